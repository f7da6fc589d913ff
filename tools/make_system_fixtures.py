#!/usr/bin/env python3
"""Turns the reference's gravity-model / initial-state configuration files into
the compact JSON files under principia_b200/data/ that the oracle, the tests and
bench.py load (the reference tree does not exist on the GPU box).

Run in the build container only:
  python tools/make_system_fixtures.py

The conversion restates, in binary64 exactly as the reference does:
* quantities/parser_body.hpp:83-320 (ParseQuantity: strtod(magnitude) *
  unit.scale, quotient split at the last '/', products at blanks, '^' via pow);
* astronomy/time_scales_body.hpp (JD epochs -> seconds from J2000);
* physics/solar_system_body.hpp:100-156 (bodies keyed by name, alphabetical;
  geopotentials truncated to degree 30), :586-645 (body parameters; min radius
  defaults to the mean radius);
* physics/oblate_body_body.hpp:30-100 (j / j2 -> normalised cos, zonalness).
Keplerian initial states (TRAPPIST-1) are converted to Cartesian, Jacobi
hierarchy with the star as primary, by a plain two-body formula; that
conversion is ours (the reference's HierarchicalSystem/KeplerOrbit is not
restated), which is fine for a synthetic ensemble seed and is stated in
DESIGN.md.
"""
import json
import math
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/astronomy'
sys.path.insert(0, os.path.join(REPO, 'tests'))

PI = 3.14159265358979323846264338327950288419716939937511
UNITS = {
    '': 1.0,
    'mm': 1e-3 * 1.0, 'cm': 1e-2 * 1.0, 'm': 1.0, 'km': 1e3 * 1.0,
    'R🜨': 6.3781e6, 'R☉': 6.957e8, 'au': 149597870700.0,
    'kg': 1.0,
    'ms': 1e-3 * 1.0, 's': 1.0, 'min': 60.0, 'h': 60.0 * 60.0,
    'd': 24 * (60.0 * 60.0),
    'GM🜨': 3.986004e14, 'GM☉': 1.3271244e20,
    'deg': PI / 180 * 1.0, '°': PI / 180 * 1.0, 'rad': 1.0,
}


def parse_unit(s):
  return UNITS[s]


def parse_exponentiation_unit(s):
  caret = s.find('^')
  if caret < 0:
    return parse_unit(s)
  left = s[:caret].rstrip(' ')
  right = int(s[caret + 1:].strip(' '))
  return math.pow(parse_unit(left), right)


def parse_product_unit(s):
  start = 0
  while True:
    first_blank = s.find(' ', start)
    if first_blank < 0:
      return parse_exponentiation_unit(s)
    m = re.compile(r'[^ ]').search(s, first_blank + 1)
    first_nonblank = m.start() if m else -1
    last_nonblank = len(s[:first_blank].rstrip(' ')) - 1
    if ((first_nonblank < 0 or s[first_nonblank] != '^') and
        (last_nonblank < 0 or s[last_nonblank] != '^')):
      break
    start = first_blank + 1
  left = parse_exponentiation_unit(s[:last_nonblank + 1])
  right = parse_product_unit(s[first_nonblank:])
  return left * right


def parse_quotient_unit(s):
  last_slash = s.rfind('/')
  if last_slash < 0:
    return parse_product_unit(s)
  right_s = s[last_slash + 1:].lstrip(' ')
  left_s = s[:last_slash].rstrip(' ')
  left = 1.0 if left_s == '' else parse_quotient_unit(left_s)
  right = parse_exponentiation_unit(right_s)
  return left / right


def parse_quantity(s):
  m = re.match(r'\s*[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?', s)
  magnitude = float(m.group(0))
  unit_string = s[m.end():].strip(' ')
  return magnitude * parse_quotient_unit(unit_string)


def parse_tt(s):
  m = re.fullmatch(r'JD(\d+)\.(\d+)', s)
  day = int(m.group(1))
  numerator = int(m.group(2))
  denominator = 10 ** len(m.group(2))
  # astronomy/date_time_body.hpp: day - 2451545 + numerator / denominator, in
  # days, times Day.
  return (day - 2451545 + numerator / denominator) * UNITS['d']


# ------------------------------------------------- protobuf text, minimal ---
TOKEN = re.compile(r'\s*(?:#[^\n]*\n\s*)*("(?:[^"\\]|\\.)*"|[{}:,]|[^\s{}:,"#]+)')


def tokenize(text):
  pos = 0
  out = []
  text += '\n'
  while True:
    m = TOKEN.match(text, pos)
    if not m:
      break
    out.append(m.group(1))
    pos = m.end()
  return out


def parse_message(tokens, i):
  """Returns (list of (key, value)), next index."""
  fields = []
  while i < len(tokens) and tokens[i] != '}':
    if tokens[i] == ',':
      i += 1
      continue
    key = tokens[i]
    i += 1
    if tokens[i] == ':':
      i += 1
    if tokens[i] == '{':
      sub, i = parse_message(tokens, i + 1)
      assert tokens[i] == '}'
      i += 1
      fields.append((key, sub))
    else:
      v = tokens[i]
      i += 1
      if v.startswith('"'):
        v = v[1:-1]
      fields.append((key, v))
  return fields, i


def load(path):
  tokens = tokenize(open(path, encoding='utf-8').read())
  msg, i = parse_message(tokens, 0)
  assert i == len(tokens)
  return msg


def get(msg, key, default=None):
  for k, v in msg:
    if k == key:
      return v
  return default


def get_all(msg, key):
  return [v for k, v in msg if k == key]


def legendre_normalization():
  from test_tables import parse_generated
  gen = parse_generated(os.path.join(REPO, 'oracle', 'generated_tables.h'))
  return gen['pb_legendre_normalization_factor']


def make_body(msg, norm, max_degree=30):
  body = {'name': get(msg, 'name'),
          'mu': parse_quantity(get(msg, 'gravitational_parameter'))}
  if get(msg, 'mean_radius') is not None:
    mean = get(msg, 'mean_radius')
    body['rotating'] = True
    body['min_radius'] = parse_quantity(get(msg, 'min_radius', mean))
    body['mean_radius'] = parse_quantity(mean)
    body['reference_angle'] = parse_quantity(get(msg, 'reference_angle'))
    body['reference_instant'] = parse_tt(get(msg, 'reference_instant'))
    body['angular_frequency'] = parse_quantity(get(msg, 'angular_frequency'))
    body['right_ascension'] = parse_quantity(get(msg, 'axis_right_ascension'))
    body['declination'] = parse_quantity(get(msg, 'axis_declination'))
  else:
    body['rotating'] = False
  body['oblate'] = False
  j2 = get(msg, 'j2')
  geo = get(msg, 'geopotential')
  cos, sin = {}, {}
  if j2 is not None:
    cos[(2, 0)] = -float(j2) / norm[2 * 3 // 2]
    degree, zonal = 2, True
  elif geo is not None:
    rows = [r for r in get_all(geo, 'row')
            if int(get(r, 'degree')) <= max_degree]
    degree = 0
    for row in rows:
      n = int(get(row, 'degree'))
      degree = max(degree, n)
      for col in get_all(row, 'column'):
        m = int(get(col, 'order'))
        c = float(get(col, 'cos', '0'))
        if m == 0 and get(col, 'j') is not None:
          c = -float(get(col, 'j')) / norm[n * (n + 1) // 2]
        cos[(n, m)] = c
        sin[(n, m)] = float(get(col, 'sin', '0'))
    zonal = all(cos.get((n, m), 0.0) == 0 and sin.get((n, m), 0.0) == 0
                for n in range(degree + 1) for m in range(1, n + 1))
    if not rows:
      geo = None
  if j2 is not None or geo is not None:
    body['oblate'] = True
    body['degree'] = degree
    body['zonal'] = zonal
    body['reference_radius'] = parse_quantity(get(msg, 'reference_radius'))
    body['cos'] = [cos.get((n, m), 0.0)
                   for n in range(degree + 1) for m in range(n + 1)]
    body['sin'] = [sin.get((n, m), 0.0)
                   for n in range(degree + 1) for m in range(n + 1)]
  return body


def kepler_to_cartesian(mu, e, period, inc, node, argp, mean_anomaly):
  n = 2 * math.pi / period
  a = (mu / (n * n)) ** (1.0 / 3.0)
  ecc_anomaly = mean_anomaly
  for _ in range(100):
    ecc_anomaly -= ((ecc_anomaly - e * math.sin(ecc_anomaly) - mean_anomaly) /
                    (1 - e * math.cos(ecc_anomaly)))
  x = a * (math.cos(ecc_anomaly) - e)
  y = a * math.sqrt(1 - e * e) * math.sin(ecc_anomaly)
  r = a * (1 - e * math.cos(ecc_anomaly))
  vx = -math.sqrt(mu * a) / r * math.sin(ecc_anomaly)
  vy = math.sqrt(mu * a * (1 - e * e)) / r * math.cos(ecc_anomaly)

  def rot(px, py):
    co, so = math.cos(node), math.sin(node)
    cw, sw = math.cos(argp), math.sin(argp)
    ci, si = math.cos(inc), math.sin(inc)
    return [(co * cw - so * sw * ci) * px + (-co * sw - so * cw * ci) * py,
            (so * cw + co * sw * ci) * px + (-so * sw + co * cw * ci) * py,
            (sw * si) * px + (cw * si) * py]
  return rot(x, y), rot(vx, vy)


def make_system(gravity_file, state_file):
  norm = legendre_normalization()
  gm = get(load(os.path.join(REF, gravity_file)), 'gravity_model')
  st = get(load(os.path.join(REF, state_file)), 'initial_state')
  bodies = {b['name']: b for b in
            (make_body(m, norm) for m in get_all(gm, 'body'))}
  out = {'gravity_model': gravity_file, 'initial_state': state_file,
         'epoch': parse_tt(get(st, 'epoch'))}
  cart = get(st, 'cartesian')
  if cart is not None:
    for b in get_all(cart, 'body'):
      body = bodies[get(b, 'name')]
      body['q'] = [parse_quantity(get(b, k)) for k in 'xyz']
      body['v'] = [parse_quantity(get(b, k)) for k in ('vx', 'vy', 'vz')]
  else:
    kep = get(st, 'keplerian')
    entries = get_all(kep, 'body')
    primary = [get(b, 'name') for b in entries if get(b, 'parent') is None][0]
    # Jacobi-like: each planet orbits the barycentre of everything interior.
    planets = [b for b in entries if get(b, 'parent') is not None]
    planets.sort(key=lambda b: parse_quantity(get(get(b, 'elements'),
                                                  'period')))
    bodies[primary]['q'] = [0.0, 0.0, 0.0]
    bodies[primary]['v'] = [0.0, 0.0, 0.0]
    inner = [primary]
    for b in planets:
      el = get(b, 'elements')
      name = get(b, 'name')
      mu_in = sum(bodies[k]['mu'] for k in inner)
      mu = mu_in + bodies[name]['mu']
      q, v = kepler_to_cartesian(
          mu, float(get(el, 'eccentricity')),
          parse_quantity(get(el, 'period')),
          parse_quantity(get(el, 'inclination')),
          parse_quantity(get(el, 'longitude_of_ascending_node')),
          parse_quantity(get(el, 'argument_of_periapsis')),
          parse_quantity(get(el, 'mean_anomaly')))
      bq = [sum(bodies[k]['mu'] * bodies[k]['q'][i] for k in inner) / mu_in
            for i in range(3)]
      bv = [sum(bodies[k]['mu'] * bodies[k]['v'][i] for k in inner) / mu_in
            for i in range(3)]
      bodies[name]['q'] = [bq[i] + q[i] for i in range(3)]
      bodies[name]['v'] = [bv[i] + v[i] for i in range(3)]
      inner.append(name)
    # Move to the barycentric frame.
    mu_tot = sum(b['mu'] for b in bodies.values())
    for i in range(3):
      bq = sum(b['mu'] * b['q'][i] for b in bodies.values()) / mu_tot
      bv = sum(b['mu'] * b['v'][i] for b in bodies.values()) / mu_tot
      for b in bodies.values():
        b['q'][i] -= bq
        b['v'][i] -= bv
  # std::map<std::string, ...> order: bytewise, equal to code point order.
  out['bodies'] = [bodies[k] for k in sorted(bodies)]
  return out


def main():
  golden = os.path.join(REPO, 'principia_b200', 'data')
  for name, gravity, state in (
      ('sol_jd2451545', 'sol_gravity_model.proto.txt',
       'sol_initial_state_jd_2451545_000000000.proto.txt'),
      # The epoch of astronomy/mercury_perihelion_test.cpp (1950-01-01).
      ('sol_jd2433282', 'sol_gravity_model.proto.txt',
       'sol_initial_state_jd_2433282_500000000.proto.txt'),
      ('trappist_jd2457000', 'trappist_gravity_model.proto.txt',
       'trappist_initial_state_jd_2457000_000000000.proto.txt')):
    system = make_system(gravity, state)
    with open(os.path.join(golden, name + '.json'), 'w') as f:
      json.dump(system, f, indent=1)
    print(name, len(system['bodies']), 'bodies,',
          sum(b['oblate'] for b in system['bodies']), 'oblate')


if __name__ == '__main__':
  main()
