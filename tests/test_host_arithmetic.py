"""CPU-only: the product's host/device arithmetic headers (compiled for the host
by tests/hostcheck) against the oracle, bit for bit.  Catches logic and
operation-order errors before GPU time is spent; the `-m gpu` tests then check
the same functions as compiled by nvcc for sm_100a.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

from principia_b200 import _capi
from principia_b200.solar_system import SolarSystem

import oracle_lib
from oracle_lib import OracleEphemeris

HERE = os.path.dirname(os.path.abspath(__file__))
_d, _i32, _vp = ctypes.c_double, ctypes.c_int32, ctypes.c_void_p
_dp, _i32p = _capi.c_double_p, _capi.c_int32_p


@pytest.fixture(scope='module')
def hc():
  directory = os.path.join(HERE, 'hostcheck')
  subprocess.check_call(
      ['/usr/bin/g++', '-std=c++17', '-O2', '-ffp-contract=off', '-mfma',
       '-fPIC', '-shared', '-w', '-o', 'libhostcheck.so', 'hostcheck.cpp'],
      cwd=directory)
  lib = ctypes.CDLL(os.path.join(directory, 'libhostcheck.so'))
  lib.hc_cr_sincos.argtypes = [_d, _dp, _dp]
  lib.hc_cr_root4.restype = _d
  lib.hc_cr_root4.argtypes = [_d]
  lib.hc_estrin.restype = _d
  lib.hc_estrin.argtypes = [_dp, _i32, _d]
  lib.hc_estrin_derivative.restype = _d
  lib.hc_estrin_derivative.argtypes = [_dp, _i32, _d]
  lib.hc_estrin_padded.argtypes = [_dp, _i32, _d, _d, _dp, _dp]
  lib.hc_chebyshev.argtypes = [_dp, _i32, _d, _d, _d, _dp, _dp]
  lib.hc_downsample.restype = ctypes.c_int64
  lib.hc_downsample.argtypes = [ctypes.c_int64, _dp, _dp, _dp, _d, _dp, _dp,
                                _dp, _dp]
  lib.hc_pow_int.restype = _d
  lib.hc_pow_int.argtypes = [_d, _i32]
  lib.hc_model_create.argtypes = [_i32, ctypes.POINTER(_capi.Body), _d,
                                  ctypes.POINTER(_vp)]
  lib.hc_model_destroy.argtypes = [_vp]
  lib.hc_model_order.argtypes = [_vp, _i32p, _i32p]
  lib.hc_model_damping.argtypes = [_vp, _i32, _i32p, _dp, _dp]
  lib.hc_geopotential_acceleration.argtypes = [_vp, _i32, _d, _dp, _i32, _dp]
  lib.hc_acceleration_on_massless_body.argtypes = [_vp, _d, _dp, _dp, _i32,
                                                   _dp, _dp]
  return lib


def make_model(hc, system, tolerance):
  bodies, keep = system.c_bodies()
  handle = _vp()
  assert hc.hc_model_create(len(system.bodies), bodies, tolerance,
                            ctypes.byref(handle)) == 0
  return handle, keep


def test_correctly_rounded_sincos_and_root(hc):
  orc = oracle_lib.library()
  rng = np.random.default_rng(7)
  xs = np.concatenate([
      rng.uniform(-10, 10, 50000), rng.uniform(-1e5, 1e5, 50000),
      rng.uniform(-2e9, 2e9, 20000), 10.0**rng.uniform(-300, 0, 5000),
      np.arange(0, 100) * (math.pi / 2), [0.0, -0.0]])
  s, c = _d(), _d()
  for x in xs:
    hc.hc_cr_sincos(x, s, c)
    assert s.value == orc.orc_cr_sin(x), x
    assert c.value == orc.orc_cr_cos(x), x
  for x in np.concatenate([10.0**rng.uniform(-30, 30, 100000),
                           rng.uniform(0.5, 2, 20000), [1.0, 16.0, 81.0]]):
    assert hc.hc_cr_root4(x) == orc.orc_root(4, x), x
  assert hc.hc_cr_root4(0.0) == 0.0
  assert hc.hc_cr_root4(math.inf) == math.inf


def test_estrin_matches_oracle(hc):
  orc = oracle_lib.library()
  rng = np.random.default_rng(3)
  for degree in range(0, 18):
    for _ in range(200):
      c = rng.normal(size=18) * 10.0**rng.uniform(-5, 5, 18)
      x = rng.uniform(-3000, 3000)
      assert (hc.hc_estrin(oracle_lib.dptr(c), degree, x) ==
              orc.orc_estrin(oracle_lib.dptr(c), degree, x))
  for k in range(0, 32):
    x = rng.uniform(0.1, 3)
    expected = x
    # Russian peasant as the reference writes it.
    def peasant(x, e):
      if e == 0:
        return 1.0
      if e == 1:
        return x
      y = peasant(x, e // 2)
      y2 = y * y
      return y2 * x if e % 2 else y2
    assert hc.hc_pow_int(x, k) == peasant(x, k)


# numerics/polynomial_evaluators_test.cpp:26-67 (PolynomialEvaluatorTest.Estrin):
# (x + 1)^degree from its binomial coefficients at the integers of
# [-degree, degree], value and derivative, EXPECT_EQ -- the oracle and the host
# build of the device header.
def test_estrin_binomial_known_answers(hc):
  orc = oracle_lib.library()
  for degree in range(1, 15):
    c = np.zeros(18)
    c[:degree + 1] = [math.comb(degree, k) for k in range(degree + 1)]
    for argument in range(-degree, degree + 1):
      x = float(argument)
      value = float((argument + 1)**degree)
      derivative = float(degree * (argument + 1)**(degree - 1))
      assert orc.orc_estrin(oracle_lib.dptr(c), degree, x) == value
      assert orc.orc_estrin_derivative(oracle_lib.dptr(c), degree,
                                       x) == derivative
      assert hc.hc_estrin(oracle_lib.dptr(c), degree, x) == value
      assert hc.hc_estrin_derivative(oracle_lib.dptr(c), degree,
                                     x) == derivative


def test_estrin_derivative_matches_oracle(hc):
  orc = oracle_lib.library()
  rng = np.random.default_rng(4)
  for degree in range(0, 18):
    for _ in range(200):
      c = rng.normal(size=18) * 10.0**rng.uniform(-5, 5, 18)
      x = rng.uniform(-3000, 3000)
      assert (hc.hc_estrin_derivative(oracle_lib.dptr(c), degree, x) ==
              orc.orc_estrin_derivative(oracle_lib.dptr(c), degree, x))


def test_chebyshev_series_match_oracle(hc):
  orc = oracle_lib.library()
  rng = np.random.default_rng(5)
  value, derivative = _d(), _d()
  for degree in range(0, 21):
    for _ in range(100):
      c = rng.normal(size=21) * 10.0**rng.uniform(-3, 3, 21)
      lower = rng.uniform(-1e6, 1e6)
      upper = lower + rng.uniform(1.0, 1e5)
      t = rng.uniform(lower, upper)
      hc.hc_chebyshev(oracle_lib.dptr(c), degree, lower, upper, t,
                      ctypes.byref(value), ctypes.byref(derivative))
      assert value.value == orc.orc_chebyshev_evaluate(
          oracle_lib.dptr(c), degree, lower, upper, t)
      assert derivative.value == orc.orc_chebyshev_evaluate_derivative(
          oracle_lib.dptr(c), degree, lower, upper, t)


def test_downsampling_matches_oracle(hc):
  """DownsamplerAppend (pb_downsampling.cuh, compiled for the host) against the
  oracle's DiscreteTrajectorySegment, bit for bit: the reference's circle and
  straight line, a Kepler ellipse with an uneven cadence, several tolerances."""
  import trajectory_factories
  from test_oracle_kats import oracle_downsample
  cases = [trajectory_factories.circular_timeline(3.0, 2.0, 10e-3, 0.0, 10.0),
           trajectory_factories.linear_timeline([1.0, 2.0, 3.0], 10e-3, 0.0,
                                                10.0)]
  rng = np.random.default_rng(21)
  mu, a, e = 3.986e14, 8.0e6, 0.3
  mean_motion = math.sqrt(mu / a**3)
  times = np.cumsum(rng.uniform(5.0, 15.0, 600))
  q, v = [], []
  for t in times:
    M = mean_motion * t
    E = M
    for _ in range(50):
      E = M + e * math.sin(E)
    r = a * (1 - e * math.cos(E))
    q.append([a * (math.cos(E) - e), a * math.sqrt(1 - e * e) * math.sin(E),
              0.1 * t])
    v.append([-a * a * mean_motion / r * math.sin(E),
              a * a * mean_motion / r * math.sqrt(1 - e * e) * math.cos(E),
              0.1])
  cases.append((times, np.array(q), np.array(v)))
  sizes = []
  for times, q, v in cases:
    for tolerance in (1e-3, 10.0, 1e4, -1.0):
      expected = oracle_downsample(times, q, v, tolerance)
      n = len(times)
      kept_times, kept_errors = np.zeros(n), np.zeros(n)
      kept_q, kept_v = np.zeros((n, 3)), np.zeros((n, 3))
      dptr = oracle_lib.dptr
      count = hc.hc_downsample(
          n, dptr(np.ascontiguousarray(times)), dptr(np.ascontiguousarray(q)),
          dptr(np.ascontiguousarray(v)), tolerance, dptr(kept_times),
          dptr(kept_q), dptr(kept_v), dptr(kept_errors))
      assert count == len(expected['times'])
      sizes.append(count)
      assert kept_times[:count].tobytes() == expected['times'].tobytes()
      assert kept_q[:count].tobytes() == expected['q'].tobytes()
      assert kept_v[:count].tobytes() == expected['v'].tobytes()
      assert kept_errors[:count].tobytes() == expected['errors'].tobytes()
  assert sizes[0] == 49 and sizes[3] == 1001 and sizes[4] == 2
  assert sizes[11] == 600 and 2 < sizes[9] < 600


def test_padded_estrin_tree_equals_the_reference_tree(hc):
  """EvaluateSegmentPadded (one tree for all degrees, -0.0 padding) against
  the run-time tree, bit for bit: every degree, both signs of x, zero and
  negative-zero coefficients."""
  rng = np.random.default_rng(11)
  reference = np.zeros(3)
  padded = np.zeros(3)
  for degree in range(0, 18):
    for trial in range(300):
      c = rng.normal(size=(18, 3)) * 10.0**rng.uniform(-8, 8, (18, 3))
      if trial % 3 == 1:
        # Exact zeros of both signs among the coefficients.
        mask = rng.random((18, 3)) < 0.5
        c[mask] = rng.choice([0.0, -0.0], size=int(mask.sum()))
      if trial % 3 == 2:
        c[:] = rng.choice([0.0, -0.0], size=(18, 3))
      c[degree + 1:, :] = -0.0
      origin = rng.uniform(-1e6, 1e6)
      t = origin + rng.choice([-1.0, 1.0, 0.0]) * rng.uniform(0, 3000)
      # The device layout is [xyz][18].
      c = np.ascontiguousarray(c.T)
      hc.hc_estrin_padded(oracle_lib.dptr(c), degree, origin, t,
                          oracle_lib.dptr(reference), oracle_lib.dptr(padded))
      assert reference.tobytes() == padded.tobytes(), (degree, trial)


def test_model_order_and_dampings_match_oracle(hc):
  system = SolarSystem.from_fixture('sol_jd2451545')
  oracle = OracleEphemeris(system)
  handle, keep = make_model(hc, system, 2.0**-24)
  order = np.zeros(len(system.bodies), dtype=np.int32)
  n_oblate = _i32()
  hc.hc_model_order(handle, order.ctypes.data_as(_i32p), n_oblate)
  expected_order, expected_oblate = oracle.internal_order()
  assert list(order) == list(expected_order)
  assert n_oblate.value == expected_oblate
  for i, b in enumerate(system.bodies):
    if not b['oblate']:
      continue
    n = _i32()
    inner = np.zeros(51)
    sectoral = _d()
    assert hc.hc_model_damping(handle, i, n, oracle_lib.dptr(inner),
                               sectoral) == 0
    expected_inner, expected_sectoral = oracle.damping(i)
    assert list(inner[:n.value]) == list(expected_inner), b['name']
    assert sectoral.value == expected_sectoral
  hc.hc_model_destroy(handle)


@pytest.mark.parametrize('unrolled', [0, 1, 2, 3])
def test_geopotential_matches_oracle_bitwise(hc, unrolled):
  """0: generic loops; 1: unrolled bodies dispatched on the degree; 2: run-time
  loops with peeled orders; 3: the unrolled body of degree 10 capped at the
  limiting degree (the adaptive flow's deep body), for the bodies whose model
  has degree <= 10."""
  system = SolarSystem.from_fixture('sol_jd2451545')
  oracle = OracleEphemeris(system)
  handle, keep = make_model(hc, system, 2.0**-24)
  rng = np.random.default_rng(42)
  out = np.zeros(3)
  checked = 0
  for i, b in enumerate(system.bodies):
    if not b['oblate']:
      continue
    if unrolled == 3 and b['degree'] > 10:
      continue
    inner, sectoral = oracle.damping(i)
    radius = b['mean_radius']
    # Distances from just above the surface to beyond every threshold, so that
    # every limiting degree, the damped zones and the zonal switch are hit.
    finite = [x for x in inner if math.isfinite(x)] + [radius]
    far = 4 * max(finite)
    n_points = 60 if b['degree'] > 12 else 150
    for _ in range(n_points):
      distance = math.exp(rng.uniform(math.log(1.01 * radius), math.log(far)))
      direction = rng.normal(size=3)
      r = distance * direction / np.linalg.norm(direction)
      t = rng.uniform(-1e6, 1e8)
      hc.hc_geopotential_acceleration(handle, i, t, oracle_lib.dptr(r),
                                      unrolled, oracle_lib.dptr(out))
      expected = oracle.geopotential_acceleration(i, t, r)
      assert list(out) == list(expected), (b['name'], distance)
      checked += 1
    # Poles and equator (r_equatorial == 0 branch).
    for r in ([0.0, 0.0, 2 * radius], [2 * radius, 0.0, 0.0]):
      r = np.array(r)
      hc.hc_geopotential_acceleration(handle, i, 0.0, oracle_lib.dptr(r),
                                      unrolled, oracle_lib.dptr(out))
      assert list(out) == list(oracle.geopotential_acceleration(i, 0.0, r))
  assert checked > 1000
  hc.hc_model_destroy(handle)


@pytest.mark.parametrize('unrolled', [0, 1])
def test_massless_acceleration_and_potential_match_oracle(hc, unrolled):
  system = SolarSystem.from_fixture('sol_jd2451545')
  oracle = OracleEphemeris(system)
  assert oracle.prolong(86400.0) == 0
  handle, keep = make_model(hc, system, 2.0**-24)
  rng = np.random.default_rng(5)
  a = np.zeros(3)
  potential = _d()
  names = system.names()
  for _ in range(300):
    t = rng.uniform(0, 86400.0)
    positions = oracle.evaluate_all_positions(t)
    centre = positions[names.index(rng.choice(['Earth', 'Moon', 'Mars',
                                                'Jupiter', 'Sun']))]
    distance = math.exp(rng.uniform(math.log(7e6), math.log(1e9)))
    direction = rng.normal(size=3)
    q = centre + distance * direction / np.linalg.norm(direction)
    error = hc.hc_acceleration_on_massless_body(
        handle, t, oracle_lib.dptr(positions), oracle_lib.dptr(q), unrolled,
        oracle_lib.dptr(a), potential)
    expected, status = oracle.acceleration_on_massless_bodies(
        t, q.reshape(3, 1))
    assert list(a) == list(expected.ravel())
    assert error == status
    assert potential.value == oracle.potential(t, q.reshape(3, 1))[0]
  # A collision: inside the Earth.
  positions = oracle.evaluate_all_positions(0.0)
  q = positions[names.index('Earth')] + np.array([1e6, 0.0, 0.0])
  error = hc.hc_acceleration_on_massless_body(
      handle, 0.0, oracle_lib.dptr(positions), oracle_lib.dptr(q), unrolled,
      oracle_lib.dptr(a), potential)
  assert error == _capi.PB_OUT_OF_RANGE
  hc.hc_model_destroy(handle)


def test_std_mt19937_64():
  """[rand.predef]: the 10000th consecutive invocation of a default-constructed
  std::mt19937_64 produces 9981545732273789042."""
  from std_random import Mt19937_64
  random = Mt19937_64()
  for _ in range(9999):
    random()
  assert random() == 9981545732273789042


# numerics/sin_cos_test.cpp:363-413 (HardRounding, HardReduction): arguments
# whose sine or cosine is hard to round, AlmostEquals(·, 0), for the oracle
# (binary128) and the host build of the device's CrSinCos (|x| < 2^31 only).
def test_sin_cos_hard_cases(hc):
  orc = oracle_lib.library()
  sin_cases = [
      (1.777288458404935767021016, 0.9787561457198967196367773),
      (5.528810471911395296729097, -0.6848332450871304488693046),
      (2.670333644894535396474566, 0.4540084183741445456039384),
      (-2.496680544289484160458414, -0.6011282027544306294509797),
      (3.348980952786005715893225, -0.2059048676737040700634683),
      (3.523452575387961971387085, -0.3726470704519433130297035),
      (1.881458091523454001503524, 0.9521314843257784876761001),
      (-1.763163156774038675678185, -0.9815544881044536151825223),
      (-2.58105062034143273308473, -0.5316453603071467637339815),
      (1.657419885978818285821035, 0.99625052493662308306103561),
      (5.094301519947547873812255, -0.9279535374988051033005616),
      (5.262137362438826571064965, -0.8526560125576488347044409),
      (float.fromhex('0x5ad7bcep-8'), 0.99999999999999999818),
  ]
  cos_cases = [
      (3.912491942337291916942377, -0.7172843528140595004137653),
      (1.486604973422413600303571, 0.0840919279825555407437241),
      (-6.265702600230396157598989, 0.99984718137127853720984932),
      (-3.885819786017697730073905, -0.7356116652133562472394118),
      (-5.026994177012682030181168, 0.3094410694753661206223057),
      (0.2388111698570396512764091, 0.9716198764286143041422587),
      (float.fromhex('0x5ad7bcep-8'), 1.9060601347136373148e-9),
  ]
  s, c = ctypes.c_double(), ctypes.c_double()
  for x, expected in sin_cases:
    assert orc.orc_cr_sin(x) == expected, x
    hc.hc_cr_sincos(x, ctypes.byref(s), ctypes.byref(c))
    assert s.value == expected, x
  for x, expected in cos_cases:
    assert orc.orc_cr_cos(x) == expected, x
    hc.hc_cr_sincos(x, ctypes.byref(s), ctypes.byref(c))
    assert c.value == expected, x
  # Muller's bad angle (beyond the device's range).
  muller = float.fromhex('0x16ac5b262ca1ffp797')
  assert orc.orc_cr_sin(muller) == 1.0
  assert orc.orc_cr_cos(muller) == -4.687165924254627611122582801963884e-19
