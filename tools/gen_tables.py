#!/usr/bin/env python3
"""Derives the numerical tables the hot path needs, from first principles.

Nothing here reads the reference.  The tables are *re-derived* with exact
rational / 60-digit arithmetic and rounded once to binary64:

* Newhall C-matrices (Чебышёв error row + monomial basis), degrees 3..17,
  8 divisions, velocity weight w = 0.4: X. X. Newhall, "Numerical
  Representation of Planetary Ephemerides", Celestial Mechanics 45:305-310,
  1989, equations (5)-(8).  The reference derives the same matrices in
  mathematica/newhall.wl and ships them as
  numerics/newhall_matrices.mathematica.h.
* Legendre normalisation factors sqrt((2-δ_m0)(2n+1)(n-m)!/(n+m)!) for
  n <= 50 (reference: numerics/legendre_normalization_factor.mathematica.h).
* max over [-1,1] of |normalised associated Legendre function| for n <= 50
  (reference: numerics/max_abs_normalized_associated_legendre_function
  .mathematica.h).

Output: a C header with `static const double` arrays, written to the paths
given on the command line (the product and the oracle each keep their own
generated copy).  `tests/test_tables.py` compares the result with the reference
headers when /root/reference is present.
"""
import sys
from fractions import Fraction

import mpmath
import numpy
import scipy.special

DIVISIONS = 8
W = Fraction(4, 10)
MIN_DEGREE, MAX_DEGREE = 3, 17
LEGENDRE_MAX = 50


# ---------------------------------------------------------------- Newhall ---
def cheb_coeffs(n):
  """Coefficients (ascending powers) of T_n as exact integers."""
  t0, t1 = [1], [0, 1]
  if n == 0:
    return t0
  for _ in range(n - 1):
    t2 = [0] + [2 * c for c in t1]
    for i, c in enumerate(t0):
      t2[i] -= c
    t0, t1 = t1, t2
  return t1


def poly_eval(c, x):
  r = Fraction(0)
  for a in reversed(c):
    r = r * x + a
  return r


def poly_deriv(c):
  return [i * a for i, a in enumerate(c)][1:] or [0]


def solve(a, b):
  """Exact Gauss-Jordan: returns a^-1 b for square a (lists of Fractions)."""
  n = len(a)
  m = [row[:] + brow[:] for row, brow in zip(a, b)]
  for col in range(n):
    piv = next(r for r in range(col, n) if m[r][col] != 0)
    m[col], m[piv] = m[piv], m[col]
    inv = 1 / m[col][col]
    m[col] = [x * inv for x in m[col]]
    for r in range(n):
      if r != col and m[r][col] != 0:
        f = m[r][col]
        m[r] = [x - f * y for x, y in zip(m[r], m[col])]
  return [row[n:] for row in m]


def newhall_c(degree):
  """Rows 0..degree of C1^-1 C2 (Newhall's c_k), exact."""
  cheb = [cheb_coeffs(j) for j in range(degree + 1)]
  dcheb = [poly_deriv(c) for c in cheb]
  xs = [Fraction(1) - Fraction(2 * i, DIVISIONS) for i in range(DIVISIONS + 1)]
  t = []
  for x in xs:  # Largest time first, as in Newhall.
    t.append([poly_eval(c, x) for c in cheb])
    t.append([poly_eval(c, x) for c in dcheb])
  rows = 2 * (DIVISIONS + 1)
  wdiag = [Fraction(1), W * W] * (DIVISIONS + 1)
  n = degree + 1
  ttw = [[t[r][i] * wdiag[r] for r in range(rows)] for i in range(n)]
  upper_left = [[sum(ttw[i][r] * t[r][j] for r in range(rows))
                 for j in range(n)] for i in range(n)]
  upper_right = [[poly_eval(cheb[i], Fraction(1)),
                  poly_eval(dcheb[i], Fraction(1)),
                  poly_eval(cheb[i], Fraction(-1)),
                  poly_eval(dcheb[i], Fraction(-1))] for i in range(n)]
  c1 = [upper_left[i] + upper_right[i] for i in range(n)]
  for k in range(4):
    c1.append([upper_right[i][k] for i in range(n)] + [Fraction(0)] * 4)
  c2 = [ttw[i][:] for i in range(n)]
  for k in (0, 1, rows - 2, rows - 1):
    c2.append([Fraction(int(r == k)) for r in range(rows)])
  c = solve(c1, c2)
  return c[:n], cheb


def machine(x):
  """binary64 value of an exact rational the way the reference's table
  generator (Mathematica's CForm of an exact matrix) produces it: numerator and
  denominator are each rounded to binary64, then divided.  For the large
  rationals of degree >= 9 this differs by 1 ULP from the correctly rounded
  quotient in 274 entries; parity with the reference needs its values."""
  return float(x.numerator) / float(x.denominator)


def newhall_tables():
  out = {}
  for degree in range(MIN_DEGREE, MAX_DEGREE + 1):
    c, cheb = newhall_c(degree)
    mono = []
    for k in range(degree + 1):
      row = [sum(c[n][col] * (cheb[n][k] if k < len(cheb[n]) else 0)
                 for n in range(degree + 1))
             for col in range(2 * (DIVISIONS + 1))]
      mono.append(row)
    out[degree] = ([machine(x) for x in c[degree]],
                   [[machine(x) for x in row] for row in mono])
  return out


# --------------------------------------------------------------- Legendre ---
def legendre_tables(max_n):
  mpmath.mp.dps = 80
  norm = {}
  maxabs = {}
  for n in range(max_n + 1):
    # Exact integer coefficients of 2^n n! P_n via Rodrigues are overkill; use
    # the Bonnet recurrence on exact rationals.
    pass
  # P_n as exact rational coefficient lists.
  p = [[Fraction(1)], [Fraction(0), Fraction(1)]]
  for n in range(2, max_n + 1):
    a = [Fraction(0)] + [Fraction(2 * n - 1, n) * c for c in p[n - 1]]
    for i, c in enumerate(p[n - 2]):
      a[i] -= Fraction(n - 1, n) * c
    p.append(a)
  for n in range(max_n + 1):
    d = p[n]
    for m in range(n + 1):
      # d = d^m/dx^m P_n, exact.
      f = mpmath.mpf(2 - (m == 0)) * (2 * n + 1) * mpmath.factorial(n - m) / \
          mpmath.factorial(n + m)
      nf = mpmath.sqrt(f)
      norm[(n, m)] = float(nf)
      # Critical points of (1-x^2)^(m/2) d(x): (1-x^2) d'(x) - m x d(x) = 0.
      dd = poly_deriv(d)
      crit = [Fraction(0)] * (max(len(dd) + 2, len(d) + 1))
      for i, c in enumerate(dd):
        crit[i] += c
        crit[i + 2] -= c
      for i, c in enumerate(d):
        crit[i + 1] -= m * c
      while len(crit) > 1 and crit[-1] == 0:
        crit.pop()

      def fval(x):
        s = mpmath.mpf(0)
        for c in reversed(d):
          s = s * x + mpmath.mpf(c.numerator) / c.denominator
        return abs(s) * (1 - x * x) ** (mpmath.mpf(m) / 2)

      best = mpmath.mpf(0)
      candidates = [mpmath.mpf(1), mpmath.mpf(-1)] if m == 0 else []
      if len(crit) > 1:
        # Locate the local maxima on a dense binary64 grid, then polish each
        # promising one with Newton's method on the critical polynomial.
        grid = numpy.cos(numpy.linspace(0.0, numpy.pi, 40001))
        vals = numpy.abs(scipy.special.lpmv(m, n, grid))
        top = vals.max()
        idx = [i for i in range(1, len(grid) - 1)
               if vals[i] >= vals[i - 1] and vals[i] >= vals[i + 1] and
               vals[i] > 0.98 * top]
        mcrit = [mpmath.mpf(c.numerator) / c.denominator for c in crit]
        mdcrit = [i * c for i, c in enumerate(mcrit)][1:]

        def horner(cs, x):
          s = mpmath.mpf(0)
          for c in reversed(cs):
            s = s * x + c
          return s

        for i in idx:
          x = mpmath.mpf(float(grid[i]))
          for _ in range(60):
            dx = horner(mcrit, x) / horner(mdcrit, x)
            x -= dx
            if abs(dx) < mpmath.mpf(10) ** -60:
              break
          assert abs(x - float(grid[i])) < 1e-3, (n, m, x, grid[i])
          candidates.append(x)
      for x in candidates:
        best = max(best, fval(x))
      maxabs[(n, m)] = float(best * nf)
      d = dd
  return norm, maxabs


# ----------------------------------------------------------------- output ---
def fmt(x):
  return float(x).hex()


def emit(path, newhall, norm, maxabs, max_n):
  with open(path, 'w') as f:
    f.write('// GENERATED by tools/gen_tables.py -- do not edit.\n')
    f.write('// Re-derived from first principles (Newhall 1989; fully '
            'normalised associated\n// Legendre functions); binary64 values '
            'printed as hex floats.\n#pragma once\n\n')
    f.write('#define PB_NEWHALL_DIVISIONS %d\n' % DIVISIONS)
    f.write('#define PB_NEWHALL_MIN_DEGREE %d\n' % MIN_DEGREE)
    f.write('#define PB_NEWHALL_MAX_DEGREE %d\n\n' % MAX_DEGREE)
    # Flattened layout: for each degree, [row][division][q,v].
    f.write('// Offsets of each degree in pb_newhall_monomial: rows 0..degree, '
            'each 9 (q, v) pairs.\n')
    offs = []
    flat = []
    err_flat = []
    for degree in range(MIN_DEGREE, MAX_DEGREE + 1):
      err, mono = newhall[degree]
      offs.append(len(flat))
      for row in mono:
        flat.extend(row)
      err_flat.extend(err)
    f.write('static const int pb_newhall_monomial_offset[%d] = {%s};\n' %
            (len(offs), ', '.join(map(str, offs))))
    f.write('static const double pb_newhall_monomial[%d] = {\n' % len(flat))
    for i in range(0, len(flat), 4):
      f.write('    ' + ', '.join(fmt(x) for x in flat[i:i + 4]) + ',\n')
    f.write('};\n')
    f.write('// Last row (degree) of the Chebyshev C matrix for each degree: '
            '18 doubles each.\n')
    f.write('static const double pb_newhall_chebyshev_last_row[%d] = {\n' %
            len(err_flat))
    for i in range(0, len(err_flat), 4):
      f.write('    ' + ', '.join(fmt(x) for x in err_flat[i:i + 4]) + ',\n')
    f.write('};\n\n')
    size = (max_n + 1) * (max_n + 2) // 2
    f.write('#define PB_LEGENDRE_MAX_DEGREE %d\n' % max_n)
    for name, tab in (('pb_legendre_normalization_factor', norm),
                      ('pb_max_abs_normalized_legendre', maxabs)):
      f.write('// Lower-triangular, index n(n+1)/2 + m.\n')
      f.write('static const double %s[%d] = {\n' % (name, size))
      vals = [tab[(n, m)] for n in range(max_n + 1) for m in range(n + 1)]
      for i in range(0, size, 4):
        f.write('    ' + ', '.join(fmt(x) for x in vals[i:i + 4]) + ',\n')
      f.write('};\n')


def main():
  newhall = newhall_tables()
  norm, maxabs = legendre_tables(LEGENDRE_MAX)
  for path in sys.argv[1:]:
    emit(path, newhall, norm, maxabs, LEGENDRE_MAX)


if __name__ == '__main__':
  main()
