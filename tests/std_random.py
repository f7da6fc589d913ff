"""std::mt19937_64 and libc++'s std::uniform_real_distribution<double>, so that
the tests draw the very samples the reference's tests draw (the reference is
built with clang + libc++, Makefile:8-10).

mt19937_64 is the standard's generator ([rand.predef]: the 10000th value of a
default-constructed one is 9981545732273789042, checked in
tests/test_host_arithmetic.py).  libc++'s generate_canonical<double, 53> with a
64-bit engine takes one draw: double(draw - min) / 2^64; the distribution
returns (b - a) * that + a.
"""

_MASK = (1 << 64) - 1


class Mt19937_64:
  N, M = 312, 156

  def __init__(self, seed=5489):
    mt = [0] * self.N
    mt[0] = seed & _MASK
    for i in range(1, self.N):
      mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & _MASK
    self.mt = mt
    self.index = self.N

  def _twist(self):
    mt, n, m = self.mt, self.N, self.M
    upper, lower = 0xFFFFFFFF80000000, 0x7FFFFFFF
    for i in range(n):
      x = (mt[i] & upper) | (mt[(i + 1) % n] & lower)
      xa = x >> 1
      if x & 1:
        xa ^= 0xB5026F5AA96619E9
      mt[i] = mt[(i + m) % n] ^ xa
    self.index = 0

  def __call__(self):
    if self.index >= self.N:
      self._twist()
    y = self.mt[self.index]
    self.index += 1
    y ^= (y >> 29) & 0x5555555555555555
    y ^= (y << 17) & 0x71D67FFFEDA60000
    y ^= (y << 37) & 0xFFF7EEE000000000
    y ^= y >> 43
    return y & _MASK


def uniform_real(random, a, b):
  """std::uniform_real_distribution<double>(a, b)(random), libc++."""
  canonical = float(random()) / 18446744073709551616.0
  return (b - a) * canonical + a
