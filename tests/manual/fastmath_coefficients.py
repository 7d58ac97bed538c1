"""Generates the polynomial coefficients of microsimulation_b200/csrc/fastmath.cuh (mpmath, Chebyshev interpolation;
near-minimax, which is all these degrees need).  python tests/manual/fastmath_coefficients.py"""
import struct

from mpmath import atanh, chebyfit, exp, log, mp, mpf, sqrt

mp.dps = 60
# exp(r) = 1 + r + r^2 q(r) on |r| <= ln(2)/2 (a little more): q of degree 9
h = log(2) / 2 * (1 + mpf(2) ** -9)


def q(r):
    if abs(r) < mpf(10) ** -12:
        return mpf(1) / 2 + r / 6 + r * r / 24
    return (exp(r) - 1 - r) / (r * r)


c, err = chebyfit(q, [-h, h], 10, error=True)
print("FM_EXPQ (lowest degree first), max error", float(err))
print(", ".join(float(x).hex() for x in reversed(c)))
# log: atanh(s)/s = 1 + z G(z), z = s^2, |s| <= (sqrt2 - 1)/(sqrt2 + 1)
smax = (sqrt(2) - 1) / (sqrt(2) + 1) * (1 + mpf(2) ** -9)


def G(z):
    if z < mpf(10) ** -20:
        return mpf(1) / 3 + z / 5
    s = sqrt(z)
    return (atanh(s) / s - 1) / z


c, err = chebyfit(G, [0, smax ** 2], 7, error=True)
print("FM_LOGG, max error", float(err))
print(", ".join(float(x).hex() for x in reversed(c)))
ln2 = log(2)
hi = float(ln2)
bits = struct.unpack("<Q", struct.pack("<d", hi))[0] & ~((1 << 21) - 1)
hi32 = struct.unpack("<d", struct.pack("<Q", bits))[0]
print("FM_K: log2(e)", float(1 / ln2).hex(), "ln2 hi/lo", hi.hex(), float(ln2 - mpf(hi)).hex(), "ln2 hi32/lo", hi32.hex(),
      float(ln2 - mpf(hi32)).hex())
