#!/usr/bin/env python3
"""Checks the Markstein-style exact division used by k_sp_pixels (surfel.cu: div100):
   q1 = RN(a*c), r = fma(-q1, 100, a), q = fma(r, c, q1), c = RN(1/100)  ==  RN(a / 100.0)
for float-valued dividends a = (double)(idiff*idiff).  fma is emulated exactly with fractions."""
import random
import struct
from fractions import Fraction


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def fma(x, y, z):
    return float(Fraction(x) * Fraction(y) + Fraction(z))


def main(n=300000):
    random.seed(0)
    c = 1.0 / 100.0
    bad = 0
    for t in range(n):
        if t % 3 == 0:
            a = f32(random.uniform(0, 65025))
        elif t % 3 == 1:
            a = f32(random.uniform(0, 4.0))
        else:
            d = f32(f32(random.uniform(0, 255)) - float(random.randint(0, 255)))
            a = f32(d * d)
        q1 = a * c
        q = fma(fma(-q1, 100.0, a), c, q1)
        bad += q != a / 100.0
    print("mismatches:", bad, "of", n)
    return bad


if __name__ == "__main__":
    raise SystemExit(1 if main() else 0)
