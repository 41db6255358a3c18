"""Rational oversampling factor: the ``pfb.rational.Rational`` the reference
harness uses (``verify/test_backends.py:51-54``, ``verify/purity.py:79-86``) and
the ``struct('nu',..,'de',..)`` of the Matlab code (normalize.m:17, multiply.m:17)."""
from __future__ import annotations

from fractions import Fraction


class Rational:
    def __init__(self, nu, de):
        self.nu, self.de = int(nu), int(de)
        if self.nu < 1 or self.de < 1:
            raise ValueError("oversampling factor must be positive")

    @classmethod
    def from_str(cls, s):
        if isinstance(s, Rational):
            return s
        if hasattr(s, "nu") and hasattr(s, "de"):
            return cls(s.nu, s.de)
        if isinstance(s, (tuple, list)):
            return cls(*s)
        parts = str(s).split("/")
        if len(parts) == 1:
            return cls(int(parts[0]), 1)
        return cls(int(parts[0]), int(parts[1]))

    def normalize(self, n):
        """n * de / nu  (an int whenever it divides exactly)."""
        v = Fraction(self.de, self.nu) * n
        return int(v) if v.denominator == 1 else float(v)

    def multiply(self, n):
        v = Fraction(self.nu, self.de) * n
        return int(v) if v.denominator == 1 else float(v)

    def __float__(self):
        return self.nu / self.de

    def __eq__(self, other):
        other = Rational.from_str(other)
        return self.nu * other.de == other.nu * self.de

    def __hash__(self):
        return hash(Fraction(self.nu, self.de))

    def __str__(self):
        return f"{self.nu}/{self.de}"

    def __repr__(self):
        return f"Rational({self.nu}, {self.de})"
