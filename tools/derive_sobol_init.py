#!/usr/bin/env python3
"""Derive the Joe-Kuo initialiser set (degree s, polynomial coefficients a, initial m_1..m_s per
dimension) from a 32 x NDIM table of Sobol direction numbers, and write it in the format of the
public `joe-kuo-old.1111` file (d s a m_i).

Development-time tool.  Input table: the reference's direction numbers, read through the compiled
reference oracle (oracle/_ref/libcfref.so, ref_sobol_dirnum -> getjkDir(), sobol.h:28,
sobol.cpp:16-3672).  Output: compfinance_b200/data/joe_kuo_old_1111.txt, from which the product
regenerates the full table with the published recurrence (Joe & Kuo 2003; Bratley & Fox Alg. 659):

    m_i = 2 a_1 m_{i-1} ^ 4 a_2 m_{i-2} ^ ... ^ 2^{s-1} a_{s-1} m_{i-s+1} ^ 2^s m_{i-s} ^ m_{i-s}
    v_i = m_i << (32 - i)        (i = 1..32)

so no table is copied: the 1101 x (s, a, m_1..m_s) generators are recovered by solving the
recurrence, and the regenerated table is checked bit-for-bit against the reference in tests.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import refapi  # noqa: E402

NDIM = 1101
NBIT = 32


def solve_dim(m):
    """m[1..32] -> (s, a) with smallest degree s such that the recurrence reproduces m[s+1..32]."""
    for s in range(1, 14):
        for a in range(1 << (s - 1)):
            ok = True
            for i in range(s + 1, NBIT + 1):
                x = m[i - s] ^ (m[i - s] << s)
                for k in range(1, s):
                    if (a >> (s - 1 - k)) & 1:
                        x ^= m[i - k] << k
                if x != m[i]:
                    ok = False
                    break
            if ok:
                return s, a
    raise RuntimeError("no primitive polynomial of degree <= 13 reproduces this dimension")


def main():
    ref = refapi.get()
    out = os.path.join(os.path.dirname(__file__), "..", "compfinance_b200", "data", "joe_kuo_old_1111.txt")
    lines = ["d       s       a       m_i\n"]
    for d in range(NDIM):
        v = [0] + [ref.sobol_dirnum(b, d) for b in range(NBIT)]
        m = [0] * (NBIT + 1)
        for i in range(1, NBIT + 1):
            assert v[i] % (1 << (NBIT - i)) == 0
            m[i] = v[i] >> (NBIT - i)
        if d == 0:
            assert all(m[i] == 1 for i in range(1, NBIT + 1))
            lines.append("1       0       0       1\n")  # dimension 1: all m_i = 1 (van der Corput)
            continue
        s, a = solve_dim(m)
        lines.append("%-7d %-7d %-7d %s\n" % (d + 1, s, a, " ".join(str(m[i]) for i in range(1, s + 1))))
    with open(out, "w") as fh:
        fh.writelines(lines)
    print("wrote", os.path.abspath(out), len(lines) - 1, "dimensions")


if __name__ == "__main__":
    main()
