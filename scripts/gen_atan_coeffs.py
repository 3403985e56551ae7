"""Generate the polynomial coefficients of the device atan / asin kernels (csrc/r2ik_math.cuh).

g(q) = q * P(q^2) on q in [0, qmax] for g = atan (default) or asin: P is the Chebyshev interpolant
(near-minimax) of f(s) = g(sqrt(s)) / sqrt(s) on s in [0, qmax^2], computed with mpmath at 60
digits and rounded to double.  Prints the monomial coefficients (c0 first) and the max relative
error of the double-rounded polynomial evaluated in exact arithmetic.

    python scripts/gen_atan_coeffs.py [qmax] [n_terms] [atan|asin]
    python scripts/gen_atan_coeffs.py 0.41421356237309515 12 atan    # tan(pi/8)
    python scripts/gen_atan_coeffs.py 0.3826834323650898 12 asin     # sin(pi/8)
"""
import sys

import mpmath as mp

mp.mp.dps = 60


G = mp.asin if (len(sys.argv) > 3 and sys.argv[3] == "asin") else mp.atan


def f(s):
    if s == 0:
        return mp.mpf(1)
    r = mp.sqrt(s)
    return G(r) / r


def fit(qmax, n):
    smax = mp.mpf(qmax) ** 2
    # Chebyshev nodes on [0, smax]
    xs = [smax * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * n))) / 2 for k in range(n)]
    # solve the Vandermonde system at high precision -> monomial coefficients
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        b[i] = f(x)
    c = mp.lu_solve(A, b)
    cd = [float(c[j]) for j in range(n)]
    # error of the rounded polynomial
    worst = mp.mpf(0)
    for k in range(2001):
        s = smax * k / 2000
        p = mp.mpf(0)
        for j in reversed(range(n)):
            p = p * s + mp.mpf(cd[j])
        worst = max(worst, abs(p / f(s) - 1))
    return cd, worst


if __name__ == "__main__":
    qmax = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    if len(sys.argv) > 2:
        cd, w = fit(qmax, int(sys.argv[2]))
        print(f"// qmax={qmax} terms={len(cd)} max rel err {mp.nstr(w, 3)}")
        for c in cd:
            print(f"  {c!r},  // {c.hex()}")
    else:
        for n in range(8, 26):
            cd, w = fit(qmax, n)
            print(n, mp.nstr(w, 3))
