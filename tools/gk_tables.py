#!/usr/bin/env python3
"""Gauss-Kronrod (2n+1)-point rule computed from first principles (mpmath, 80 digits): the n Gauss-Legendre nodes,
the n+1 zeros of the Stieltjes polynomial E_{n+1} (orthogonal to all polynomials of degree <= n with weight P_n), and
the weights from the moment equations. Prints the table in the layout artis_b200/csrc/gk31.h uses: non-negative
abscissae in increasing order starting with 0, Kronrod weights, and the weights of the embedded Gauss rule.

  python tools/gk_tables.py [n]     (default n = 15 -> the 31-point rule used by select_continuum_nu)
"""
import sys

import mpmath as mp

mp.mp.dps = 80


def legendre_coeffs(n):
    """coefficients (ascending powers) of P_n"""
    p0, p1 = [mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]
    if n == 0:
        return p0
    for k in range(1, n):
        # (k+1) P_{k+1} = (2k+1) x P_k - k P_{k-1}
        xp = [mp.mpf(0)] + [(2 * k + 1) * c for c in p1]
        pm = [k * c for c in p0] + [mp.mpf(0)] * (len(xp) - len(p0))
        p0, p1 = p1, [(a - b) / (k + 1) for a, b in zip(xp, pm)]
    return p1


def polymul(a, b):
    out = [mp.mpf(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] += x * y
    return out


def integrate_m1_1(c):
    return sum(2 * v / (k + 1) for k, v in enumerate(c) if k % 2 == 0)


def gauss_kronrod(n):
    pn = legendre_coeffs(n)
    m = n + 1  # degree of the Stieltjes polynomial, same parity as n + 1
    powers = list(range(m % 2, m, 2))  # unknown coefficients below the leading (monic) one
    rows, rhs = [], []
    for k in range(m % 2, m, 2):  # moments of the same parity are the non-trivial conditions ... one per unknown
        xk = [mp.mpf(0)] * k + [mp.mpf(1)]
        w = polymul(pn, xk)  # P_n x^k
        # the parity of the product P_n * E_m * x^k must be even for a non-trivial condition
        if (n + m + k) % 2 != 0:
            continue
        rows.append([integrate_m1_1(polymul(w, [mp.mpf(0)] * p + [mp.mpf(1)])) for p in powers])
        rhs.append(-integrate_m1_1(polymul(w, [mp.mpf(0)] * m + [mp.mpf(1)])))
    if len(rows) != len(powers):
        # conditions come from k of the parity that makes the integrand even
        rows, rhs = [], []
        for k in range((n + m) % 2, n + 1, 2):
            xk = [mp.mpf(0)] * k + [mp.mpf(1)]
            w = polymul(pn, xk)
            rows.append([integrate_m1_1(polymul(w, [mp.mpf(0)] * p + [mp.mpf(1)])) for p in powers])
            rhs.append(-integrate_m1_1(polymul(w, [mp.mpf(0)] * m + [mp.mpf(1)])))
        rows, rhs = rows[-len(powers):], rhs[-len(powers):]
    sol = mp.lu_solve(mp.matrix(rows), mp.matrix(rhs))
    e = [mp.mpf(0)] * (m + 1)
    e[m] = mp.mpf(1)
    for p, v in zip(powers, sol):
        e[p] = v
    kron = [mp.re(r) for r in mp.polyroots(list(reversed(e)), maxsteps=2000, extraprec=400)]
    gauss = [mp.re(r) for r in mp.polyroots(list(reversed(pn)), maxsteps=2000, extraprec=400)]
    nodes = sorted(kron + gauss)
    nn = len(nodes)
    a = mp.matrix(nn, nn)
    b = mp.matrix(nn, 1)
    for k in range(nn):
        for i, x in enumerate(nodes):
            a[k, i] = x ** k
        b[k] = mp.mpf(2) / (k + 1) if k % 2 == 0 else mp.mpf(0)
    wk = mp.lu_solve(a, b)
    ag = mp.matrix(n, n)
    bg = mp.matrix(n, 1)
    gs = sorted(gauss)
    for k in range(n):
        for i, x in enumerate(gs):
            ag[k, i] = x ** k
        bg[k] = mp.mpf(2) / (k + 1) if k % 2 == 0 else mp.mpf(0)
    wg = mp.lu_solve(ag, bg)
    nonneg = [(x, w) for x, w in zip(nodes, wk) if x > -mp.mpf(10) ** -40]
    nonneg[0] = (mp.mpf(0), nonneg[0][1])
    gnonneg = [w for x, w in zip(gs, wg) if x > -mp.mpf(10) ** -40]
    return nonneg, gnonneg


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    nodes, gw = gauss_kronrod(n)
    print(f"// {2 * n + 1}-point Gauss-Kronrod rule (tools/gk_tables.py)")
    print("abscissa:")
    for x, _ in nodes:
        print("  " + mp.nstr(x, 25) + ",")
    print("weights:")
    for _, w in nodes:
        print("  " + mp.nstr(w, 25) + ",")
    print("gauss_weights:")
    for w in gw:
        print("  " + mp.nstr(w, 25) + ",")


if __name__ == "__main__":
    main()
