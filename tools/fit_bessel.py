#!/usr/bin/env python3
"""Fit the polynomial kernels used by the device Bessel evaluation and write
qbold_vi_b200/csrc/bessel_coef.h.

Device algorithm (qbold_vi_b200/csrc/bessel.cuh), three ranges:
  small x <= X1      : 1 - J0(x) = z * S0(z),  J1(x) = x * S1(z),  z = x^2
                       (no cancellation: the quadrature needs 1 - J0, not J0)
  mid   X1 < x <= X2 : 1 - J0 and J1 as polynomials in t = x - XC; valid on [MID_LO, X2]
  big   x > X2       : valid from BIG_LO;  J_n(x) = rsqrt(x) * A_n(w) * cos(x - (2n+1)pi/4 + q * F_n(w)),
                       q = 1/x, w = q^2  (modulus / phase form; A_n, F_n polynomials in w)
  cos on [-pi/2, pi/2] after a 2-constant Cody-Waite reduction mod pi: polynomial in r^2.

These are our own near-minimax fits (Chebyshev interpolation in float64, rounded to
float32), not the Cephes coefficients the oracle uses; the script reports the float32
(FMA-emulated) error of every piece against scipy/mpmath-grade float64 values.
"""
import os
import sys

import numpy as np
import scipy.special as sp
from numpy.polynomial import chebyshev as C, polynomial as P

X1, X2 = 3.0, 9.0            # per-lane split points
MID_LO, BIG_LO = 2.0, 6.5    # the mid / big kernels stay valid down to here, so a pass (or phase) that
                             # straddles a split point can still run ONE kernel on all lanes
SMALL_FIT_HI = 3.1           # fit a little beyond the split
MID_FIT = (MID_LO - 0.05, X2 + 0.05)
XC = 0.5 * (MID_FIT[0] + MID_FIT[1])
BIG_FIT_LO = BIG_LO - 0.1
DEG_SMALL = 5
DEG_MID = 12
DEG_AMP = 2
DEG_PHASE = 2
DEG_COS = 4                  # in r^2, |r| <= pi/2 + 0.02


def cheb_fit(f, lo, hi, deg, n=6000):
    k = np.arange(n)
    t = np.cos(np.pi * (k + 0.5) / n)
    x = 0.5 * (hi - lo) * t + 0.5 * (hi + lo)
    c = C.chebfit(t, f(x), deg)
    p = C.cheb2poly(c)
    a, b = 2 / (hi - lo), -(hi + lo) / (hi - lo)
    return P.Polynomial(p)(P.Polynomial([b, a])).coef      # ascending powers


def fma32(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(np.float32) if np.isscalar(b) else \
        (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, dtype=np.float64)).astype(np.float32)


def horner32(coef, x):
    """Horner with float32 FMA steps; coef ascending."""
    x = x.astype(np.float32)
    acc = np.full_like(x, np.float32(coef[-1]))
    for c in coef[-2::-1]:
        acc = (acc.astype(np.float64) * x.astype(np.float64) + np.float64(np.float32(c))).astype(np.float32)
    return acc


def amp_phase(order, x):
    J, Y = sp.jv(order, x), sp.yv(order, x)
    M = np.sqrt(J * J + Y * Y)
    th = np.arctan2(Y, J)
    base = x - (0.25 + 0.5 * order) * np.pi
    phi = (th - base + np.pi) % (2 * np.pi) - np.pi
    return M, phi


def flit(v):
    """float32 value -> C float literal that round-trips."""
    t = '%.9g' % np.float32(v)
    if not any(ch in t for ch in '.e'):
        t += '.0'
    return t + 'f'


def main():
    out = {}
    # ---- small
    f0 = lambda z: np.where(z > 1e-10, (1 - sp.j0(np.sqrt(np.maximum(z, 0)))) / np.maximum(z, 1e-300), 0.25 - z / 64)
    f1 = lambda z: np.where(z > 1e-10, sp.j1(np.sqrt(np.maximum(z, 0))) / np.sqrt(np.maximum(z, 1e-300)), 0.5 - z / 16)
    out['S0'] = cheb_fit(f0, 0.0, SMALL_FIT_HI ** 2, DEG_SMALL)
    out['S1'] = cheb_fit(f1, 0.0, SMALL_FIT_HI ** 2, DEG_SMALL)
    # ---- mid
    out['M0'] = cheb_fit(lambda t: 1 - sp.j0(t + XC), MID_FIT[0] - XC, MID_FIT[1] - XC, DEG_MID)
    out['M1'] = cheb_fit(lambda t: sp.j1(t + XC), MID_FIT[0] - XC, MID_FIT[1] - XC, DEG_MID)
    # ---- big
    wmax = 1.0 / BIG_FIT_LO ** 2
    for order in (0, 1):
        fa = lambda w, o=order: (lambda x: amp_phase(o, x)[0] * np.sqrt(x))(1 / np.sqrt(w))
        fp = lambda w, o=order: (lambda x: amp_phase(o, x)[1] * x)(1 / np.sqrt(w))
        out['A%d' % order] = cheb_fit(fa, 1e-7, wmax, DEG_AMP)
        out['F%d' % order] = cheb_fit(fp, 1e-7, wmax, DEG_PHASE)
    # ---- cos(r), |r| <= pi/2 + margin, polynomial in s = r^2
    rmax = np.pi / 2 + 0.02
    fc = lambda s: np.cos(np.sqrt(np.maximum(s, 0)))
    out['CS'] = cheb_fit(fc, 0.0, rmax ** 2, DEG_COS)

    # ---- report float32 accuracy of the composed evaluation
    x = np.linspace(1e-4, X1, 300001).astype(np.float32)
    xd = x.astype(np.float64)
    z = (x * x).astype(np.float32)
    v0 = z * horner32(out['S0'], z)
    v1 = x * horner32(out['S1'], z)
    e0 = np.abs(v0 - (1 - sp.j0(xd)))
    print('small: 1-J0 abs %.2e rel %.2e | J1 abs %.2e' % (e0.max(), (e0 / (1 - sp.j0(xd))).max(),
                                                         np.abs(v1 - sp.j1(xd)).max()))
    x = np.linspace(MID_LO, X2, 300001).astype(np.float32)
    xd = x.astype(np.float64)
    t = (x - np.float32(XC)).astype(np.float32)
    print('mid  : 1-J0 abs %.2e | J1 abs %.2e' % (np.abs(horner32(out['M0'], t) - (1 - sp.j0(xd))).max(),
                                                 np.abs(horner32(out['M1'], t) - sp.j1(xd)).max()))
    x = np.linspace(BIG_LO, 40, 600001).astype(np.float32)
    xd = x.astype(np.float64)
    r = (1 / np.sqrt(xd)).astype(np.float32)
    q = (r * r).astype(np.float32)
    w = (q * q).astype(np.float32)
    INV_PI = np.float32(1 / np.pi)
    PI_HI = np.float32(3.140625)
    PI_LO = np.float32(np.pi - 3.140625)
    for order in (0, 1):
        A = r * horner32(out['A%d' % order], w)
        F = horner32(out['F%d' % order], w)
        y = (x + np.float32(-(0.25 + 0.5 * order) * np.pi)).astype(np.float32)
        th = fma32(q, F, y)
        n = np.rint(th.astype(np.float64) * INV_PI).astype(np.float32)
        rr = fma32(n, -PI_HI, th)
        rr = fma32(n, -PI_LO, rr)
        s = (rr * rr).astype(np.float32)
        c = horner32(out['CS'], s)
        val = A * c * np.where(n.astype(np.int64) % 2 == 0, 1, -1).astype(np.float32)
        print('big  : J%d abs %.2e   (|rr| max %.4f)' % (order, np.abs(val - sp.jv(order, xd)).max(), np.abs(rr).max()))

    # ---- emit header
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        'qbold_vi_b200', 'csrc', 'bessel_coef.h')
    with open(path, 'w') as f:
        f.write('// GENERATED by tools/fit_bessel.py -- do not edit.\n')
        f.write('// Near-minimax float32 polynomial kernels for 1-J0 / J1 (see bessel.cuh).\n#pragma once\n')
        f.write('namespace qb {\nnamespace coef {\n')
        f.write('constexpr float kX1 = %s, kX2 = %s, kXC = %s, kMidLo = %s, kBigLo = %s;\n'
                % (flit(X1), flit(X2), flit(XC), flit(MID_LO), flit(BIG_LO)))
        for name, coef in out.items():
            f.write('// %s: ascending powers, degree %d\n' % (name, len(coef) - 1))
            f.write('struct %s {\n    static constexpr int N = %d;\n    __host__ __device__ static constexpr float c(int i) {\n'
                    '        constexpr float v[%d] = {%s};\n        return v[i];\n    }\n};\n'
                    % (name, len(coef), len(coef), ', '.join(flit(c) for c in coef)))
        f.write('}  // namespace coef\n}  // namespace qb\n')
    print('wrote', path)


if __name__ == '__main__':
    sys.exit(main())
