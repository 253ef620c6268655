"""A numpy model of how pvo_b200/csrc/ba.cu solves the reduced camera system of a large window (ba_solve_kernel, packed branch), checked
against a dense solve: the band-packed storage ba_combine_band_kernel writes, the T | S | B split of the two-sided factorisation, the
index-reversed copy the second half of the block works on, the merge of the two Schur complements into the separator and the push-style
back substitution restricted to a column range.  It pins the index arithmetic of the kernel in a form that runs without a GPU; the kernel
itself is tested against the oracle in tests/test_geom_gpu.py (F = 34 / 64: two-sided; F = 50, and F = 64 at radius 8: windowed)."""
import numpy as np
import pytest


def pack(A, band):
    """[n][band + 6]: row r keeps columns max(0, 6*(r//6) - band) .. that + band + 5, lower triangle only (ba_combine_band_kernel)"""
    n, W = A.shape[0], band + 6
    P = np.zeros((n, W))
    for r in range(n):
        c0 = max(0, 6 * (r // 6) - band)
        for cc in range(W):
            c = c0 + cc
            if c <= r:
                P[r, cc] = A[r, c]
    return P


class Band:
    """SysMat<true>: element (r, c) of a band-packed lower triangle with pitch ld"""
    def __init__(self, n, band, ld):
        self.n, self.band, self.ld = n, band, ld
        self.p = np.zeros(n * ld)

    def off(self, r):
        return r * self.ld - max(0, 6 * (r // 6) - self.band)

    def get(self, r, c):
        return self.p[self.off(r) + c]

    def set(self, r, c, v):
        self.p[self.off(r) + c] = v


def forward(M, y, dinv, n, band, kb0, kb1, factor_first=True, factor_last=True):
    """ldl_forward, arithmetic only (right-looking by 6x6 block columns; the forward substitution rides along)"""
    def diag(c0):
        for j in range(6):
            d = M.get(c0 + j, c0 + j)
            assert d > 0
            dinv[c0 + j] = 1.0 / d
            v = {i: M.get(c0 + i, c0 + j) for i in range(j + 1, 6)}      # column j, still L_ij d_j
            for i in range(j + 1, 6):
                l = v[i] / d
                y[c0 + i] -= l * y[c0 + j]
                for c in range(j + 1, i + 1):
                    M.set(c0 + i, c0 + c, M.get(c0 + i, c0 + c) - l * v[c])
                M.set(c0 + i, c0 + j, l)
    if factor_first:
        diag(6 * kb0)
    for kb in range(kb0, kb1):
        c0 = 6 * kb
        rend = min(n, c0 + 6 + band)
        LD = {}
        for r in range(c0 + 6, rend):                       # panel: L_rk and L_rk d
            x = np.zeros(6)
            v = np.zeros(6)
            for j in range(6):
                v[j] = M.get(r, c0 + j) - sum(x[t] * M.get(c0 + j, c0 + t) * M.get(c0 + t, c0 + t) for t in range(j))
                x[j] = v[j] * dinv[c0 + j]
                y[r] -= x[j] * y[c0 + j]
            for j in range(6):
                M.set(r, c0 + j, x[j])
            LD[r] = (x, v)
        for i in range(c0 + 6, rend):                       # trailing update, lower triangle
            for c in range(c0 + 6, i + 1):
                M.set(i, c, M.get(i, c) - float(np.dot(LD[i][1], LD[c][0])))
        if rend > c0 + 6 and (factor_last or kb + 1 < kb1):
            diag(c0 + 6)


def backward(M, y, n, band, kb_hi, kb_lo):
    """ldl_backward: rows below the range are pushed first; a column outside [kb_lo, kb_hi] is never written"""
    clo = 6 * kb_lo

    def push(rb, cb):
        for c in range(max(clo, rb - band), cb):
            y[c] -= sum(M.get(rb + j, c) * y[rb + j] for j in range(6))
    ctop = 6 * (kb_hi + 1)
    for rb in range(ctop, min(n, ctop + band), 6):
        push(rb, ctop)
    for kb in range(kb_hi, kb_lo - 1, -1):
        c0 = 6 * kb
        for j in range(5, -1, -1):
            y[c0 + j] -= sum(M.get(c0 + t, c0 + j) * y[c0 + t] for t in range(j + 1, 6))
        push(c0, c0)


def two_sided_solve(packed, b, P, bwb):
    """ba_solve_kernel's two-sided branch on the packed block; returns x"""
    n, band = 6 * P, 6 * bwb
    W, Ws = band + 6, band + 7
    Tb = (P - bwb + 1) // 2
    Bb = P - bwb - Tb
    assert Bb >= 2
    n1, nB = 6 * (Tb + bwb), 6 * Bb
    n2 = nB + band
    M1, M2 = Band(n1, band, Ws), Band(n2, band, Ws)
    for r in range(n1):                                     # matrix 1 = the first n1 packed rows
        for cc in range(W):
            M1.p[r * Ws + cc] = packed[r, cc]
    for R in range(6 * Tb + band, n):                       # matrix 2: element (R, C) of the rows of B lands at (n-1-C, n-1-R)
        c0 = max(0, 6 * (R // 6) - band)
        for cc in range(W):
            C = c0 + cc
            if C > R:
                continue
            r, c = n - 1 - C, n - 1 - R
            M2.p[r * Ws + c - max(0, 6 * (r // 6) - band)] = packed[R, cc]
    y1 = b[:n1].copy()
    y2 = np.array([b[n - 1 - l] if l < nB else 0.0 for l in range(n2)])
    d1, d2 = np.zeros(n1), np.zeros(n2)
    forward(M1, y1, d1, n1, band, 0, Tb, True, False)       # the two halves of the block, side by side on the GPU
    forward(M2, y2, d2, n2, band, 0, Bb, True, False)
    for a in range(band):                                   # S += B's Schur complement
        for bb in range(a + 1):
            M1.set(6 * Tb + a, 6 * Tb + bb, M1.get(6 * Tb + a, 6 * Tb + bb) + M2.get(n - 1 - (6 * Tb + bb), n - 1 - (6 * Tb + a)))
        y1[6 * Tb + a] += y2[n - 1 - (6 * Tb + a)]
    forward(M1, y1, d1, n1, band, Tb, Tb + bwb, True, True)
    y1 *= d1
    y2[:nB] *= d2[:nB]
    backward(M1, y1, n1, band, Tb + bwb - 1, Tb)            # x_S
    for l in range(nB, n2):
        y2[l] = y1[n - 1 - l]
    backward(M1, y1, n1, band, Tb - 1, 0)
    backward(M2, y2, n2, band, Bb - 1, 0)
    return np.array([y1[i] if i < n1 else y2[n - 1 - i] for i in range(n)])


def banded_spd(P, bwb, seed):
    g = np.random.default_rng(seed)
    n = 6 * P
    A = np.zeros((n, n))
    for i in range(P):
        for j in range(max(0, i - bwb), i + 1):
            blk = g.standard_normal((6, 6))
            A[6 * i:6 * i + 6, 6 * j:6 * j + 6] = blk
            A[6 * j:6 * j + 6, 6 * i:6 * i + 6] = blk.T
    A = 0.5 * (A + A.T)
    A += np.eye(n) * (np.abs(A).sum(1).max() + 1.0)         # diagonally dominant => SPD
    return A, g.standard_normal(n)


@pytest.mark.parametrize('P,bwb', [(63, 6), (31, 6), (32, 2), (12, 2), (20, 4)])
def test_two_sided_band_solve_model(P, bwb):
    A, b = banded_spd(P, bwb, seed=P * 10 + bwb)
    x = two_sided_solve(pack(A, 6 * bwb), b, P, bwb)
    want = np.linalg.solve(A, b)
    assert np.abs(x - want).max() / np.abs(want).max() < 1e-11


@pytest.mark.parametrize('P,bwb,kcount', [(20, 3, 5), (17, 2, 4)])
def test_windowed_band_solve_model(P, bwb, kcount):
    """the windowed schedule: forward over block-column ranges that continue each other (factor_first only in the first), backward over the
    same ranges in reverse, each pushing the rows below its range first"""
    A, b = banded_spd(P, bwb, seed=7 * P + bwb)
    n, band = 6 * P, 6 * bwb
    M = Band(n, band, band + 7)
    pk = pack(A, band)
    for r in range(n):
        M.p[r * (band + 7):r * (band + 7) + band + 6] = pk[r]
    y, dinv = b.copy(), np.zeros(n)
    starts = list(range(0, P, kcount))
    for k0 in starts:
        forward(M, y, dinv, n, band, k0, min(P, k0 + kcount), factor_first=(k0 == 0), factor_last=True)
    y *= dinv
    for k0 in reversed(starts):
        backward(M, y, n, band, min(P, k0 + kcount) - 1, k0)
    want = np.linalg.solve(A, b)
    assert np.abs(y - want).max() / np.abs(want).max() < 1e-11
