"""Test helper: a slow numpy interpreter of the plan tables with exactly the semantics that
include/e2enet_b200.h documents for e2e_pack_weights / e2e_gather_gemm / e2e_gather_wgrad /
e2e_unpack_wgrad.  It lets the CPU suite validate the host-side plan builder (shift folding,
virtual concat, stride-parity dgrad variants, transposed convs) against the oracle without a GPU.
Test infrastructure only."""
import numpy as np


def to_c8(x):
    """(B,C,D,H,W) fp32 -> (B,Cb,D,H,W,8) fp32 (zero padded)"""
    B, C = x.shape[:2]
    Cb = (C + 7) // 8
    y = np.zeros((B, Cb * 8) + x.shape[2:], np.float64)
    y[:, :C] = x
    return y.reshape((B, Cb, 8) + x.shape[2:]).transpose(0, 1, 3, 4, 5, 2).copy()


def from_c8(y, C):
    B, Cb, D, H, W, _ = y.shape
    return y.transpose(0, 1, 5, 2, 3, 4).reshape(B, Cb * 8, D, H, W)[:, :C].copy()


def pack(plan, w, mask=None):
    wf = w.reshape(-1).astype(np.float64)
    mf = None if mask is None else mask.reshape(-1).astype(np.float64)
    out = np.zeros((plan.n_cent // 2, plan.n_taps, 2, plan.Npad, 8))
    for e in range(plan.n_cent):
        for t in range(plan.n_taps):
            for n in range(plan.Npad):
                ro = plan.rowoff[n]
                if ro < 0:
                    continue
                if plan.emask is not None and not ((int(plan.emask[e]) >> int(plan.rclass[n])) & 1):
                    continue                      # this entry does not feed this column's class
                for j in range(8):
                    co = plan.centoff[e * 8 + j]
                    if co < 0:
                        continue
                    idx = ro + co + plan.tapoff[t]
                    v = wf[idx]
                    if mf is not None:
                        v *= mf[idx]
                    out[e // 2, t, e % 2, n, j] = v
    return out


def _gather(plan, srcs, src_grid, o, e, t):
    """8-vector fetched for iteration voxel o=(b,od,oh,ow), entry e, tap t (zeros if out of range)"""
    b, od, oh, ow = o
    src, blk, dd, dh, dw = plan.cents[e]
    td, th, tw = plan.taps[t]
    d = od * plan.istride[0] + plan.ivoff[0] + dd + td
    h = oh * plan.istride[1] + plan.ivoff[1] + dh + th
    w = ow * plan.istride[2] + plan.ivoff[2] + dw + tw
    D, H, W = src_grid
    if 0 <= d < D and 0 <= h < H and 0 <= w < W:
        return srcs[src][b, blk, d, h, w]
    return np.zeros(8)


def gemm(plan, wp, srcs, src_grid, iter_grid, B, dsts, dst_grid):
    """dsts: list of C8 arrays (written in place where the plan says so)."""
    Do, Ho, Wo = iter_grid
    for b in range(B):
        for od in range(Do):
            for oh in range(Ho):
                for ow in range(Wo):
                    acc = np.zeros(plan.Npad)
                    for e in range(plan.n_cent):
                        for t in range(plan.n_taps):
                            a = _gather(plan, srcs, src_grid, (b, od, oh, ow), e, t)
                            acc += wp[e // 2, t, e % 2] @ a
                    for q in range(plan.Npad // 8):
                        dst, blk, chmask, cd, ch, cw = plan.cols[q]
                        if dst < 0 or chmask == 0:
                            continue
                        d = od * plan.ostride[0] + cd
                        h = oh * plan.ostride[1] + ch
                        w = ow * plan.ostride[2] + cw
                        if not (0 <= d < dst_grid[0] and 0 <= h < dst_grid[1] and 0 <= w < dst_grid[2]):
                            continue
                        for j in range(8):
                            if chmask & (1 << j):
                                dsts[dst][b, blk, d, h, w, j] = acc[q * 8 + j]


def gemm_planar(plan, wp, srcs, src_grid, iter_grid, B, C):
    Do, Ho, Wo = iter_grid
    out = np.zeros((B, C, Do, Ho, Wo))
    for b in range(B):
        for od in range(Do):
            for oh in range(Ho):
                for ow in range(Wo):
                    acc = np.zeros(plan.Npad)
                    for e in range(plan.n_cent):
                        for t in range(plan.n_taps):
                            acc += wp[e // 2, t, e % 2] @ _gather(plan, srcs, src_grid, (b, od, oh, ow), e, t)
                    out[b, :, od, oh, ow] = acc[:C]
    return out


def wgrad(plan, srcs, src_grid, iter_grid, B, grad, weight_shape):
    """grad: C8 on the iteration grid.  Returns the dense weight gradient (reference layout)."""
    Do, Ho, Wo = iter_grid
    gcb = grad.shape[1]
    dwp = np.zeros((plan.n_cent // 2, plan.n_taps, 2, plan.Npad, 8))
    for b in range(B):
        for od in range(Do):
            for oh in range(Ho):
                for ow in range(Wo):
                    gv = np.zeros(plan.Npad)
                    flat = grad[b, :, od, oh, ow, :].reshape(-1)
                    n = min(plan.Npad, gcb * 8)
                    gv[:n] = flat[:n]
                    for e in range(plan.n_cent):
                        for t in range(plan.n_taps):
                            a = _gather(plan, srcs, src_grid, (b, od, oh, ow), e, t)
                            dwp[e // 2, t, e % 2] += np.outer(gv, a)
    gw = np.zeros(int(np.prod(weight_shape)))
    for e in range(plan.n_cent):
        for t in range(plan.n_taps):
            for n in range(plan.Npad):
                ro = plan.rowoff[n]
                if ro < 0:
                    continue
                for j in range(8):
                    co = plan.centoff[e * 8 + j]
                    if co >= 0:
                        gw[ro + co + plan.tapoff[t]] = dwp[e // 2, t, e % 2, n, j]
    return gw.reshape(weight_shape)
