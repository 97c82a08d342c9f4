"""Host-side plan builder for the plan-driven gather GEMMs (include/e2enet_b200.h).

A plan is a handful of small int32 tables that tell the CUDA kernels
  * which 8-channel blocks of which source tensors form the K dimension, and at which
    spatial offset each block is fetched (this is where the reference's depth shift,
    unetpp_d.py:45-59, and its torch.cat, unetpp_d.py:453-478, disappear),
  * which filter taps exist and where each packed weight comes from in the reference's
    fp32 parameter tensor (so state_dict layout stays the reference's),
  * where each 8-column block of the result is stored.
Plans depend only on layer geometry; they are built once and cached on the device.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

SHIFT_SIZE = 5


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def channel_shifts(C: int, shift_size: int = SHIFT_SIZE) -> np.ndarray:
    """shift of every channel of a C-channel (concatenated) conv input: torch.chunk(x, 5, 1)
    makes groups of ceil(C/5); group k is rolled by k - 2 (unetpp_d.py:54-56)."""
    g = -(-C // shift_size)
    return (np.arange(C) // g - shift_size // 2).astype(np.int64)


@dataclass
class GemmPlan:
    cents: np.ndarray            # (n_cent, 5) src, blk, dd, dh, dw
    taps: np.ndarray             # (n_taps, 3)
    cols: np.ndarray             # (Npad/8, 6) dst, blk, chmask, od, oh, ow
    rowoff: np.ndarray           # (Npad,)
    centoff: np.ndarray          # (n_cent*8,)
    tapoff: np.ndarray           # (n_taps,)
    Npad: int
    istride: Tuple[int, int, int] = (1, 1, 1)
    ivoff: Tuple[int, int, int] = (0, 0, 0)
    ostride: Tuple[int, int, int] = (1, 1, 1)
    out_mode: int = 0
    # iteration grid of a data-gradient variant over a destination grid (D, H, W) with conv stride s:
    # it = ceil((D - iter_off) / s) + iter_extra   (see ShiftConvPlan.dgrad_iter_grid)
    iter_off: Tuple[int, int, int] = (0, 0, 0)
    iter_extra: Tuple[int, int, int] = (0, 0, 0)
    halo: bool = False          # stride-1 3x3 halo form (tcgen05 kernel eligible)
    col_bounds: int = 7         # bit 0/1/2: destination depth / row / column of a column block may leave the grid
    emask: Optional[np.ndarray] = None     # (n_cent,) bitmask of column classes an entry feeds (None: all)
    rclass: Optional[np.ndarray] = None    # (Npad,) column class 0..31 (None: 0)
    useful: float = 1.0         # fraction of the (entry, column) weight pairs that are not masked out (FLOP accounting)
    _dev: Dict = field(default_factory=dict, repr=False)

    @property
    def n_cent(self) -> int:
        return int(self.cents.shape[0])

    @property
    def n_taps(self) -> int:
        return int(self.taps.shape[0])

    @property
    def packed_numel(self) -> int:
        return self.n_cent * self.n_taps * self.Npad * 8

    def dev(self, device):
        """device copies of the tables (uploaded once per device)."""
        import torch
        key = str(device)
        if key not in self._dev:
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)
            self._dev[key] = dict(cents=t(self.cents), taps=t(self.taps), cols=t(self.cols), rowoff=t(self.rowoff),
                                  centoff=t(self.centoff), tapoff=t(self.tapoff))
            if self.emask is not None:
                self._dev[key]["emask"] = t(self.emask)
                self._dev[key]["rclass"] = t(self.rclass if self.rclass is not None else np.zeros(self.Npad))
        return self._dev[key]


def _pad_even(cents: List[List[int]], centoff: List[List[int]]):
    if len(cents) % 2:
        cents.append([0, 0, 0, 0, 0])
        centoff.append([-1] * 8)


def _finish(cents, centoff, taps, tapoff, cols, rowoff, emask=None, rclass=None, **kw) -> GemmPlan:
    Npad = ceil_to(max(len(rowoff), 1), 16)
    if emask is not None:
        emask = np.asarray(list(emask) + [0] * (len(cents) - len(emask)), np.int32)      # padded entries feed nothing
        rclass = np.asarray(list(rclass) + [0] * (Npad - len(rclass)), np.int32)
        kw = dict(kw, emask=emask, rclass=rclass)
    rowoff = list(rowoff) + [-1] * (Npad - len(rowoff))
    cols = list(cols)
    while len(cols) < Npad // 8:
        cols.append([-1, 0, 0, 0, 0, 0])
    return GemmPlan(cents=np.asarray(cents, np.int32).reshape(-1, 5), taps=np.asarray(taps, np.int32).reshape(-1, 3),
                    cols=np.asarray(cols, np.int32).reshape(-1, 6), rowoff=np.asarray(rowoff, np.int32),
                    centoff=np.asarray(centoff, np.int32).reshape(-1), tapoff=np.asarray(tapoff, np.int32),
                    Npad=Npad, **kw)


# ----------------------------------------------------------------------------------------
# depth-shifted (1,3,3) conv over a virtual concat of sources
# ----------------------------------------------------------------------------------------
def _col_chunks(n_blocks: int, max_blocks: int = 32) -> List[Tuple[int, int]]:
    """balanced split of n_blocks 8-column blocks into chunks of <= max_blocks (even sizes: Npad % 16 == 0)"""
    n_chunks = -(-n_blocks // max_blocks)
    per = -(-n_blocks // n_chunks)
    per += per % 2
    return [(c0, min(n_blocks, c0 + per)) for c0 in range(0, n_blocks, per)]


@dataclass
class ShiftConvPlan:
    src_channels: List[int]
    cin: int
    cout: int
    stride: Tuple[int, int, int]
    fwd: GemmPlan                # all Cout columns
    fwd_chunks: List[GemmPlan]   # forward GEMMs: column chunks of <= 256
    wgrad: GemmPlan              # gather plan of the weight gradient (fwd, or its point form for tiny Cin)
    fwd3: Optional[GemmPlan]     # narrow stride-1 layers: kw-stacked forward (3 taps kh, N = 3 x Cout), tcgen05 only
    dgrad: List[GemmPlan]        # variants; every element of every source gradient is written at most once
    dgrad_needs_zero: bool       # strided convs: the variants write only the voxels that receive a contribution

    def out_grid(self, D, H, W):
        sd, sh, sw = self.stride
        return (D - 1) // sd + 1, (H + 2 - 3) // sh + 1, (W + 2 - 3) // sw + 1

    @property
    def dgrad_groups(self) -> List[List[GemmPlan]]:
        """variants grouped into GEMMs: the column chunks of one GEMM share the output parity"""
        groups: Dict[Tuple[int, int, int], List[GemmPlan]] = {}
        for v in self.dgrad:
            groups.setdefault(tuple(v.iter_off), []).append(v)
        return list(groups.values())

    def dgrad_iter_grid(self, var: GemmPlan, D, H, W):
        """iteration grid of data-gradient variant `var` for a source grid (D, H, W)."""
        return tuple((n - o + s - 1) // s + e
                     for n, o, s, e in zip((D, H, W), var.iter_off, self.stride, var.iter_extra))


S3_MIN_PAIRS = int(__import__("os").environ.get("E2E_S3_MIN_PAIRS", "1"))


def build_shiftconv_plan(src_channels: Sequence[int], cout: int, stride=(1, 1, 1), shift: bool = True) -> ShiftConvPlan:
    src_channels = [int(c) for c in src_channels]
    cin = sum(src_channels)
    stride = tuple(int(s) for s in stride)
    unit = stride == (1, 1, 1)
    assert cout % 8 == 0, "Cout must be a multiple of 8 (C8 layout)"
    sh_c = channel_shifts(cin) if shift else np.zeros(cin, np.int64)
    src_off = np.concatenate([[0], np.cumsum(src_channels)]).astype(int)
    sd, shh, sww = stride

    # ---- K entries of the forward: (source block, shift)
    cents, centoff = [], []
    for i, ci in enumerate(src_channels):
        for blk in range((ci + 7) // 8):
            chans = [src_off[i] + blk * 8 + j if blk * 8 + j < ci else -1 for j in range(8)]
            for s in sorted({int(sh_c[c]) for c in chans if c >= 0}):
                cents.append([i, blk, -s, 0, 0])                  # x~[d] = x[d - s]
                centoff.append([c * 9 if (c >= 0 and sh_c[c] == s) else -1 for c in chans])
    cols = [[0, q, 0xff, 0, 0, 0] for q in range(cout // 8)]
    rowoff = [n * cin * 9 for n in range(cout)]
    base_cents, base_centoff = [list(c) for c in cents], [list(c) for c in centoff]

    def point_form():
        pc, po = [], []
        for kh in range(3):
            for kw in range(3):
                for ce, co in zip(base_cents, base_centoff):
                    pc.append([ce[0], ce[1], ce[2], kh - 1, kw - 1])
                    po.append([v + kh * 3 + kw if v >= 0 else -1 for v in co])
        _pad_even(pc, po)
        return pc, po

    if unit:
        # halo form: 9 taps over every entry (one haloed window serves all taps)
        _pad_even(cents, centoff)
        taps = [[0, kh - 1, kw - 1] for kh in range(3) for kw in range(3)]
        tapoff = [kh * 3 + kw for kh in range(3) for kw in range(3)]
    else:
        # point form: every (tap, entry) is its own K entry fetched at o*stride + (kh-1, kw-1); conv
        # padding and the shift's zero fill are out-of-range reads
        cents, centoff = point_form()
        taps, tapoff = [[0, 0, 0]], [0]
    mk = lambda cc, rr: _finish([list(c) for c in cents], [list(c) for c in centoff], taps, tapoff, cc, rr,
                                istride=stride, halo=unit, col_bounds=0)
    fwd = mk(cols, rowoff)
    wgrad = fwd
    if unit and 9 * len(base_cents) <= 16:
        # tiny Cin (the network input): the 9 taps of the few entries fill ONE 16-entry group of the
        # weight-gradient kernel as point entries instead of 9 mostly empty tap accumulators
        pc, po = point_form()
        wgrad = _finish(pc, po, [[0, 0, 0]], [0], cols, rowoff, istride=stride)
    chunks = _col_chunks(len(cols))
    fwd_chunks = [fwd] if len(chunks) == 1 else [mk(cols[a:b], rowoff[8 * a:8 * b]) for a, b in chunks]
    fwd3 = None
    npo = ceil_to(cout, 16)
    # (measured: 17 % faster than the 9-tap form on loc4; since the role warps of conv_tc3 became single threads it also
    # wins on the shortest K loop, the 1-channel input layer: E2E_S3_MIN_PAIRS restores the round-1 threshold of 5)
    if unit and npo <= 80 and len(cents) // 2 >= S3_MIN_PAIRS and (len(cents) // 2) * 3 * 2 * 16 * 3 * npo <= 120 * 1024:
        # kw-stacked forward: column kw*Np + n holds W[:, :, kh, kw] of output channel n; the kernel adds the
        # three column groups with a W shift of -1 / 0 / +1 (one A read per three taps)
        row3 = [(n * cin * 9 + kw) if n < cout else -1 for kw in range(3) for n in range(npo)]
        fwd3 = _finish([list(c) for c in cents], [list(c) for c in centoff], [[0, kh - 1, 0] for kh in range(3)],
                       [kh * 3 for kh in range(3)], [list(c) for c in cols], row3, istride=stride, col_bounds=0)
        assert fwd3.Npad == 3 * npo

    # ---- dgrad: source of the GEMM is d(raw) on the conv's output grid, K = Cout x taps.
    # Columns = (source block, shift group); the columns of a group shifted by s are stored at depth
    # o*sd + (smin_or_0 - s):  dx[c, d] = dx~[c, d + s_c], dx~ on depths that are multiples of sd.
    g_cents = [[0, e, 0, 0, 0] for e in range(cout // 8)]
    g_centoff = [[(e * 8 + j) * cin * 9 for j in range(8)] for e in range(cout // 8)]
    shifts = sorted({int(v) for v in sh_c})
    smin, smax = shifts[0], shifts[-1]
    ucols, urow = [], []
    for i, ci in enumerate(src_channels):
        for blk in range((ci + 7) // 8):
            chans = [src_off[i] + blk * 8 + j if blk * 8 + j < ci else -1 for j in range(8)]
            for s in sorted({int(sh_c[c]) for c in chans if c >= 0}):
                m = 0
                for j, c in enumerate(chans):
                    if c >= 0 and sh_c[c] == s:
                        m |= 1 << j
                ucols.append((i, blk, m, s))
                urow.append([c * 9 if (c >= 0 and sh_c[c] == s) else -1 for c in chans])
    variants: List[GemmPlan] = []
    if unit:
        # One halo-form GEMM over the depth range of d(raw) that any shift group needs: iteration
        # depth o reads d(raw) at depth o + smin; depths outside d(raw) read zeros, which also writes
        # the zero slices of dx the shift leaves uncovered.
        gc, go = [list(c) for c in g_cents], [list(c) for c in g_centoff]
        _pad_even(gc, go)
        taps_d = [[0, 1 - kh, 1 - kw] for kh in range(3) for kw in range(3)]
        tapoff_d = [kh * 3 + kw for kh in range(3) for kw in range(3)]
        for a, b in _col_chunks(len(ucols)):
            cc = [[i, blk, m, smin - s, 0, 0] for (i, blk, m, s) in ucols[a:b]]
            rr = [v for r in urow[a:b] for v in r]
            variants.append(_finish([list(c) for c in gc], [list(c) for c in go], taps_d, tapoff_d, cc, rr,
                                    istride=(1, 1, 1), ivoff=(smin, 0, 0), ostride=(1, 1, 1),
                                    iter_extra=(smax - smin, 0, 0), halo=True, col_bounds=1))
        return ShiftConvPlan(src_channels, cin, cout, stride, fwd, fwd_chunks, wgrad, fwd3, variants, False)
    if shh == 2 and sww == 2:
        # stride 2 in H and W: ALL FOUR output parities in ONE point-form GEMM.  dx[2u + p] (p = parity per axis)
        # receives tap k = 1 (p = 0) from d(raw)[u], and taps k = 2 / k = 0 (p = 1) from d(raw)[u] / d(raw)[u + 1].
        # K entries = (offset class per axis: 0 reads u, 1 reads u + 1) x Cout blocks; the weight of (class, parity)
        # is tap k = ke + p with ke = 1 (class 0) or -1 (class 1), and class 1 does not reach parity 0: those
        # (entry, column) pairs are masked to zero at pack time (emask / rclass).  Columns = (ph, pw, source block,
        # shift group) stored at (depth o*sd - s, row 2u + ph, column 2w + pw): the two W parities of a voxel pair are
        # written by the same CTA back to back, so every 32-byte sector is completed by one kernel (four separate
        # parity launches wrote half sectors: measured 3.7x the algorithmic DRAM traffic) and d(raw) is read once.
        vc, vo, em = [], [], []
        for ch_ in (0, 1):
            for cw_ in (0, 1):
                keh, kew = (1 if ch_ == 0 else -1), (1 if cw_ == 0 else -1)
                classes = sum(1 << (ph * 2 + pw) for ph in (0, 1) for pw in (0, 1) if ph >= ch_ and pw >= cw_)
                for ce, co in zip(g_cents, g_centoff):
                    vc.append([0, ce[1], 0, ch_, cw_])
                    vo.append([v + (keh + 1) * 3 + (kew + 1) for v in co])      # kept >= 0: negative = padded channel;
                    em.append(classes)                                          # the -4 lives in tapoff below
        _pad_even(vc, vo)
        pcols, prow, pcls = [], [], []
        for ph in (0, 1):
            for pw in (0, 1):
                for (i, blk, m, s_), r in zip(ucols, urow):
                    pcols.append([i, blk, m, -s_, ph, pw])
                    prow.append([(v + ph * 3 + pw) if v >= 0 else -1 for v in r])
                    pcls.append(ph * 2 + pw)
        for a, b in _col_chunks(len(pcols)):
            rr = [v for r in prow[a:b] for v in r]
            rc = [c for c in pcls[a:b] for _ in range(8)]
            variants.append(_finish([list(c) for c in vc], [list(c) for c in vo], [[0, 0, 0]], [-4], [list(c) for c in pcols[a:b]],
                                    rr, emask=em, rclass=rc, istride=(1, 1, 1), ivoff=(0, 0, 0), ostride=stride,
                                    iter_off=(0, 0, 0), col_bounds=7, useful=9.0 / 16.0))
        # depth stride 2: the odd depth slices receive nothing and stay zero (the caller zero-fills); depth stride 1
        # with the shift still leaves the slices the shift pushes out uncovered -> zero-fill as well
        return ShiftConvPlan(src_channels, cin, cout, stride, fwd, fwd_chunks, wgrad, None, variants, True)
    # strided (general): one point-form GEMM per (H, W) output parity; its K entries are the taps that reach
    # that parity, each fetched from d(raw) at o + (p - k + 1) / stride.  dx is zeroed by the caller
    # (depths / voxels that no tap reaches stay zero).
    for ph in range(shh):
        for pw in range(sww):
            vc, vo = [], []
            for kh in range(3):
                if (ph - kh + 1) % shh:
                    continue
                for kw in range(3):
                    if (pw - kw + 1) % sww:
                        continue
                    for ce, co in zip(g_cents, g_centoff):
                        vc.append([0, ce[1], 0, (ph - kh + 1) // shh, (pw - kw + 1) // sww])
                        vo.append([v + kh * 3 + kw for v in co])
            if not vc:
                continue
            _pad_even(vc, vo)
            for a, b in _col_chunks(len(ucols)):
                cc = [[i, blk, m, -s, ph, pw] for (i, blk, m, s) in ucols[a:b]]
                rr = [v for r in urow[a:b] for v in r]
                variants.append(_finish([list(c) for c in vc], [list(c) for c in vo], [[0, 0, 0]], [0], cc, rr,
                                        istride=(1, 1, 1), ivoff=(0, 0, 0), ostride=stride, iter_off=(0, ph, pw),
                                        col_bounds=1))
    return ShiftConvPlan(src_channels, cin, cout, stride, fwd, fwd_chunks, wgrad, None, variants, True)


# ----------------------------------------------------------------------------------------
# ConvTranspose3d with kernel == stride (the up* modules), weight (Cin, Cout, kd, kh, kw)
# ----------------------------------------------------------------------------------------
@dataclass
class TConvPlan:
    cin: int
    cout: int
    k: Tuple[int, int, int]
    fwd: List[GemmPlan]           # column chunks of <= 256 (N = taps x Cout)
    dgrad: List[GemmPlan]         # column chunks of <= 256 (N = Cin)
    wgrad: GemmPlan               # gather plan of the weight gradient (dgrad's K entries, all Cin columns, grad = x)


def build_tconv_plan(cin: int, cout: int, k) -> TConvPlan:
    k = tuple(int(v) for v in k)
    kd, kh, kw = k
    kv = kd * kh * kw
    assert cin % 8 == 0 and cout % 8 == 0
    tl = [(a, b, c) for a in range(kd) for b in range(kh) for c in range(kw)]
    # forward: K = Cin, N = (tap, Cout)
    cents = [[0, e, 0, 0, 0] for e in range(cin // 8)]
    centoff = [[(e * 8 + j) * cout * kv for j in range(8)] for e in range(cin // 8)]
    _pad_even(cents, centoff)
    # column order (a, b) -> channel block -> c: the kw column blocks that land in adjacent 16-byte
    # halves of the same 32-byte sectors sit next to each other and are stored back to back
    cols, rowoff = [], []
    for a in range(kd):
        for b in range(kh):
            for q in range(cout // 8):
                for c in range(kw):
                    t = (a * kh + b) * kw + c
                    cols.append([0, q, 0xff, a, b, c])
                    rowoff.extend([(q * 8 + j) * kv + t for j in range(8)])
    fwd = [_finish([list(c) for c in cents], [list(c) for c in centoff], [[0, 0, 0]], [0], cols[a:b],
                   rowoff[8 * a:8 * b], ostride=k, col_bounds=0) for a, b in _col_chunks(len(cols))]
    # dgrad: K = (tap, Cout block) gathered from dy at u*k + tap, N = Cin
    cents, centoff = [], []
    for t, (a, b, c) in enumerate(tl):
        for q in range(cout // 8):
            cents.append([0, q, a, b, c])
            centoff.append([(q * 8 + j) * kv + t for j in range(8)])
    _pad_even(cents, centoff)
    cols = [[0, q, 0xff, 0, 0, 0] for q in range(cin // 8)]
    rowoff = [n * cout * kv for n in range(cin)]
    wgrad = _finish([list(c) for c in cents], [list(c) for c in centoff], [[0, 0, 0]], [0], cols, rowoff, istride=k)
    dgrad = [_finish([list(c) for c in cents], [list(c) for c in centoff], [[0, 0, 0]], [0], cols[a:b],
                     rowoff[8 * a:8 * b], istride=k, col_bounds=0) for a, b in _col_chunks(len(cols))]
    return TConvPlan(cin, cout, k, fwd, dgrad, wgrad)


# ----------------------------------------------------------------------------------------
# 1x1x1 seg head, weight (ncls, C, 1, 1, 1), logits returned as fp32 NCDHW
# ----------------------------------------------------------------------------------------
@dataclass
class SegHeadPlan:
    cin: int
    ncls: int
    fwd: GemmPlan
    dgrad: GemmPlan


def build_seghead_plan(cin: int, ncls: int) -> SegHeadPlan:
    assert cin % 8 == 0
    cents = [[0, e, 0, 0, 0] for e in range(cin // 8)]
    centoff = [[e * 8 + j for j in range(8)] for e in range(cin // 8)]
    _pad_even(cents, centoff)
    rowoff = [n * cin for n in range(ncls)]
    fwd = _finish(cents, centoff, [[0, 0, 0]], [0], [], rowoff, out_mode=1)
    ncb = (ncls + 7) // 8
    cents = [[0, e, 0, 0, 0] for e in range(ncb)]
    centoff = [[(e * 8 + j) * cin if e * 8 + j < ncls else -1 for j in range(8)] for e in range(ncb)]
    _pad_even(cents, centoff)
    cols = [[0, q, 0xff, 0, 0, 0] for q in range(cin // 8)]
    dgrad = _finish(cents, centoff, [[0, 0, 0]], [0], cols, list(range(cin)), col_bounds=0)
    return SegHeadPlan(cin, ncls, fwd, dgrad)
