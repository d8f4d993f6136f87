"""Host-side planning for the tap-GEMM engine: tile boxes, K schedules and weight-packing schedules.

Pure Python (no CUDA), so the CPU test-suite can check every schedule against ``torch.nn.functional.conv2d``.

The K loop of a convolution is a list of *load groups*.  A group (src, dw, dh0, c0) is ONE TMA box: 64 input
channels starting at ``c0`` of view ``src``, read at the output tile shifted by (dh0, dw) and ``th + R - 1`` rows
tall.  It feeds ``R`` vertically adjacent filter taps (dh = dh0 .. dh0+R-1): tap r simply starts ``r*tw`` rows
further down the same shared-memory box, so the rows a 3x3 window shares vertically are fetched once instead of
three times.  ``R`` is uniform per layer (3 for 3x3, 2 for the stride-2 and the nearest-2x convs, 1 for 1x1 or
whenever a tile spans several samples).  The packed weight has one 64-wide K block per (group, r), in that order;
the packing entry (cin0, nvalid, tapmask) says which slice of the fp32 ``[Cout][Cin_total][taps]`` weight lands
there (``tapmask`` sums several filter taps: nearest-2x fold).

Reference call sites: CCDM_unified/models/unet.py:74-81 (Upsample / Downsample), :139 (3x3), :165,195,198 (1x1).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

KB = 64          # channels per K block
TILE_M = 128     # output positions per tile


def _pow2_floor_div(n: int, cap: int) -> int:
    """Largest power of two <= cap that divides n (1 if n is odd)."""
    p = 1
    while p * 2 <= cap and n % (p * 2) == 0:
        p *= 2
    return p


def _pow2_ceil(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def tile_box(gw: int, gh: int, force_tb1: bool = False, square: bool = False) -> Tuple[int, int, int]:
    """(tw, th, tb) with tw*th*tb == 128 for an output grid of gw x gh positions per sample.

    Prefers boxes that divide the grid exactly (no masked rows); falls back to the next power of two for odd
    sizes (3x3, 6x6 bottlenecks of the 192-px model).  ``square`` asks for a 16x8 box when the grid allows it:
    with vertical tap reuse the halo overhead is (th+2)/th, so tall boxes fetch less.  ``force_tb1`` keeps a tile
    inside one sample (per-sample weights) at the price of masked rows on tiny grids.
    """
    cap_w = 16 if (square and gw % 16 == 0 and gh % 8 == 0) else 64
    tw = _pow2_floor_div(gw, cap_w)
    if tw < 4 and tw < gw:
        tw = min(_pow2_ceil(gw), 64)
    cap_h = TILE_M // tw
    th = _pow2_floor_div(gh, cap_h)
    if th < min(4, cap_h) and th < gh:
        th = min(_pow2_ceil(gh), cap_h)
    if force_tb1:
        th = cap_h
    tb = TILE_M // (tw * th)
    assert tw * th * tb == TILE_M
    return tw, th, tb


def can_reuse_rows(tile: Tuple[int, int, int]) -> bool:
    """Vertical tap reuse needs one sample per tile and 8-row-aligned tap offsets inside the shared-memory box."""
    tw, th, tb = tile
    return tb == 1 and tw % 8 == 0


@dataclass
class ConvPlan:
    """Everything ccdm_tapgemm / ccdm_pack_weights need for one layer, except pointers.

    For the data-gradient plans (``*_dgrad``) the roles swap: the GEMM's input channels are the forward layer's
    OUTPUT channels (``cins`` = (Cout,)), its output channels the forward layer's input channels, and packing entries
    index the forward weight as W[k_channel][n_channel][tap] (``transposed``)."""
    kind: str
    cins: Tuple[int, ...]                 # channels of each concatenated source
    cout: int
    ntaps: int                            # kh*kw of the fp32 weight
    nz: int
    ngroups: int                          # load groups per sub-problem
    R: int                                # vertical taps served by one load group
    sched: List[Tuple[int, int, int, int]] = field(default_factory=list)    # [nz*ngroups] (src, dw, dh0, c0)
    psched: List[Tuple[int, int, int, int]] = field(default_factory=list)   # [nz*nkb] (cin0, nvalid, tapmask, 0)
    n_views: int = 1                      # views per source tensor (4 parity planes for the stride-2 conv)
    out_parity: bool = False              # nz == 4 output parity planes (nearest-2x + 3x3)
    stride: int = 1
    transposed: bool = False              # data-gradient plan: pack W with input/output channel roles swapped
    halo: bool = False                    # 3x3: ONE box {64, tw+2, th+2} per 64-channel block feeds all nine taps (R == 9;
                                          # K block g*9 + r*3 + q = tap (r, q)); tile 8 x 16 x 1

    @property
    def nkb(self) -> int:
        return self.ngroups * self.R

    @property
    def ksteps(self):
        """[nz*ngroups] K steps (16 channels) of each load group's 64-channel block that hold data, or None when every block is
        full (or the plan has no packing schedule / uses halo boxes): ``ccdm_tapgemm_args.ksteps``.  Cached on the plan."""
        tab = self.__dict__.get("_ksteps", 0)
        if tab == 0:
            tab = None
            if self.psched and not self.halo and len(self.psched) == self.nz * self.nkb:
                t = [[max(1, -(-self.psched[g * self.R][1] // 16))] for g in range(self.nz * self.ngroups)]
                if any(v[0] != 4 for v in t):
                    tab = t
            self.__dict__["_ksteps"] = tab
        return tab

    @property
    def n_src(self) -> int:
        return len(self.cins) * self.n_views


def _chunks(c: int):
    return [(c0, min(KB, c - c0)) for c0 in range(0, c, KB)]


HALO_TILE = (8, 16, 1)


def halo_ok(kind: str, gw: int, gh: int) -> bool:
    """A 3x3 stride-1 layer whose output grid is covered by 8 x 16 tiles can read its taps out of one halo box per tile."""
    return kind == "3x3" and gw % HALO_TILE[0] == 0 and gh % HALO_TILE[1] == 0


def plan_conv(kind: str, cins: Sequence[int], cout: int, reuse_rows: bool = False, halo: bool = False) -> ConvPlan:
    """kind: '1x1' | '3x3' | 'down4x4s2' | 'down3x3s2' | 'up2x3x3' | 'up2x1x1' | 'stem7' | '<kind>_dgrad'.  ``reuse_rows`` groups vertically adjacent taps (R > 1)."""
    cins = tuple(int(c) for c in cins)
    offs = [sum(cins[:i]) for i in range(len(cins))]
    sched, psched = [], []

    def emit(src_view, s, dw, rows):
        """rows: [(dh, tapmask)] vertically consecutive taps of one column offset dw for source s."""
        for c0, nv in _chunks(cins[s]):
            if reuse_rows:
                assert all(rows[i + 1][0] == rows[i][0] + 1 for i in range(len(rows) - 1))
                sched.append((src_view, dw, rows[0][0], c0))
                for _, mask in rows:
                    psched.append((offs[s] + c0, nv, mask, 0))
            else:
                for dh, mask in rows:
                    sched.append((src_view, dw, dh, c0))
                    psched.append((offs[s] + c0, nv, mask, 0))

    if kind == "stem7":
        # 7x7 stem over the 64-channel im2row tensor (ccdm_stem_im2row, H+1 rows): vertical tap pair g sits at row
        # offset 2g-2.  Weights come from ccdm_stem_pack, so there is no packing schedule.
        sched = [(0, 0, 2 * g - 2, 0) for g in range(4)]
        return ConvPlan(kind, cins, cout, 49, 1, 4, 1, sched, [])
    if kind == "1x1":
        for s in range(len(cins)):
            emit(s, s, 0, [(0, 1)])
        R = 1
        return ConvPlan(kind, cins, cout, 1, 1, len(sched), 1, sched, psched)
    if kind == "3x3" and halo:
        for s in range(len(cins)):
            for c0, nv in _chunks(cins[s]):
                sched.append((s, -1, -1, c0))
                for t in range(9):
                    psched.append((offs[s] + c0, nv, 1 << t, 0))          # tap (r, q) = divmod(t, 3)
        return ConvPlan(kind, cins, cout, 9, 1, len(sched), 9, sched, psched, halo=True)
    if kind == "3x3":
        for q in range(3):
            for s in range(len(cins)):
                emit(s, s, q - 1, [(r - 1, 1 << (r * 3 + q)) for r in range(3)])
        R = 3 if reuse_rows else 1
        return ConvPlan(kind, cins, cout, 9, 1, len(sched), R, sched, psched)
    if kind == "down4x4s2":
        # out[ho,wo] = sum_{r,q} W[r,q] x[2ho-1+r, 2wo-1+q];  2ho-1+r = 2(ho+dr)+pr  with (dr,pr) below.
        # The 4 parity planes x[pr::2, pq::2] are separate strided TMA views: view index = src*4 + pr*2 + pq.
        taps = {1: [(-1, 0), (0, 2)], 0: [(0, 1), (1, 3)]}       # parity -> [(shift, filter index)]
        for pr in (1, 0):
            for pq in (1, 0):
                for dq, q in taps[pq]:
                    for s in range(len(cins)):
                        emit(s * 4 + pr * 2 + pq, s, dq, [(dr, 1 << (r * 4 + q)) for dr, r in taps[pr]])
        R = 2 if reuse_rows else 1
        return ConvPlan(kind, cins, cout, 16, 1, len(sched), R, sched, psched, n_views=4, stride=2)
    if kind == "down3x3s2":
        # the vanilla UNet's Downsample (CCDM_vanilla/.../models/unet.py:193-198): 3x3, stride 2, padding 1.
        # out[ho,wo] = sum_{r,q} W[r,q] x[2ho-1+r, 2wo-1+q]; row r lives on parity plane (r+1)%2 at shift -1 (r = 0) or 0.
        # With vertical reuse R must be uniform, so the even plane's single row is paired with a zero-weight row.
        taps = {1: [(-1, 0), (0, 2)], 0: [(0, 1)]}               # parity -> [(shift, filter index)]
        for pr in (1, 0):
            for pq in (1, 0):
                for dq, q in taps[pq]:
                    rows = [(dr, 1 << (r * 3 + q)) for dr, r in taps[pr]]
                    if reuse_rows and len(rows) == 1:
                        rows.append((rows[0][0] + 1, 0))
                    for s in range(len(cins)):
                        emit(s * 4 + pr * 2 + pq, s, dq, rows)
        R = 2 if reuse_rows else 1
        return ConvPlan(kind, cins, cout, 9, 1, len(sched), R, sched, psched, n_views=4, stride=2)
    if kind == "up2x3x3":
        # nearest-2x then 3x3/p1: output (2a+pa, 2b+pb) sees a 2x2 window of the low-res input; taps that hit the
        # same low-res pixel are summed at packing time (9 -> 4 taps, 2.25x fewer MACs).
        win = {0: [(-1, (0,)), (0, (1, 2))], 1: [(0, (0, 1)), (1, (2,))]}
        for pa in range(2):
            for pb in range(2):
                for dw, qset in win[pb]:
                    rows = []
                    for dh, rset in win[pa]:
                        mask = 0
                        for r in rset:
                            for q in qset:
                                mask |= 1 << (r * 3 + q)
                        rows.append((dh, mask))
                    for s in range(len(cins)):
                        emit(s, s, dw, rows)
        R = 2 if reuse_rows else 1
        return ConvPlan(kind, cins, cout, 9, 4, len(sched) // 4, R, sched, psched, out_parity=True)
    if kind == "up2x1x1":
        # nearest-2x then 1x1 (the generator's bypass, sngan.py:66-70): every output parity plane is the same 1x1 conv
        for _ in range(4):
            for s in range(len(cins)):
                emit(s, s, 0, [(0, 1)])
        return ConvPlan(kind, cins, cout, 1, 4, len(sched) // 4, 1, sched, psched, out_parity=True)
    if kind.endswith("_dgrad"):
        return _plan_dgrad(kind[:-6], cins, cout, reuse_rows, emit, sched, psched, halo)
    raise ValueError(kind)


def _plan_dgrad(fwd_kind, cins, cout, reuse_rows, emit, sched, psched, halo=False):
    """dX = conv^T(dY): `cins` = (forward Cout,) are the channels of dY, `cout` the forward input channels produced.

    3x3 / 1x1:  dx[h,w] = sum_{r,q} dy[h+1-r, w+1-q] W[r,q]           -> same taps with the filter flipped
    4x4 / s2:   dx[2a+ph, 2b+pw] = sum over the two filter rows r = ph+1 (mod 2): a 2x2 conv of dy per input parity
                plane (exactly the structure of the nearest-2x forward, without tap sums)
    up2x + 3x3: dx[a,b] = sum over the four output parity planes of dy of 2x2 convs with the folded weights
                (exactly the structure of the stride-2 forward: four strided views of dy)"""
    assert len(cins) == 1
    if fwd_kind == "1x1":
        emit(0, 0, 0, [(0, 1)])
        return ConvPlan("1x1_dgrad", cins, cout, 1, 1, len(sched), 1, sched, psched, transposed=True)
    if fwd_kind == "3x3" and halo:                     # one halo box per 64-channel block of dy, flipped filter
        for c0, nv in _chunks(cins[0]):
            sched.append((0, -1, -1, c0))
            for t in range(9):
                r, q = divmod(t, 3)
                psched.append((c0, nv, 1 << ((2 - r) * 3 + (2 - q)), 0))
        return ConvPlan("3x3_dgrad", cins, cout, 9, 1, len(sched), 9, sched, psched, transposed=True, halo=True)
    if fwd_kind == "3x3":
        for q in range(3):
            emit(0, 0, q - 1, [(r - 1, 1 << ((2 - r) * 3 + (2 - q))) for r in range(3)])
        return ConvPlan("3x3_dgrad", cins, cout, 9, 1, len(sched), 3 if reuse_rows else 1, sched, psched,
                        transposed=True)
    if fwd_kind == "down4x4s2":
        # parity p of the input row: [(shift of dy row, filter row)] in increasing shift
        rows = {0: [(-1, 3), (0, 1)], 1: [(0, 2), (1, 0)]}
        for ph in range(2):
            for pw in range(2):
                for dq, q in rows[pw]:
                    emit(0, 0, dq, [(dr, 1 << (r * 4 + q)) for dr, r in rows[ph]])
        return ConvPlan("down4x4s2_dgrad", cins, cout, 16, 4, len(sched) // 4, 2 if reuse_rows else 1, sched, psched,
                        out_parity=True, transposed=True)
    if fwd_kind == "down3x3s2":
        # x row 2a+ph receives filter row r where 2a+ph = 2ho-1+r: ph = 0 -> r = 1 (ho = a); ph = 1 -> r = 2 (ho = a) and
        # r = 0 (ho = a+1).  Parity-0 rows / columns are padded with a zero-weight tap so that every parity plane has the
        # same 2x2 tap structure (uniform ngroups and R, as the kernel requires).
        rows = {0: [(0, 1), (1, None)], 1: [(0, 2), (1, 0)]}
        for ph in range(2):
            for pw in range(2):
                for dq, q in rows[pw]:
                    emit(0, 0, dq, [(dr, 0 if (r is None or q is None) else 1 << (r * 3 + q)) for dr, r in rows[ph]])
        return ConvPlan("down3x3s2_dgrad", cins, cout, 9, 4, len(sched) // 4, 2 if reuse_rows else 1, sched, psched,
                        out_parity=True, transposed=True)
    if fwd_kind == "up2x3x3":
        # forward window of parity p: [(dh, filter rows)]; the gradient reads dy plane p at the negated shift
        win = {0: [(-1, (0,)), (0, (1, 2))], 1: [(0, (0, 1)), (1, (2,))]}
        for pa in range(2):
            for pb in range(2):
                for dw, qset in sorted((-dw_, qs) for dw_, qs in win[pb]):
                    rows = []
                    for dh, rset in sorted((-dh_, rs) for dh_, rs in win[pa]):
                        mask = 0
                        for r in rset:
                            for q in qset:
                                mask |= 1 << (r * 3 + q)
                        rows.append((dh, mask))
                    emit(pa * 2 + pb, 0, dw, rows)
        return ConvPlan("up2x3x3_dgrad", cins, cout, 9, 1, len(sched), 2 if reuse_rows else 1, sched, psched,
                        n_views=4, transposed=True)
    raise ValueError(fwd_kind)


def n_tiling(cout: int, full_row: bool, long_k: bool = False) -> Tuple[int, int]:
    """(n_rows, n_tile): packed row count and output channels per CTA.

    ``full_row`` (RMSNorm / sum-of-squares epilogues) keeps every channel of a pixel in one CTA (<= 512 TMEM
    columns).  Otherwise wide outputs are split into 128-channel tiles so that several CTAs share an SM -- or, for layers
    with a long K loop (``long_k``: >= 16 K blocks, which run as CTA pairs), into 256-channel tiles: a pair then reads each
    A box for twice as many columns and the B operand fetch per MMA balances the tensor time (4x4-level 3x3 convs of the
    RC-49 model: 52 -> 41 us, profiles/r2_notes.md).
    """
    pad32 = (cout + 31) // 32 * 32
    if full_row:
        if pad32 > 512:
            raise ValueError(f"a fused channel-norm epilogue needs Cout <= 512 (got {cout})")
        return pad32, pad32
    if pad32 <= 256:
        return pad32, pad32
    n_tile = 256 if long_k else 128
    return (cout + n_tile - 1) // n_tile * n_tile, n_tile
