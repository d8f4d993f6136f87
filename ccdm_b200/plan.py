"""Host-side planning for the tap-GEMM engine: tile boxes, K-block schedules and weight-packing schedules.

Pure Python (no CUDA), so the CPU test-suite can check every schedule against ``torch.nn.functional.conv2d``.

A K block is (source view, dw, dh, c0): 64 input channels starting at ``c0`` of view ``src`` read at the output
position shifted by (dh, dw).  The matching packing entry (cin0, nvalid, tapmask) says which slice of the fp32
``[Cout][Cin_total][taps]`` weight lands in that block (``tapmask`` sums several filter taps: nearest-2x fold).

Reference call sites: CCDM_unified/models/unet.py:74-81 (Upsample / Downsample), :139 (3x3), :165,195,198 (1x1).
"""
from __future__ import annotations

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


def tile_box(gw: int, gh: int, force_tb1: bool = False) -> Tuple[int, int, int]:
    """(tw, th, tb) with tw*th*tb == 128 for an output grid of gw x gh positions per sample.

    Prefers boxes that divide the grid exactly (no masked rows); falls back to the next power of two for odd
    sizes (3x3, 6x6 bottlenecks of the 192-px model).  ``force_tb1`` keeps a tile inside one sample (per-sample
    weights) at the price of masked rows on tiny grids.
    """
    tw = _pow2_floor_div(gw, 64)
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


@dataclass
class ConvPlan:
    """Everything ccdm_tapgemm / ccdm_pack_weights need for one layer, except pointers."""
    kind: str
    cins: Tuple[int, ...]                 # channels of each concatenated source
    cout: int
    ntaps: int                            # kh*kw of the fp32 weight
    nz: int
    nkb: int
    sched: List[Tuple[int, int, int, int]] = field(default_factory=list)    # [nz*nkb] (src, dw, dh, c0)
    psched: List[Tuple[int, int, int, int]] = field(default_factory=list)   # [nz*nkb] (cin0, nvalid, tapmask, 0)
    n_views: int = 1                      # views per source tensor (4 parity planes for the stride-2 conv)
    out_parity: bool = False              # nz == 4 output parity planes (nearest-2x + 3x3)
    stride: int = 1

    @property
    def n_src(self) -> int:
        return len(self.cins) * self.n_views


def _chunks(c: int):
    return [(c0, min(KB, c - c0)) for c0 in range(0, c, KB)]


def plan_conv(kind: str, cins: Sequence[int], cout: int) -> ConvPlan:
    """kind: '1x1' | '3x3' | 'down4x4s2' | 'up2x3x3'."""
    cins = tuple(int(c) for c in cins)
    offs = [sum(cins[:i]) for i in range(len(cins))]
    sched, psched = [], []
    if kind == "1x1":
        for s, c in enumerate(cins):
            for c0, nv in _chunks(c):
                sched.append((s, 0, 0, c0))
                psched.append((offs[s] + c0, nv, 1, 0))
        return ConvPlan(kind, cins, cout, 1, 1, len(sched), sched, psched)
    if kind == "3x3":
        for r in range(3):
            for q in range(3):
                for s, c in enumerate(cins):
                    for c0, nv in _chunks(c):
                        sched.append((s, q - 1, r - 1, c0))
                        psched.append((offs[s] + c0, nv, 1 << (r * 3 + q), 0))
        return ConvPlan(kind, cins, cout, 9, 1, len(sched), sched, psched)
    if kind == "down4x4s2":
        # out[ho,wo] = sum_{r,q} W[r,q] x[2ho-1+r, 2wo-1+q];  2ho-1+r = 2(ho+dr)+pr  with (dr,pr) below.
        # The 4 parity planes x[pr::2, pq::2] are separate strided TMA views: view index = src*4 + pr*2 + pq.
        split = {0: (-1, 1), 1: (0, 0), 2: (0, 1), 3: (1, 0)}
        for r in range(4):
            for q in range(4):
                dr, pr = split[r]
                dq, pq = split[q]
                for s, c in enumerate(cins):
                    for c0, nv in _chunks(c):
                        sched.append((s * 4 + pr * 2 + pq, dq, dr, c0))
                        psched.append((offs[s] + c0, nv, 1 << (r * 4 + q), 0))
        return ConvPlan(kind, cins, cout, 16, 1, len(sched), sched, psched, n_views=4, stride=2)
    if kind == "up2x3x3":
        # nearest-2x then 3x3/p1: output (2a+pa, 2b+pb) sees a 2x2 window of the low-res input; taps that hit the
        # same low-res pixel are summed at packing time (9 -> 4 taps, 2.25x fewer MACs).
        rows = {0: [(-1, (0,)), (0, (1, 2))], 1: [(0, (0, 1)), (1, (2,))]}
        for pa in range(2):
            for pb in range(2):
                for dh, rset in rows[pa]:
                    for dw, qset in rows[pb]:
                        mask = 0
                        for r in rset:
                            for q in qset:
                                mask |= 1 << (r * 3 + q)
                        for s, c in enumerate(cins):
                            for c0, nv in _chunks(c):
                                sched.append((s, dw, dh, c0))
                                psched.append((offs[s] + c0, nv, mask, 0))
        nkb = len(sched) // 4
        return ConvPlan(kind, cins, cout, 9, 4, nkb, sched, psched, out_parity=True)
    raise ValueError(kind)


def n_tiling(cout: int, full_row: bool) -> Tuple[int, int]:
    """(n_rows, n_tile): packed row count and output channels per CTA.

    ``full_row`` (RMSNorm / sum-of-squares epilogues) keeps every channel of a pixel in one CTA (<= 512 TMEM
    columns).  Otherwise wide outputs are split into 128-channel tiles so that several CTAs share an SM.
    """
    pad32 = (cout + 31) // 32 * 32
    if full_row:
        if pad32 > 512:
            raise ValueError(f"a fused channel-norm epilogue needs Cout <= 512 (got {cout})")
        return pad32, pad32
    if pad32 <= 256:
        return pad32, pad32
    n_tile = 128
    return (cout + n_tile - 1) // n_tile * n_tile, n_tile
