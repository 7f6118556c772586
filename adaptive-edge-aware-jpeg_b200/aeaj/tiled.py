"""Halo-split of ONE large image over G GPUs (BASELINE config C4).

Design: shard the heavy, replicate the cheap.  Rank r owns the band of full-resolution rows
[r*H/G, (r+1)*H/G) (a multiple of 256 rows, so no chroma cell, filter tile or quadtree leaf straddles
two ranks).  The pipeline runs phase by phase through ``aeaj_encode_phase`` and between phases the small
planes are exchanged over NVLink with NCCL:

    colour + subsample + u8 cast (band)
    CLAHE histograms (band)            -> all-reduce SUM  (3 x 16 x 256 counters)
                                       -> all-gather      u8 planes            (1.5 B / px)
    CLAHE LUT + Gaussian + bilateral (band), histogram of the result
                                       -> all-reduce SUM  (3 x 256 counters), all-gather filtered u8 planes
    thresholds + Sobel / NMS (band)    -> all-gather      strong / weak bitmaps (2 bit / sample)
    hysteresis + quadtree              replicated on every rank (bit maps only; no exchange rounds needed)
    DCT + quantise (band's leaves)     (optional: all-reduce SUM of the zero-initialised coefficient stream)

so every rank ends with the complete, reference-ordered result (leaves, states, coefficients), bit-identical
to the single-GPU path.  Decode mirrors it: IDCT per band, all-gather of the chroma layers (the bilinear
upsample reads one row across the band edge), inverse colour per band.

``emulate`` runs the G bands one after the other on ONE GPU through exactly the same phase calls (the buffers
are shared, so no collective is needed); the GPU tests use it to check the band-restricted kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import native
from .codec import DeviceCodec, EncodedBatch, _stream

PH_COLOR, PH_HIST, PH_PREFILTER, PH_NMS, PH_TREE, PH_DCT = range(6)
DPH_IDCT, DPH_COLOR = range(2)
BAND_ALIGN = 256


def band_of(rank: int, world: int, H: int, block_max: int = 128):
    """rows [lo, hi) of the full-resolution image owned by `rank` (bands are multiples of 2 x max(128, block_max) rows)"""
    align = max(BAND_ALIGN, 2 * block_max)
    if H % (world * align) != 0:
        raise ValueError(f"halo-split needs the image height ({H}) to be a multiple of {align} x world size ({world})")
    hb = H // world
    return rank * hb, (rank + 1) * hb


class TiledCodec:
    def __init__(self, codec: DeviceCodec, rank: int = 0, world: int = 1, group=None, emulate: int = 0):
        self.codec, self.rank, self.world, self.group, self.emulate = codec, rank, world, group, emulate
        self.lib = codec.lib

    # -- views of the plan workspace ------------------------------------------------------------
    def _views(self, p):
        buf = native.PlanBuffers()
        native.check(self.lib.aeaj_plan_buffers(p.ptr, p.workspace.data_ptr(), C.byref(buf)), "aeaj_plan_buffers")
        base = p.workspace.data_ptr()

        def view(ptr, nbytes, dtype, shape):
            off = ptr - base
            return p.workspace[off:off + nbytes].view(dtype).view(shape)

        v = dict(u8a=[], u8b=[], strong=[], weak=[], layer=[])
        for l in range(3):
            h, w, wpr = buf.h[l], buf.w[l], buf.wpr[l]
            v["u8a"].append(view(buf.u8a[l], h * w, torch.uint8, (h, w)))
            v["u8b"].append(view(buf.u8b[l], h * w, torch.uint8, (h, w)))
            v["strong"].append(view(buf.strong[l], h * wpr * 4, torch.int32, (h, wpr)))
            v["weak"].append(view(buf.weak[l], h * wpr * 4, torch.int32, (h, wpr)))
            v["layer"].append(view(buf.layer[l], h * w * 4, torch.float32, (h, w)))
        v["clahe_hist"] = view(buf.clahe_hist, buf.clahe_hist_bytes, torch.int32, (-1,))
        v["hist"] = view(buf.hist, buf.hist_bytes, torch.int32, (-1,))
        return v

    def _gather_rows(self, full: torch.Tensor, H_layer: int):
        """in-place all-gather of equal row bands of a (rows, cols) plane"""
        import torch.distributed as dist
        rows = H_layer // self.world
        dist.all_gather_into_tensor(full.view(-1), full[self.rank * rows:(self.rank + 1) * rows].reshape(-1), group=self.group)

    def _reduce(self, t: torch.Tensor):
        import torch.distributed as dist
        dist.all_reduce(t, group=self.group)

    # -- encode ---------------------------------------------------------------------------------
    def encode(self, rgb: torch.Tensor, H: int, W: int, space: str, qrange, brange, exchange_coef: bool = False) -> EncodedBatch:
        """rgb: float32 CUDA tensor.  Real multi-rank mode: this rank's band [Hb, W, 3].  Emulation: the whole [H, W, 3].
        Leaves, states and counts are complete on every rank.  Coefficients: each rank holds the ranges of its own band's
        leaves (at their global offsets) -- all its decoder phase and its host-side D2H need; `exchange_coef` additionally
        all-reduces the zero-filled streams so that every rank holds the whole stream (used by the parity check)."""
        c = self.codec
        p = c._plan(1, H, W, space, brange, qrange)
        o = p.out
        io = native.EncodeIO()
        for l in range(3):
            io.coef[l], io.leaves[l], io.states[l] = o.coef[l].data_ptr(), o.leaves[l].data_ptr(), o.states[l].data_ptr()
        io.counts, io.status = o.counts.data_ptr(), o.status.data_ptr()
        ws = p.workspace.data_ptr()
        G = self.emulate or self.world
        ranks = range(G) if self.emulate else [self.rank]
        rgb = rgb.contiguous()

        def run(phase, r):
            lo, hi = band_of(r, G, H, brange[1])
            # the kernel indexes rows from the top of the full image: hand it the (virtual) address of row 0
            io.rgb = rgb.data_ptr() - (0 if self.emulate else lo * W * 12)
            native.check(self.lib.aeaj_encode_phase(p.ptr, C.byref(io), ws, _stream(), phase, lo, hi), f"aeaj_encode_phase({phase})")

        multi = not self.emulate and self.world > 1
        v = self._views(p) if multi else None
        for r in ranks:
            run(PH_COLOR, r)                  # also clears the histogram accumulators (before any histogram work)
        for r in ranks:
            run(PH_HIST, r)
        if multi:
            self._reduce(v["clahe_hist"])
            for l in range(3):
                self._gather_rows(v["u8a"][l], p.info.layer_h[l])
        for r in ranks:
            run(PH_PREFILTER, r)
        if multi:
            self._reduce(v["hist"])
            for l in range(3):
                self._gather_rows(v["u8b"][l], p.info.layer_h[l])
        for r in ranks:
            run(PH_NMS, r)
        if multi:
            for l in range(3):
                self._gather_rows(v["strong"][l], p.info.layer_h[l])
                self._gather_rows(v["weak"][l], p.info.layer_h[l])
        run(PH_TREE, ranks[0])                # whole image, replicated (the band argument is irrelevant for this phase)
        if multi and exchange_coef:
            for l in range(3):
                o.coef[l].zero_()
        for r in ranks:
            run(PH_DCT, r)
        if multi and exchange_coef:
            for l in range(3):
                self._reduce(o.coef[l])
        return o

    # -- decode ---------------------------------------------------------------------------------
    def decode(self, enc: EncodedBatch, H: int, W: int, space: str, qrange, brange) -> torch.Tensor:
        """Returns the full [H, W, 3] buffer; in real multi-rank mode only this rank's band rows are valid."""
        c = self.codec
        p = c._plan(1, H, W, space, brange, qrange)
        io = native.DecodeIO()
        for l in range(3):
            io.coef[l], io.leaves[l] = enc.coef[l].data_ptr(), enc.leaves[l].data_ptr()
        io.counts, io.rgb = enc.counts.data_ptr(), p.rgb_out.data_ptr()
        ws = p.workspace.data_ptr()
        G = self.emulate or self.world
        ranks = range(G) if self.emulate else [self.rank]
        multi = not self.emulate and self.world > 1

        def run(phase, r):
            lo, hi = band_of(r, G, H, brange[1])
            native.check(self.lib.aeaj_decode_phase(p.ptr, C.byref(io), ws, _stream(), phase, lo, hi), f"aeaj_decode_phase({phase})")

        for r in ranks:
            run(DPH_IDCT, r)
        if multi:
            v = self._views(p)
            for l in (1, 2):
                if p.info.layer_h[l] != H or p.info.layer_w[l] != W:
                    self._gather_rows(v["layer"][l], p.info.layer_h[l])
        for r in ranks:
            run(DPH_COLOR, r)
        return p.rgb_out[0]
