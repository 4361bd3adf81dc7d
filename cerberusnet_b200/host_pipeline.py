"""Host-buffer pipeline for the fused level op: inputs and outputs live in (pinned) host memory.

`cerb_warp_corr_forward_host` (C ABI) does H2D -> kernel -> D2H on ONE stream, which is what a
single TensorRT-style `enqueue` can do.  A caller that streams image pairs from the host can do
better: PCIe is full duplex, so the copy-in of pair k+1, the kernels of pair k and the copy-out of
pair k-1 overlap when they run on three streams over double-buffered device staging.  This class
is that pipeline; `bench.py` reports it as the end-to-end number.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import WARP_TORCH

LevelShape = Tuple[int, int, int, bool]  # (C, H, W, warped)


class HostPipeline:
    def __init__(self, levels: Sequence[LevelShape], batch: int = 1, depth: int = 2, device=None, pad_size: int = 4,
                 max_displacement: int = 4, warp_mode: int = WARP_TORCH, leaky_slope: Optional[float] = 0.1,
                 dtype: torch.dtype = torch.float32, x2_roll: int = 0, shared_features: bool = False):
        """``shared_features`` with ``x2_roll = batch // 2``: each level takes ONE feature tensor holding
        [features of image 1; features of image 2] that serves as both correlation inputs -- both flow
        directions per launch, and every feature map crosses PCIe once instead of twice."""
        self.device = torch.device(device if device is not None else "cuda")
        self.levels, self.batch, self.depth = list(levels), batch, depth
        self.cfg = (pad_size, 1, max_displacement, 1, 1, 1, warp_mode, leaky_slope)
        self.x2_roll, self.shared = int(x2_roll), bool(shared_features)
        d2 = (2 * max_displacement + 1) ** 2
        self.slots = []
        for _ in range(depth):
            bufs = []
            for (C, H, W, warped) in self.levels:
                oh, ow = H + 2 * pad_size - 2 * max_displacement, W + 2 * pad_size - 2 * max_displacement
                bufs.append((torch.empty(batch, C, H, W, dtype=dtype, device=self.device),
                             None if self.shared else torch.empty(batch, C, H, W, dtype=dtype, device=self.device),
                             torch.empty(batch, 2, H, W, dtype=torch.float32, device=self.device) if warped else None,
                             torch.empty(batch, d2, oh, ow, dtype=dtype, device=self.device)))
            self.slots.append(bufs)
        # optional packed arenas: one pinned host buffer + one device buffer per direction per slot,
        # so a pyramid pass costs ONE memcpy each way instead of 3 per level / 1 per level
        self._arenas = None
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self._n = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for lv in self.slots[0] for t in lv[:3] if t is not None)
        self.d2h_bytes = sum(lv[3].numel() * lv[3].element_size() for lv in self.slots[0])

    def submit(self, host_in: Sequence[Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]],
               host_out: Sequence[torch.Tensor]) -> None:
        """Queue one pyramid pass: `host_in[l] = (x1, x2, flow-or-None)` and `host_out[l]` are host
        tensors (pinned for real overlap).  Returns immediately; call `synchronize()` before
        reading `host_out`."""
        k = self._n % self.depth
        bufs = self.slots[k]
        # (waiting on an event that was never recorded is a no-op, so the first uses of a slot need no special case)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_cmp[k])   # the kernels that last read this slot's inputs are done
            for (d1, d2_, dfl, _), (h1, h2, hfl) in zip(bufs, host_in):
                d1.copy_(h1, non_blocking=True)
                if d2_ is not None:
                    d2_.copy_(h2, non_blocking=True)
                if dfl is not None:
                    dfl.copy_(hfl, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(self.ev_in[k])
            self.s_cmp.wait_event(self.ev_out[k])  # the previous result of this slot has left the device
            self._launch(bufs)
            self.ev_cmp[k].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_cmp[k])
            for (_, _, _, dout), hout in zip(bufs, host_out):
                hout.copy_(dout, non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self._n += 1

    def _launch(self, bufs) -> None:
        pad, ks, md, s1, s2, mult, mode, slope = self.cfg
        for (d1, d2_, dfl, dout) in bufs:
            ops.warp_corr_forward(d1, d1 if d2_ is None else d2_, dfl, pad, ks, md, s1, s2, mult, mode, slope, out=dout,
                                  x2_roll=self.x2_roll)

    # ---------------------------------------------------------------- packed arenas
    def enable_arenas(self) -> None:
        """Re-home the staging buffers inside contiguous arenas (device) mirrored by pinned host
        arenas.  Fill `host_inputs(slot)` views, call `submit_packed(slot)`, read
        `host_outputs(slot)` after `synchronize()`."""
        if self._arenas is not None:
            return
        def align(n):  # keep every view 256-byte aligned (TMA needs 16)
            return (n + 63) // 64 * 64
        dtype = self.slots[0][0][0].dtype
        if dtype != torch.float32:
            raise NotImplementedError("arenas are implemented for float32 staging")
        n_in = sum(align(t.numel()) for lv in self.slots[0] for t in lv[:3] if t is not None)
        n_out = sum(align(lv[3].numel()) for lv in self.slots[0])
        self._arenas = []
        for k in range(self.depth):
            d_in = torch.empty(n_in, dtype=torch.float32, device=self.device)
            d_out = torch.empty(n_out, dtype=torch.float32, device=self.device)
            h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
            h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
            oi = oo = 0
            dev_views, host_in_views, host_out_views = [], [], []
            for lv in self.slots[k]:
                dv, hv = [], []
                for t in lv[:3]:
                    if t is None:
                        dv.append(None); hv.append(None)
                        continue
                    n = t.numel()
                    dv.append(d_in[oi:oi + n].view(t.shape)); hv.append(h_in[oi:oi + n].view(t.shape))
                    oi += align(n)
                n = lv[3].numel()
                dv.append(d_out[oo:oo + n].view(lv[3].shape))
                host_out_views.append(h_out[oo:oo + n].view(lv[3].shape))
                oo += align(n)
                dev_views.append(tuple(dv)); host_in_views.append(tuple(hv))
            self.slots[k] = dev_views
            self._arenas.append((d_in, d_out, h_in, h_out, host_in_views, host_out_views))

    def host_inputs(self, slot: int):
        return self._arenas[slot][4]

    def host_outputs(self, slot: int):
        return self._arenas[slot][5]

    def submit_packed(self, slot: int) -> None:
        """One pyramid pass from the pinned arena of `slot`: a single H2D copy, the kernels, a single
        D2H copy, on the three streams."""
        k = slot
        d_in, d_out, h_in, h_out, _, _ = self._arenas[k]
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_cmp[k])    # no-op until the slot has been used once
            d_in.copy_(h_in, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(self.ev_in[k])
            self.s_cmp.wait_event(self.ev_out[k])
            self._launch(self.slots[k])
            self.ev_cmp[k].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_cmp[k])
            h_out.copy_(d_out, non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self._n += 1

    def synchronize(self) -> None:
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.synchronize()
