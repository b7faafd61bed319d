"""EventDetector -- same constructor / run() / Event surface as the reference
(event_detection/event_detector.py:14-83), executed by the K1 CUDA kernel."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class Event:
    """event_detector.py:14-24"""
    __slots__ = ("start", "length", "mean", "stdv")

    def __init__(self, start: int, length: int, mean: float, stdv: float) -> None:
        self.start, self.length, self.mean, self.stdv = start, length, mean, stdv

    @property
    def end(self) -> int:
        return self.start + self.length

    def __repr__(self):
        return f"Event(start={self.start}, length={self.length}, mean={self.mean:.3f}, stdv={self.stdv:.3f})"


class EventDetector:
    """Two-window t-statistic segmentation on the GPU.

    ``run(raw)`` keeps the reference contract (one read -> list[Event]).
    ``detect_batch`` is the batched form the data path uses: a ragged set of
    reads in one launch, results left on the device as a structure of arrays."""

    def __init__(self, window_length1=3, window_length2=6, threshold1=1.4, threshold2=9., peak_height=0.2,
                 device: int | None = None, warmup: int = -1):
        self.params = {'window_length1': int(window_length1), 'window_length2': int(window_length2),
                       'threshold1': float(threshold1), 'threshold2': float(threshold2),
                       'peak_height': float(peak_height)}
        if _lib.device_count() == 0:
            raise _lib.RavventError(_lib.RVB_ERR_CUDA, "no CUDA device: the event scan has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.warmup = warmup

    def reset(self):   # the reference's streaming state does not exist here; kept for API compatibility
        return None

    # ------------------------------------------------------------------
    def detect_batch(self, signal, read_offsets):
        """signal: 1-D integer tensor/array with all reads concatenated; read_offsets: n_reads+1.
        -> dict(start,length [int32], mean,stdv [float64], count [int32] per read,
                event_offsets [int64 numpy], all device tensors except the offsets)."""
        offs = np.ascontiguousarray(np.asarray(read_offsets, dtype=np.int64))
        n_reads = offs.size - 1
        if isinstance(signal, torch.Tensor):
            sig = signal.to(self.device)
            if sig.dtype not in (torch.int16, torch.int32):
                sig = sig.to(torch.int32)
        else:
            arr = np.asarray(signal)
            if arr.dtype != np.int16:
                if not np.issubdtype(arr.dtype, np.integer):
                    if not np.all(arr == np.rint(arr)):
                        raise ValueError("EventDetector expects integer-valued raw samples (data_loader.py:114)")
                arr = arr.astype(np.int32)
            sig = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
        sig = sig.contiguous()
        lens = np.diff(offs)
        ev_offs = np.zeros(n_reads + 1, dtype=np.int64)
        np.cumsum(lens // 2 + 2, out=ev_offs[1:])
        cap = int(ev_offs[-1])
        with torch.cuda.device(self.device):
            start = torch.empty(cap, dtype=torch.int32, device=self.device)
            length = torch.empty(cap, dtype=torch.int32, device=self.device)
            mean = torch.empty(cap, dtype=torch.float64, device=self.device)
            stdv = torch.empty(cap, dtype=torch.float64, device=self.device)
            count = torch.zeros(max(n_reads, 1), dtype=torch.int32, device=self.device)
            nbytes = C.c_size_t(0)
            _lib.check(_lib.lib.rvb_event_detect_workspace_bytes(offs.ctypes.data, n_reads, C.byref(nbytes)))
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            p = self.params
            _lib.check(_lib.lib.rvb_event_detect(
                sig.data_ptr(), sig.element_size(), offs.ctypes.data, n_reads,
                p['window_length1'], p['window_length2'], p['threshold1'], p['threshold2'], p['peak_height'],
                ev_offs.ctypes.data, start.data_ptr(), length.data_ptr(), mean.data_ptr(), stdv.data_ptr(),
                count.data_ptr(), ws.data_ptr(), nbytes.value, self.warmup,
                torch.cuda.current_stream(self.device).cuda_stream))
        return {"start": start, "length": length, "mean": mean, "stdv": stdv, "count": count[:n_reads],
                "event_offsets": ev_offs}

    def run(self, raw):
        """One read -> list[Event] (event_detector.py:75-83)."""
        raw = np.asarray(raw)
        out = self.detect_batch(raw, [0, raw.size])
        n = int(out["count"][0].item()) if raw.size else 0
        st = out["start"][:n].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        ln = out["length"][:n].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        mu = out["mean"][:n].cpu().numpy()
        sd = out["stdv"][:n].cpu().numpy()
        return [Event(int(a), int(b), float(c), float(d)) for a, b, c, d in zip(st, ln, mu, sd)]
