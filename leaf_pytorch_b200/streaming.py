"""Chunked / streaming inference with carried PCEN state (BASELINE.json configs[4]: long-form
audio, "chunked PCEN with carried IIR state").

The reference has no streaming mode: its smoother always restarts from the first frame
(reference postprocessing.py:15) and the whole (B,2F,T) activation is materialised, so long clips
are handled only by brute force.  Here the time axis is processed in chunks of frames; each chunk
re-reads the K-1 halo samples it shares with its neighbours and receives / returns the smoother
state, so the concatenated result equals the un-chunked forward of the whole clip (checked against
the reference's 60 s golden vector in tests/test_streaming_gpu.py).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import functional as LF


def needed_samples(spec: LF.LeafSpec, T_total: int, n_begin: int, n_count: int):
    """[lo, hi): samples of the clip that frames [n_begin, n_begin+n_count) depend on."""
    K, H = spec.K, spec.H
    pad_l = K // 2 + K % 2 - 1
    lo = max(0, n_begin * H - 2 * pad_l)
    hi = min(T_total, (n_begin + n_count - 1) * H - 2 * pad_l + 2 * K - 1)
    return lo, hi


def forward_chunked(leaf, x: torch.Tensor, chunk_frames: int = 1000) -> torch.Tensor:
    """Whole-clip features computed chunk by chunk over frames; ``x`` (B,1,T) is device resident.
    Peak scratch is that of one chunk instead of the whole clip.  No autograd."""
    leaf._require_plain("forward_chunked")
    spec = leaf.spec
    prm = [None if p is None else p.detach() for p in leaf._param_tuple()]
    B, _, T = x.shape
    N = spec.num_frames(T)
    out = torch.empty((B, spec.F, N), dtype=spec.out_dtype, device=x.device)     # bf16 features if the module asks
    if x.dtype not in (torch.float32, torch.int16) or not x.is_contiguous():
        x = LF._check_input(x)
    # one scratch buffer for every chunk: the banks written for the first chunk serve the others (same parameters),
    # and the chunks read their sample windows in place (views of x, no copies)
    ws = torch.empty(LF.workspace_bytes(spec, B, min(chunk_frames, N), x.dtype), dtype=torch.uint8, device=x.device)
    state = None
    for n0 in range(0, N, chunk_frames):
        cnt = min(chunk_frames, N - n0)
        lo, hi = needed_samples(spec, T, n0, cnt)
        _, state = LF.forward_window(spec, x[:, :, lo:hi], T, lo, n0, cnt, *prm, ema_state=state,
                                     out=out[:, :, n0:n0 + cnt], workspace=ws, reuse_banks=n0 > 0)
    return out


class LeafStream:
    """Online frontend for an unbounded stream: push sample blocks, get the frames that became
    computable.  Keeps the tail of the waveform (<= 2K samples) and the PCEN smoother state.

    Frames are emitted as soon as every sample of their analysis window has arrived, so the
    output is identical to the offline forward of the full signal except for the frames whose
    window would reach past the (unknown) end; ``flush()`` emits those, zero-padding like the
    reference does at a clip's end (reference convolution.py:92, pooling.py:37)."""

    def __init__(self, leaf, batch: int, device=None):
        leaf._require_plain("LeafStream")
        self.leaf = leaf
        self.spec = leaf.spec
        self.B = batch
        self.device = torch.device(device) if device is not None else leaf._complex_conv._kernel.device
        self.buf = torch.empty((batch, 1, 0), dtype=torch.float32, device=self.device)
        self.buf_off = 0          # absolute index of buf[...,0]
        self.n_seen = 0           # samples received so far
        self.n_done = 0           # frames emitted so far
        self.state: Optional[torch.Tensor] = None

    def _emit(self, T_total: int, n_count: int) -> torch.Tensor:
        prm = [None if p is None else p.detach() for p in self.leaf._param_tuple()]
        lo, hi = needed_samples(self.spec, T_total, self.n_done, n_count)
        win = self.buf[:, :, lo - self.buf_off:hi - self.buf_off].contiguous()
        out, self.state = LF.forward_window(self.spec, win, T_total, lo, self.n_done, n_count, *prm,
                                            ema_state=self.state)
        self.n_done += n_count
        # drop samples no future frame needs
        keep_from, _ = needed_samples(self.spec, 1 << 29, self.n_done, 1)
        cut = max(0, min(keep_from, self.n_seen) - self.buf_off)
        if cut:
            self.buf = self.buf[:, :, cut:].contiguous()
            self.buf_off += cut
        return out

    def push(self, block: torch.Tensor) -> torch.Tensor:
        """block (B,1,t) -> (B,F,n_new) frames (n_new may be 0)."""
        block = block.to(self.device, torch.float32)
        self.buf = torch.cat([self.buf, block], dim=2)
        self.n_seen += block.shape[2]
        K, H = self.spec.K, self.spec.H
        pad_l = K // 2 + K % 2 - 1
        # frame n is final once sample n*H - 2*pad_l + 2K - 2 has arrived
        n_ready = (self.n_seen - 1 + 2 * pad_l - 2 * K + 2) // H + 1 if self.n_seen >= 2 * K - 1 - 2 * pad_l else 0
        n_ready = max(0, n_ready)
        if n_ready <= self.n_done:
            return torch.empty((self.B, self.spec.F, 0), dtype=self.spec.out_dtype, device=self.device)
        # pretend the clip is very long: only frames whose window is complete are requested
        return self._emit(self.n_seen + (1 << 20), n_ready - self.n_done)

    def flush(self) -> torch.Tensor:
        """End of stream: emit the remaining frames with the true clip length."""
        N = self.spec.num_frames(self.n_seen) if self.n_seen > 0 else 0
        if N <= self.n_done:
            return torch.empty((self.B, self.spec.F, 0), dtype=self.spec.out_dtype, device=self.device)
        return self._emit(self.n_seen, N - self.n_done)
