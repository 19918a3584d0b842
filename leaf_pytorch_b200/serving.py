"""Host-side inference pipeline: several batches in flight from pinned host buffers.

``HostPipeline`` wraps ``leafk_forward_host_async``.  Each in-flight batch owns a buffer set (device
waveform, device features, workspace, two events); the library orders three streams so that the H2D copy of
batch i+1 runs under the kernels of batch i and the D2H of batch i under the kernels of batch i+1.  Inside a
batch the persistent tensor-core kernel starts on the first slice of the copy (ready flags), so a single
batch is already pipelined; this class removes the remaining per-call bubbles of a synchronous loop.

    pipe = HostPipeline(leaf, B, T, depth=2)
    t0 = pipe.submit(x0)                 # pinned (B,1,T) float32 or int16
    t1 = pipe.submit(x1)
    y0 = pipe.result(t0)                 # pinned (B,F,N) float32 (bfloat16 if leaf.out_dtype is), valid on return
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _native as N
from . import functional as LF


class _Set:
    __slots__ = ("dev_x", "dev_out", "ws", "ev_compute", "ev_out", "out_host", "busy", "keep", "ticket", "status")


class HostPipeline:
    def __init__(self, leaf, batch: int, n_samples: int, depth: int = 2, n_slices: int = 2,
                 input_dtype: torch.dtype = torch.float32, device=None):
        self.lib = N.lib()
        leaf._require_plain("HostPipeline")
        self.leaf = leaf
        self.spec = leaf.spec
        self.B, self.T = int(batch), int(n_samples)
        self.n_frames = self.spec.num_frames(self.T)
        self.dtype = input_dtype
        self.n_slices = n_slices
        self.device = torch.device(device) if device is not None else leaf._complex_conv._kernel.device
        if self.device.type != "cuda":
            raise N.LeafNativeError("HostPipeline needs the module on a CUDA device; there is no CPU fallback")
        if not LF.tc_supported(self.spec.F, self.spec.K, self.spec.H) or self.spec.algo == "fp32":
            raise N.LeafNativeError("HostPipeline needs the tensor-core kernel (geometry not covered or algo='fp32')")
        self.out_dtype = self.spec.out_dtype          # fixed at construction: the buffer sets are typed
        self.cfg = self.spec.config(input_dtype, out_dtype=self.out_dtype)
        with torch.cuda.device(self.device):
            ws_bytes = self.lib.leafk_workspace_bytes(C.byref(self.cfg), self.B, self.n_frames)
            self.copy_stream = torch.cuda.Stream(device=self.device)
            self.d2h_stream = torch.cuda.Stream(device=self.device)
            self.compute_stream = torch.cuda.Stream(device=self.device)
            self.sets: List[_Set] = []
            for _ in range(max(1, depth)):
                s = _Set()
                s.dev_x = torch.empty(self.B * self.T, dtype=input_dtype, device=self.device)
                s.dev_out = torch.empty(self.B * self.spec.F * self.n_frames, dtype=self.out_dtype, device=self.device)
                s.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
                s.ev_compute = self.lib.leafk_event_create()
                s.ev_out = self.lib.leafk_event_create()
                if not s.ev_compute or not s.ev_out:
                    raise N.LeafNativeError("could not create CUDA events")
                s.status = torch.zeros(1, dtype=torch.int32).pin_memory()       # asynchronous error word, arrives with the result
                s.out_host, s.busy, s.keep, s.ticket = None, False, None, -1
                self.sets.append(s)
        self._next = 0                    # monotonically increasing ticket of the next batch
        self._done = {}                   # ticket -> features of batches collected early (their set was needed again)

    def submit(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> int:
        """Enqueue one batch; returns a ticket for ``result`` (tickets never repeat).  Blocks only when every buffer
        set is in flight; the batch that has to make room is collected and kept until its ticket is asked for."""
        if x_host.is_cuda or x_host.dtype != self.dtype or tuple(x_host.shape) != (self.B, 1, self.T):
            raise ValueError(f"x_host must be a CPU {self.dtype} tensor of shape {(self.B, 1, self.T)}")
        ticket = self._next
        self._next += 1
        s = self.sets[ticket % len(self.sets)]
        if s.busy:
            self._done[s.ticket] = self._collect(s)       # oldest batch not collected yet: wait for it, keep its result
        # parameters may have been updated on the caller's stream (optimizer.step, load_state_dict, .to()): the
        # pipeline's own streams must see those writes
        self.compute_stream.wait_stream(torch.cuda.current_stream(self.device))
        if out_host is None:
            out_host = torch.empty((self.B, self.spec.F, self.n_frames), dtype=self.out_dtype, pin_memory=True)
        elif (out_host.is_cuda or out_host.dtype != self.out_dtype or not out_host.is_contiguous()
              or tuple(out_host.shape) != (self.B, self.spec.F, self.n_frames)):
            raise ValueError(f"out_host must be a contiguous CPU {self.out_dtype} tensor of shape "
                             f"{(self.B, self.spec.F, self.n_frames)}")
        x_host = x_host.contiguous()
        prm_t = [None if p is None else p.detach() for p in self.leaf._param_tuple()]
        prm, keep = LF._params_struct(self.spec, *prm_t, self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.leafk_forward_host_async(
                C.byref(self.cfg), C.byref(prm), C.c_void_p(x_host.data_ptr()), self.B, self.T,
                C.c_void_p(out_host.data_ptr()), int(self.n_slices), C.c_void_p(s.dev_x.data_ptr()),
                C.c_void_p(s.dev_out.data_ptr()), C.c_void_p(s.ws.data_ptr()), s.ws.numel(),
                C.c_void_p(self.compute_stream.cuda_stream), C.c_void_p(self.copy_stream.cuda_stream),
                C.c_void_p(self.d2h_stream.cuda_stream), C.c_void_p(s.ev_compute), C.c_void_p(s.ev_out),
                C.c_void_p(s.status.data_ptr()))
        N.check(rc, "leafk_forward_host_async")
        s.out_host, s.busy, s.keep, s.ticket = out_host, True, (keep, x_host), ticket
        return ticket

    def _collect(self, s: _Set) -> torch.Tensor:
        N.check(self.lib.leafk_event_synchronize(C.c_void_p(s.ev_out)), "leafk_event_synchronize")
        s.busy, s.keep = False, None
        # a stalled host-to-device slice is reported here (the kernels record it instead of trapping; the word was copied
        # back right behind the features, so reading it costs nothing)
        N.check(self.lib.leafk_status_message(int(s.status[0])), "leafk_forward_host_async")
        return s.out_host

    def result(self, ticket: int) -> torch.Tensor:
        """Block until the batch of ``ticket`` is on the host and return its (B,F,N) features.  Each ticket can be
        collected once; an unknown or already collected ticket raises."""
        if ticket in self._done:
            return self._done.pop(ticket)
        s = self.sets[ticket % len(self.sets)] if 0 <= ticket < self._next else None
        if s is None or not s.busy or s.ticket != ticket:
            raise ValueError(f"no batch in flight for ticket {ticket}")
        return self._collect(s)

    def close(self) -> None:
        for s in self.sets:
            if s.busy:
                self._collect(s)
            if s.ev_compute:
                self.lib.leafk_event_destroy(C.c_void_p(s.ev_compute)); s.ev_compute = None
            if s.ev_out:
                self.lib.leafk_event_destroy(C.c_void_p(s.ev_out)); s.ev_out = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
