"""Functional entry points over libleafk.so: the fused LEAF forward / backward on CUDA tensors.

``leaf_forward`` is what ``Leaf.forward`` (frontend.py here, reference frontend.py:78-89) calls;
it is differentiable with respect to the seven frontend parameters (reference train.py:258 only
ever needs those) and the waveform.

Training (parameters require grad and autograd is recording): on geometries the tensor-core training kernel covers
the forward is ``leafk_forward_train``, which also pools the bilinear forms the parameter gradients need, and the
backward (``leafk_backward_saved``) runs no correlation; other geometries use ``leafk_forward`` + the generic FP32
``leafk_backward``.  Under ``torch.no_grad()`` the plain forward runs and nothing is saved.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _native as N

PCEN_FLOOR = 1e-12     # reference frontend.py:70
CLAMP_MIN = 1e-5       # reference frontend.py:84


@dataclass(frozen=True)
class LeafSpec:
    """Static geometry of one frontend (what reference frontend.py:38-39,65-75 fixes at init)."""
    F: int
    K: int
    H: int
    compression: bool = True
    algo: str = "auto"
    pcen_floor: float = PCEN_FLOOR
    clamp_min: float = CLAMP_MIN
    out_dtype: torch.dtype = torch.float32    # torch.bfloat16: K2 writes bf16 features (inference; LEAFK_OUTPUT_BF16)

    def config(self, input_dtype=torch.float32, reuse_banks: bool = False, out_dtype=None, prep=None) -> N.Config:
        algo = N.ALGOS[self.algo] | (N.REUSE_BANKS if reuse_banks else 0)
        out_dtype = self.out_dtype if out_dtype is None else out_dtype
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"out_dtype must be float32 or bfloat16, got {out_dtype}")
        cfg = N.Config(self.F, self.K, self.H, self.pcen_floor, self.clamp_min, int(self.compression),
                       algo, 1 if input_dtype == torch.int16 else 0, 1 if out_dtype == torch.bfloat16 else 0, None)
        if prep is not None:
            cfg.prep = C.pointer(prep.struct)            # `prep` (and through it the device arrays) must outlive the call
        return cfg

    def num_frames(self, T: int) -> int:
        lo = self.K // 2 + self.K % 2 - 1
        hi = self.K // 2
        return (T + lo + hi - self.K) // self.H + 1


class ClipPrep:
    """Per-clip preparation applied on the fly by the kernels (leafk_clip_prep): sample i of prepared clip b is
    ``raw[b, 0, i + start[b]]`` inside ``[0, length[b])`` -- outside, by ``pad_mode``: zero, the index wrapped around,
    the clip's edge sample, or ``pad_value[b]`` -- divided by ``divisor[b]``.  Built by :func:`prepare_clips`."""

    def __init__(self, n_samples: int, start=None, length=None, divisor=None, ld: int = 0, pad_mode: str = "zero",
                 pad_value=None):
        self.n_samples = int(n_samples)
        self.start, self.length, self.divisor, self.pad_value = start, length, divisor, pad_value
        p = lambda t: None if t is None else t.data_ptr()
        self.struct = N.ClipPrep(p(start), p(length), p(divisor), int(ld), N.PAD_MODES[pad_mode], p(pad_value))


def prepare_clips(spec: "LeafSpec", x_raw: torch.Tensor, n_samples: int, raw_lengths=None, starts="center",
                  pad_mode: str = "wrap", peak_normalize: bool = True, only_too_loud: bool = True) -> ClipPrep:
    """GPU side of the reference's per-clip input transforms: crop every raw clip to ``n_samples`` (``starts``:
    "center" = CenterCrop, or an int tensor of crop offsets = RandomCrop drawn by the caller), pad shorter clips
    (``pad_mode``: "edge" = what the reference's PadToSize(mode='wrap') does -- F.pad 'replicate' --, "min" = its
    PadToSize(mode='constant'), which pads with the clip's minimum, "wrap" = np.pad 'wrap' (PadToSize_NP), "zero" = the
    collate function's zero padding) and, optionally, peak-normalise clips whose peak exceeds 1
    (PeakNormalization(only_too_loud_sounds)).  ``x_raw`` (B,1,Traw) float32 or
    int16 on the GPU holds the raw clips row by row, ``raw_lengths`` (B,) their true lengths (default Traw).  Nothing is
    copied: the result only describes the view; one small kernel computes the peak divisors."""
    if not x_raw.is_cuda or x_raw.dim() != 3 or x_raw.shape[1] != 1 or not x_raw.is_contiguous():
        raise ValueError("x_raw must be a contiguous CUDA tensor of shape (B,1,Traw)")
    if pad_mode not in N.PAD_MODES:
        raise ValueError(f"pad_mode must be one of {sorted(N.PAD_MODES)}")
    B, _, Traw = x_raw.shape
    dev = x_raw.device
    length = torch.full((B,), Traw, dtype=torch.int32, device=dev) if raw_lengths is None else \
        torch.as_tensor(raw_lengths, dtype=torch.int32, device=dev).contiguous()
    if length.numel() != B or int(length.max()) > Traw or int(length.min()) < 1:
        raise ValueError("raw_lengths must hold B values in [1, Traw]")
    if isinstance(starts, str):
        if starts != "center":
            raise ValueError("starts must be 'center' or a tensor of crop offsets")
        # longer clips: (len - n)//2 (CenterCrop); shorter clips: -(n - len)//2, the offset PadToSize puts in front
        diff = length.to(torch.int64) - n_samples
        start = torch.where(diff >= 0, diff // 2, -((-diff) // 2)).to(torch.int32)
    else:
        start = torch.as_tensor(starts, dtype=torch.int32, device=dev).contiguous()
        if start.numel() != B:
            raise ValueError("starts must hold B crop offsets")
    prep = ClipPrep(n_samples, start, length, None, ld=Traw, pad_mode=pad_mode)
    padval = None
    with torch.cuda.device(dev):
        if pad_mode == "min":
            padval = torch.empty(B, dtype=torch.float32, device=dev)
            cfg = spec.config(x_raw.dtype, prep=prep)
            N.check(N.lib().leafk_clip_minimum(C.byref(cfg), _ptr(x_raw), B, int(n_samples), _ptr(padval), _stream_ptr(dev)),
                    "leafk_clip_minimum")
            prep = ClipPrep(n_samples, start, length, None, ld=Traw, pad_mode=pad_mode, pad_value=padval)
        if peak_normalize:
            div = torch.empty(B, dtype=torch.float32, device=dev)
            cfg = spec.config(x_raw.dtype, prep=prep)
            N.check(N.lib().leafk_peak_divisors(C.byref(cfg), _ptr(x_raw), B, int(n_samples), int(only_too_loud), _ptr(div),
                                                _stream_ptr(dev)), "leafk_peak_divisors")
            prep = ClipPrep(n_samples, start, length, div, ld=Traw, pad_mode=pad_mode, pad_value=padval)
    return prep


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_param(name: str, t: Optional[torch.Tensor], numel: int, device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.device != device:
        raise ValueError(f"parameter {name} lives on {t.device}, input on {device}")
    if t.dtype != torch.float32:
        raise TypeError(f"parameter {name} must be float32, got {t.dtype}")
    if t.numel() != numel:
        raise ValueError(f"parameter {name} has {t.numel()} elements, expected {numel}")
    return t if t.is_contiguous() else t.contiguous()     # only the address is used: no detach / new tensor needed


def _check_input(x: torch.Tensor) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError("input must be a torch.Tensor")
    if not x.is_cuda:
        raise N.LeafNativeError(
            "leaf_pytorch_b200 runs only on CUDA tensors (sm_100a kernels); there is no CPU fallback. "
            "Move the module and the input to a GPU.")
    if x.dtype not in (torch.float32, torch.int16):
        raise TypeError(f"input must be float32 (or int16 PCM, read as s/32768), got {x.dtype}")
    if x.dim() != 3 or x.shape[1] != 1:
        raise ValueError(f"input must have shape (B,1,T), got {tuple(x.shape)}")
    if x.shape[0] < 1 or x.shape[2] < 1:
        raise ValueError(f"empty input {tuple(x.shape)}")
    return x.contiguous()


def _params_struct(spec: LeafSpec, kernel, pool_w, pool_b, alpha, delta, root, ema_w, device):
    F = spec.F
    keep = [
        _check_param("kernel", kernel, 2 * F, device), _check_param("pool_w", pool_w, F, device),
        _check_param("pool_b", pool_b, F, device), _check_param("alpha", alpha, F, device),
        _check_param("delta", delta, F, device), _check_param("root", root, F, device),
        _check_param("ema_w", ema_w, F, device),
    ]
    return N.Params(*[None if t is None else t.data_ptr() for t in keep]), keep


def forward_raw(spec: LeafSpec, x, kernel, pool_w, pool_b, alpha, delta, root, ema_w,
                save_p: bool = False, prep: Optional[ClipPrep] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """One fused forward on the current stream; no autograd.  Returns (out, saved_p or None).  With ``prep`` the
    batch is ``prep.n_samples`` long per clip and ``x`` holds the raw rows (see prepare_clips)."""
    L = N.lib()
    x = _check_input(x)
    B, _, T = x.shape
    if prep is not None:
        T = prep.n_samples
    n = spec.num_frames(T)
    cfg = spec.config(x.dtype, prep=prep)
    prm, keep = _params_struct(spec, kernel, pool_w, pool_b, alpha, delta, root, ema_w, x.device)
    with torch.cuda.device(x.device):
        out = torch.empty((B, spec.F, n), dtype=spec.out_dtype, device=x.device)
        saved = torch.empty((B, spec.F, n), dtype=torch.float32, device=x.device) if save_p else None
        ws_bytes = L.leafk_workspace_bytes(C.byref(cfg), B, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        rc = L.leafk_forward(C.byref(cfg), C.byref(prm), _ptr(x), B, T, _ptr(out), _ptr(saved), _ptr(ws),
                             ws_bytes, _stream_ptr(x.device))
    N.check(rc, "leafk_forward")
    del keep
    return out, saved


def train_supported(spec: LeafSpec) -> bool:
    """True when the tensor-core training kernel covers this geometry (else training uses the generic FP32 backward)."""
    return spec.algo != "fp32" and bool(N.lib().leafk_train_supported(spec.F, spec.K, spec.H))


def forward_train_raw(spec: LeafSpec, x, kernel, pool_w, pool_b, alpha, delta, root, ema_w, prep: Optional[ClipPrep] = None):
    """Training forward on the current stream; no autograd.  Returns (out (B,F,N), saved (4,B,F,N)):
    saved[0] = floored pooled energies, saved[1:4] = pooled bilinear forms of the parameter gradients."""
    L = N.lib()
    x = _check_input(x)
    B, _, T = x.shape
    if prep is not None:
        T = prep.n_samples
    n = spec.num_frames(T)
    cfg = spec.config(x.dtype, out_dtype=torch.float32, prep=prep)
    prm, keep = _params_struct(spec, kernel, pool_w, pool_b, alpha, delta, root, ema_w, x.device)
    with torch.cuda.device(x.device):
        out = torch.empty((B, spec.F, n), dtype=torch.float32, device=x.device)
        saved = torch.empty((4, B, spec.F, n), dtype=torch.float32, device=x.device)
        ws_bytes = L.leafk_train_workspace_bytes(C.byref(cfg), B, T)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
        rc = L.leafk_forward_train(C.byref(cfg), C.byref(prm), _ptr(x), B, T, _ptr(out), _ptr(saved), _ptr(ws),
                                   ws_bytes, _stream_ptr(x.device))
    N.check(rc, "leafk_forward_train")
    del keep
    return out, saved


def workspace_bytes(spec: LeafSpec, B: int, n_frames: int, input_dtype=torch.float32) -> int:
    cfg = spec.config(input_dtype)
    return int(N.lib().leafk_workspace_bytes(C.byref(cfg), int(B), int(n_frames)))


def forward_window(spec: LeafSpec, x_win, T_total: int, t_off: int, n_begin: int, n_count: int,
                   kernel, pool_w, pool_b, alpha, delta, root, ema_w, ema_state=None,
                   out: Optional[torch.Tensor] = None, want_state: bool = True,
                   workspace: Optional[torch.Tensor] = None, reuse_banks: bool = False):
    """Frames [n_begin, n_begin+n_count) of clips of length T_total from a sample window
    (x_win[b,0,i] = sample t_off+i).  Returns (out (B,F,n_count), new ema state (B,F) or None).
    ``x_win`` may be a view into a longer buffer (unit stride along samples, any clip stride): no copy is made.
    ``workspace`` (uint8, at least workspace_bytes(spec, B, n_count)) lets consecutive chunks share one scratch
    buffer; with ``reuse_banks`` the bank prologue is skipped -- the caller guarantees that the previous call used the
    same workspace and the same parameters (LEAFK_REUSE_BANKS)."""
    L = N.lib()
    if isinstance(x_win, torch.Tensor) and x_win.is_cuda and x_win.dim() == 3 and x_win.shape[1] == 1 \
            and x_win.shape[0] >= 1 and x_win.shape[2] >= 1 and x_win.dtype in (torch.float32, torch.int16) \
            and x_win.stride(2) == 1 and (x_win.shape[0] == 1 or x_win.stride(0) >= x_win.shape[2]):
        ldx = x_win.stride(0) if x_win.shape[0] > 1 else x_win.shape[2]          # strided view: used in place
    else:
        x_win = _check_input(x_win)
        ldx = x_win.shape[2]
    B, _, T_win = x_win.shape
    prm, keep = _params_struct(spec, kernel, pool_w, pool_b, alpha, delta, root, ema_w, x_win.device)
    with torch.cuda.device(x_win.device):
        if out is None:
            out = torch.empty((B, spec.F, n_count), dtype=spec.out_dtype, device=x_win.device)
        if out.dtype not in (torch.float32, torch.bfloat16) or out.dim() != 3 or out.shape[0] != B \
                or out.shape[1] != spec.F or out.shape[2] != n_count or out.stride(2) != 1:
            raise ValueError("out must be a float32 / bfloat16 (B,F,n_count) view with unit stride along frames")
        cfg = spec.config(x_win.dtype, reuse_banks=reuse_banks, out_dtype=out.dtype)   # the format follows the buffer
        state_out = None
        if spec.compression and want_state:
            state_out = torch.empty((B, spec.F), dtype=torch.float32, device=x_win.device)
        if ema_state is not None:
            ema_state = _check_param("ema_state", ema_state, B * spec.F, x_win.device)
        ws_bytes = L.leafk_workspace_bytes(C.byref(cfg), B, n_count)
        if workspace is None:
            if reuse_banks:
                raise ValueError("reuse_banks needs the workspace of the previous call")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x_win.device)
        else:
            ws = workspace
            if ws.dtype != torch.uint8 or not ws.is_contiguous() or ws.device != x_win.device or ws.numel() < ws_bytes:
                raise ValueError(f"workspace must be a contiguous uint8 tensor of >= {ws_bytes} bytes on the input's device")
        rc = L.leafk_forward_window(C.byref(cfg), C.byref(prm), _ptr(x_win), B, int(ldx), int(T_total), int(t_off),
                                    T_win, int(n_begin), int(n_count), _ptr(ema_state), _ptr(state_out),
                                    _ptr(out), None, out.stride(0), out.stride(1), _ptr(ws), ws.numel(),
                                    _stream_ptr(x_win.device))
    N.check(rc, "leafk_forward_window")
    del keep
    return out, state_out


class _LeafFunction(torch.autograd.Function):
    """autograd node of reference frontend.py:78-89.  forward = leafk_forward_train (tensor-core geometries) or
    leafk_forward; backward = leafk_backward_saved or the generic leafk_backward; the gradient w.r.t. the waveform is
    computed only when the waveform requires grad."""

    @staticmethod
    def forward(ctx, spec: LeafSpec, x, kernel, pool_w, pool_b, alpha, delta, root, ema_w, prep=None):
        if spec.out_dtype != torch.float32:
            raise TypeError("training needs float32 features (out_dtype=torch.bfloat16 is an inference option)")
        if prep is not None and ctx.needs_input_grad[1]:
            raise TypeError("no gradient with respect to a waveform that is cropped / normalised on the fly")
        ctx.spec = spec
        ctx.prep = prep
        ctx.has_bias = pool_b is not None
        ctx.has_pcen = alpha is not None
        ctx.fused = train_supported(spec)
        if ctx.fused:
            out, saved = forward_train_raw(spec, x, kernel, pool_w, pool_b, alpha, delta, root, ema_w, prep=prep)
        else:
            out, saved = forward_raw(spec, x, kernel, pool_w, pool_b, alpha, delta, root, ema_w, save_p=True, prep=prep)
        # the waveform is needed again only by the generic backward and for its own gradient
        ctx.keep_x = (not ctx.fused) or ctx.needs_input_grad[1]
        tensors = [x if ctx.keep_x else None, saved, kernel, pool_w] + ([pool_b] if ctx.has_bias else []) + \
                  ([alpha, delta, root, ema_w] if ctx.has_pcen else [])
        ctx.save_for_backward(*tensors)
        ctx.x_shape, ctx.x_dtype = tuple(x.shape), x.dtype
        ctx.n_samples = x.shape[2] if prep is None else prep.n_samples
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        L = N.lib()
        spec: LeafSpec = ctx.spec
        saved = list(ctx.saved_tensors)
        x, p = saved[0], saved[1]
        kernel, pool_w = saved[2], saved[3]
        idx = 4
        pool_b = None
        if ctx.has_bias:
            pool_b = saved[idx]; idx += 1
        alpha = delta = root = ema_w = None
        if ctx.has_pcen:
            alpha, delta, root, ema_w = saved[idx:idx + 4]
        B, T = ctx.x_shape[0], ctx.n_samples
        if x is not None:
            x = _check_input(x)
        cfg = spec.config(ctx.x_dtype, out_dtype=torch.float32, prep=ctx.prep)
        dev = p.device
        prm, keep = _params_struct(spec, kernel, pool_w, pool_b, alpha, delta, root, ema_w, dev)
        grad_out = grad_out.contiguous().to(torch.float32)
        want_gx = ctx.needs_input_grad[1]
        if want_gx and ctx.x_dtype != torch.float32:
            raise TypeError("no gradient with respect to an int16 waveform")
        with torch.cuda.device(dev):
            def z(t):                                    # the backward writes (does not accumulate) every entry
                return None if t is None else torch.empty(t.numel(), dtype=torch.float32, device=dev)
            g = [z(kernel), z(pool_w), z(pool_b), z(alpha), z(delta), z(root), z(ema_w)]
            grads = N.Grads(*[None if t is None else t.data_ptr() for t in g])
            gx = torch.empty(ctx.x_shape, dtype=torch.float32, device=dev) if want_gx else None
            if ctx.fused:
                ws_bytes = L.leafk_backward_saved_workspace_bytes(C.byref(cfg), B, T, int(want_gx))
                ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
                rc = L.leafk_backward_saved(C.byref(cfg), C.byref(prm), _ptr(x), B, T, _ptr(grad_out), _ptr(p),
                                            C.byref(grads), _ptr(gx), _ptr(ws), ws_bytes, _stream_ptr(dev))
                what = "leafk_backward_saved"
            else:
                ws_bytes = L.leafk_backward_workspace_bytes(C.byref(cfg), B, T)
                ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
                rc = L.leafk_backward(C.byref(cfg), C.byref(prm), _ptr(x), B, T, _ptr(grad_out), _ptr(p),
                                      C.byref(grads), _ptr(gx), _ptr(ws), ws_bytes, _stream_ptr(dev))
                what = "leafk_backward"
        N.check(rc, what)
        del keep

        def shaped(gt, like):
            return None if gt is None else gt.view(like.shape)
        return (None, gx, shaped(g[0], kernel), shaped(g[1], pool_w), shaped(g[2], pool_b) if pool_b is not None else None,
                shaped(g[3], alpha) if alpha is not None else None, shaped(g[4], delta) if delta is not None else None,
                shaped(g[5], root) if root is not None else None, shaped(g[6], ema_w) if ema_w is not None else None, None)


def leaf_forward(spec: LeafSpec, x, kernel, pool_w, pool_b=None, alpha=None, delta=None, root=None, ema_w=None,
                 prep: Optional[ClipPrep] = None):
    """Differentiable fused LEAF forward: (B,1,T) float32 CUDA waveform -> (B,F,N).  ``prep``: per-clip crop / pad /
    peak normalisation applied on the fly (prepare_clips); ``x`` then holds the raw clips."""
    if spec.compression and any(t is None for t in (alpha, delta, root, ema_w)):
        raise ValueError("compression=True needs alpha, delta, root and ema_w")
    tensors = (x, kernel, pool_w, pool_b, alpha, delta, root, ema_w)
    if not torch.is_grad_enabled() or not any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        # inference: nothing is saved for a backward that will not come
        return forward_raw(spec, x, *[None if t is None else t.detach() for t in tensors[1:]], prep=prep)[0]
    return _LeafFunction.apply(spec, x, kernel, pool_w, pool_b, alpha, delta, root, ema_w, prep)


class _PreEmphasis(torch.autograd.Function):
    """y[b,0,t] = w[0] x[b,0,t] + w[1] x[b,0,t+1]  (x[T] = 0): leafk_preemp_forward / leafk_preemp_backward."""

    @staticmethod
    def forward(ctx, x, w):
        L = N.lib()
        x = _check_input(x)
        if x.dtype != torch.float32:
            raise TypeError("pre-emphasis needs a float32 waveform")
        w2 = _check_param("preemp weight", w, 2, x.device)
        B, _, T = x.shape
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            N.check(L.leafk_preemp_forward(_ptr(x), _ptr(w2), B, T, _ptr(y), _stream_ptr(x.device)), "leafk_preemp_forward")
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        L = N.lib()
        x, w = ctx.saved_tensors
        w2 = _check_param("preemp weight", w, 2, x.device)
        B, _, T = x.shape
        gy = gy.contiguous().to(torch.float32)
        with torch.cuda.device(x.device):
            gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            gw = torch.empty(2, dtype=torch.float32, device=x.device)
            nws = int(L.leafk_preemp_backward_workspace_bytes())
            ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
            N.check(L.leafk_preemp_backward(_ptr(x), _ptr(w2), _ptr(gy), B, T, _ptr(gx), _ptr(gw), _ptr(ws), nws,
                                            _stream_ptr(x.device)), "leafk_preemp_backward")
        return gx, gw.view(w.shape)


def pre_emphasis(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Learnable 2-tap pre-emphasis in front of the Gabor bank (original LEAF; reference frontend.py:40-41 stub)."""
    return _PreEmphasis.apply(x, weight)


class _InstanceNorm(torch.autograd.Function):
    """Per-(clip, filter) mean / variance normalisation over frames: leafk_instnorm_forward / _backward."""

    @staticmethod
    def forward(ctx, v, eps):
        L = N.lib()
        if not v.is_cuda or v.dtype != torch.float32 or v.dim() != 3:
            raise TypeError("instance norm expects float32 CUDA features of shape (B,F,N)")
        v = v.contiguous()
        rows, n = v.shape[0] * v.shape[1], v.shape[2]
        out = torch.empty_like(v)
        stats = torch.empty((rows, 2), dtype=torch.float32, device=v.device)
        with torch.cuda.device(v.device):
            N.check(L.leafk_instnorm_forward(_ptr(v), rows, n, float(eps), _ptr(out), _ptr(stats), _stream_ptr(v.device)),
                    "leafk_instnorm_forward")
        ctx.save_for_backward(v, stats)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        L = N.lib()
        v, stats = ctx.saved_tensors
        rows, n = v.shape[0] * v.shape[1], v.shape[2]
        g = g.contiguous().to(torch.float32)
        gv = torch.empty_like(v)
        with torch.cuda.device(v.device):
            N.check(L.leafk_instnorm_backward(_ptr(v), _ptr(stats), _ptr(g), rows, n, _ptr(gv), _stream_ptr(v.device)),
                    "leafk_instnorm_backward")
        return gv, None


def instance_norm(v: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """mean_var_norm of the original LEAF (reference frontend.py:62-63 stub): nn.InstanceNorm1d without affine."""
    return _InstanceNorm.apply(v, eps)


_host_cache = {}                 # (device, shapes, geometry, algo, dtype) -> device scratch of forward_host
_host_cache_lock = threading.Lock()
_HOST_CACHE_MAX = 8


def forward_host(spec: LeafSpec, x_host: torch.Tensor, kernel, pool_w, pool_b, alpha, delta, root, ema_w,
                 out_host: Optional[torch.Tensor] = None, n_slices: int = 8, device=None) -> torch.Tensor:
    """End-to-end call on HOST buffers through leafk_forward_host: the H2D copy is issued in
    ``n_slices`` pieces on a side stream, each followed by a stream-ordered flag write, and ONE
    persistent launch of the tensor-core kernel consumes clips as their slice lands (the FP32 kernel
    falls back to per-slice launches).  ``x_host``
    (B,1,T) float32 or int16 PCM, CPU (pinned for full speed); returns ``out_host`` (B,F,N) pinned CPU, float32 or
    -- when ``spec.out_dtype`` is bfloat16 -- bf16 written by the PCEN kernel (half the read-back).  The call
    synchronises the compute stream before returning (the result is on the host)."""
    L = N.lib()
    if x_host.is_cuda or x_host.dtype not in (torch.float32, torch.int16) or x_host.dim() != 3 or x_host.shape[1] != 1:
        raise ValueError("x_host must be a float32 (or int16 PCM) CPU tensor of shape (B,1,T)")
    device = torch.device(device if device is not None else kernel.device)
    if device.type != "cuda":
        raise N.LeafNativeError("forward_host needs the parameters on a CUDA device; there is no CPU fallback")
    x_host = x_host.contiguous()
    B, _, T = x_host.shape
    n = spec.num_frames(T)
    odt = spec.out_dtype
    cfg = spec.config(x_host.dtype, out_dtype=odt)
    prm, keep = _params_struct(spec, kernel, pool_w, pool_b, alpha, delta, root, ema_w, device)
    if out_host is None:
        out_host = torch.empty((B, spec.F, n), dtype=odt, pin_memory=True)
    if out_host.is_cuda or out_host.dtype != odt or tuple(out_host.shape) != (B, spec.F, n) \
            or not out_host.is_contiguous():
        raise ValueError(f"out_host must be a contiguous {odt} CPU tensor of shape (B,F,N)")
    with torch.cuda.device(device):
        ws_bytes = L.leafk_workspace_bytes(C.byref(cfg), B, n)
        key = (device.index, B, T, spec.F, spec.K, spec.H, spec.algo, n, x_host.dtype, odt)
        with _host_cache_lock:
            bufs = _host_cache.get(key)
            if bufs is None or bufs[2].numel() < ws_bytes:
                while len(_host_cache) >= _HOST_CACHE_MAX:          # bounded: drop the oldest shape, keep the rest
                    _host_cache.pop(next(iter(_host_cache)))
                bufs = (torch.empty(B * T, dtype=x_host.dtype, device=device),
                        torch.empty(B * spec.F * n, dtype=odt, device=device),
                        torch.empty(ws_bytes, dtype=torch.uint8, device=device), torch.cuda.Stream(device=device),
                        torch.zeros(1, dtype=torch.int32).pin_memory())
                _host_cache[key] = bufs
        dev_x, dev_out, ws, side, status = bufs
        cur = torch.cuda.current_stream(device)
        side.wait_stream(cur)
        rc = L.leafk_forward_host(C.byref(cfg), C.byref(prm), C.c_void_p(x_host.data_ptr()), B, T,
                                  C.c_void_p(out_host.data_ptr()), int(n_slices), _ptr(dev_x), _ptr(dev_out),
                                  _ptr(ws), ws.numel(), C.c_void_p(cur.cuda_stream), C.c_void_p(side.cuda_stream),
                                  C.c_void_p(status.data_ptr()))
        N.check(rc, "leafk_forward_host")
        cur.synchronize()
        # a stalled H2D slice is reported here (the status word came back with the features) instead of a trap in the kernel
        N.check(L.leafk_status_message(int(status[0])), "leafk_forward_host")
    del keep
    return out_host


def k1_clock_probe(spec: LeafSpec, x, kernel, pool_w, pool_b, alpha, delta, root, ema_w):
    """Run one forward with profiling on and return (SM cycles, nanoseconds) of the tensor-core kernel's CTA 0:
    cycles/ns is the SM clock (GHz) the kernel actually ran at."""
    L = N.lib()
    x = _check_input(x)
    B, _, T = x.shape
    n = spec.num_frames(T)
    cfg = spec.config(x.dtype)
    prm, keep = _params_struct(spec, kernel, pool_w, pool_b, alpha, delta, root, ema_w, x.device)
    with torch.cuda.device(x.device):
        out = torch.empty((B, spec.F, n), dtype=torch.float32, device=x.device)
        ws_bytes = L.leafk_workspace_bytes(C.byref(cfg), B, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        L.leafk_profile_begin()
        rc = L.leafk_forward(C.byref(cfg), C.byref(prm), _ptr(x), B, T, _ptr(out), None, _ptr(ws), ws_bytes,
                             _stream_ptr(x.device))
        torch.cuda.synchronize(x.device)
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        L.leafk_profile_end(C.byref(a), C.byref(b), C.byref(c))
        N.check(rc, "leafk_forward")
        cyc, ns = C.c_longlong(0), C.c_longlong(0)
        N.check(L.leafk_profile_k1_clock(C.byref(cfg), B, T, _ptr(ws), ws_bytes, C.byref(cyc), C.byref(ns)),
                "leafk_profile_k1_clock")
    del keep
    return int(cyc.value), int(ns.value)



def tc_schedule(spec: LeafSpec, x, kernel, pool_w, pool_b, alpha, delta, root, ema_w):
    """Run one forward and read back the support-pruning schedule of the tensor-core kernel (profiling / tests).
    Returns dict(n_groups, channels_per_group, n_ksteps, active=[per group: channels running per k-step],
    active_all_products=[per group: channels running all three split products per k-step],
    executed_fraction = executed / unpruned tensor work)."""
    L = N.lib()
    x = _check_input(x)
    B, _, T = x.shape
    n = spec.num_frames(T)
    cfg = spec.config(x.dtype)
    prm, keep = _params_struct(spec, kernel, pool_w, pool_b, alpha, delta, root, ema_w, x.device)
    with torch.cuda.device(x.device):
        out = torch.empty((B, spec.F, n), dtype=torch.float32, device=x.device)
        ws_bytes = L.leafk_workspace_bytes(C.byref(cfg), B, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        N.check(L.leafk_forward(C.byref(cfg), C.byref(prm), _ptr(x), B, T, _ptr(out), None, _ptr(ws), ws_bytes,
                                _stream_ptr(x.device)), "leafk_forward")
        torch.cuda.synchronize(x.device)
        ng, cg, ks = C.c_int(0), C.c_int(0), C.c_int(0)
        codes = (C.c_int * (64 * 128))()
        N.check(L.leafk_profile_tc_schedule(C.byref(cfg), B, T, _ptr(ws), ws_bytes, C.byref(ng), C.byref(cg), C.byref(ks),
                                            codes, 64 * 128), "leafk_profile_tc_schedule")
    del keep

    def table(g, shift):
        return [16 * ((codes[g * ks.value + s] >> shift) & 15) for s in range(ks.value)]
    active = [table(g, 0) for g in range(ng.value)]          # channels that run (main product)
    active3 = [table(g, 4) for g in range(ng.value)]         # channels that run all three products
    total = sum(sum(a) for a in active) + 2 * sum(sum(a) for a in active3)
    return {"n_groups": ng.value, "channels_per_group": cg.value, "n_ksteps": ks.value, "active": active,
            "active_all_products": active3,
            "executed_fraction": total / float(3 * ng.value * cg.value * ks.value)}


def profile_begin() -> None:
    N.lib().leafk_profile_begin()


def profile_end():
    """-> (n_forwards, mean ms of K0, K1, K2)"""
    a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
    n = N.lib().leafk_profile_end(C.byref(a), C.byref(b), C.byref(c))
    return int(n), float(a.value), float(b.value), float(c.value)


def describe_plan(F: int, K: int, H: int) -> dict:
    """Host-only: how the tensor-core kernels would run this geometry (leafk_describe_plan)."""
    a, b, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    N.check(N.lib().leafk_describe_plan(int(F), int(K), int(H), C.byref(a), C.byref(b), C.byref(c), C.byref(d)),
            "leafk_describe_plan")
    return {"forward_groups": a.value, "forward_channels_per_group": b.value, "train_filters_per_group": c.value,
            "frame_slots": d.value}


def tc_supported(F: int, K: int, H: int) -> bool:
    """True when the tcgen05 kernel covers this geometry (else algo="auto" uses the fp32 kernel)."""
    return bool(N.lib().leafk_tc_supported(int(F), int(K), int(H)))


def launch_count(reset: bool = False) -> int:
    """Kernels launched through libleafk.so by this thread (bench.py's gpu_launches)."""
    return int(N.lib().leafk_launch_count(int(reset)))
