"""Drop-in ``Leaf`` module backed by the sm_100a kernels in libleafk.so.

Mirrors the public surface of the reference module (reference leaf_pytorch/frontend.py:22-89):
same constructor keywords and defaults, same sub-module / parameter names and shapes (so
``state_dict`` round-trips with reference checkpoints, reference train.py:36,94 and
frontend_helper.py:53), same exceptions for the options the reference declares but does not
implement.  What differs is the inside: the sub-modules here only *hold* parameters; the whole
forward (constraint, filter synthesis, correlation, modulus, pooling, floor, PCEN) is one call into
the fused kernels, and so is the backward.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Union

import torch
from torch import nn

from . import functional as LF
from .initializers import MelGaborInit, same_padding


class _FusedStage(nn.Module):
    """Parameter holder for one stage of the fused pipeline; not callable on its own."""

    def forward(self, *_args, **_kw):
        raise RuntimeError(
            f"{type(self).__name__} is fused into Leaf.forward on the GPU (libleafk.so) and cannot be "
            "called stand-alone; call the Leaf module.")


class GaborConstraint(_FusedStage):
    """Bounds applied to (centre, width) inside the kernel prologue (reference convolution.py:10-22)."""

    def __init__(self, kernel_size: int):
        super().__init__()
        self._kernel_size = kernel_size

    def bounds(self):
        root_2ln2 = math.sqrt(2.0 * math.log(2.0))
        return (0.0, math.pi), (4 * root_2ln2 / math.pi, self._kernel_size * root_2ln2 / math.pi)


class GaborConv1d(_FusedStage):
    """Holds ``_kernel`` (F,2): per-filter centre frequency (rad/sample) and Gaussian width (samples).
    Attribute names follow reference convolution.py:25-69."""

    def __init__(self, filters: int, kernel_size: int, strides: int = 1, padding: str = "same",
                 initializer: Union[str, Callable, None] = None, use_bias: bool = False,
                 sort_filters: bool = False, use_legacy_complex: bool = False):
        super().__init__()
        self._filters = filters // 2
        self._kernel_size = kernel_size
        self._strides = strides
        self._padding = padding
        self._use_bias = use_bias
        # sort_filters (reference convolution.py:74-75 raises NotImplementedError): as in the original LEAF, the rows of
        # the constrained kernel are ordered by centre frequency before the bank is synthesised
        self._sort_filters = sort_filters
        if use_bias:
            raise NotImplementedError("the fused kernel has no per-channel conv bias (Leaf never enables it)")
        shape = (self._filters, 2)
        if callable(initializer):
            start = initializer(shape)
        elif initializer == "random":
            start = torch.randn(*shape)
        elif initializer == "xavier_normal":
            start = nn.init.xavier_normal_(torch.randn(*shape))
        elif initializer == "kaiming_normal":
            start = nn.init.kaiming_normal_(torch.randn(*shape))
        else:
            raise ValueError("unsupported initializer")
        self.constraint = GaborConstraint(kernel_size)
        self._kernel = nn.Parameter(torch.as_tensor(start, dtype=torch.float32).clone())
        self._pad_value = same_padding(kernel_size) if padding.lower() == "same" else padding
        self._bias = None
        # both complex formulations of the reference give the same filters (<=4e-9); the kernel
        # has a single synthesis path, the flag is kept for config compatibility only
        self.use_legacy_complex = use_legacy_complex


class PreEmphasis(_FusedStage):
    """Learnable 2-tap pre-emphasis (original LEAF; the reference declares ``preemp`` but raises, frontend.py:40-41).
    ``weight`` has the shape of an nn.Conv1d(1, 1, 2, bias=False) kernel and starts at (-0.97, 1)."""

    def __init__(self, coeff: float = 0.97):
        super().__init__()
        self.weight = nn.Parameter(torch.tensor([[[-coeff, 1.0]]], dtype=torch.float32))


class InstanceNorm(_FusedStage):
    """mean_var_norm (original LEAF; the reference raises, frontend.py:62-63): every (clip, filter) row is normalised
    over its frames, no affine parameters (nn.InstanceNorm1d semantics, biased variance)."""

    def __init__(self, eps: float = 1e-5):
        super().__init__()
        self.eps = eps


class SquaredModulus(_FusedStage):
    """re^2 + im^2 of each filter pair, done in registers in the conv epilogue (reference frontend.py:10-19)."""


class GaussianLowPass(_FusedStage):
    """Holds the per-filter pooling width ``weights`` (1,1,F,1) and ``_bias`` (F,)
    (reference pooling.py:8-29)."""

    def __init__(self, in_channels: int, kernel_size: int, strides: int = 1, padding: str = "same",
                 use_bias: bool = True):
        super().__init__()
        self.kernel_size = kernel_size
        self.strides = strides
        self.padding = padding
        self.use_bias = use_bias
        self.in_channels = in_channels
        self.weights = nn.Parameter(torch.full((1, 1, in_channels, 1), 0.4))
        self._bias = nn.Parameter(torch.ones(in_channels)) if use_bias else None
        self.pad_value = same_padding(kernel_size) if padding.lower() == "same" else padding


class ExponentialMovingAverage(_FusedStage):
    """Holds the smoothing coefficient(s) of the PCEN IIR (reference postprocessing.py:5-11)."""

    def __init__(self, in_channels: int, coeff_init: float, per_channel: bool = False):
        super().__init__()
        self._coeff_init = coeff_init
        self._per_channel = per_channel
        n = in_channels if per_channel else 1
        self._weights = nn.Parameter(torch.ones(n) * coeff_init)


class PCENLayer(_FusedStage):
    """Holds alpha / delta / root and the smoother (reference postprocessing.py:31-60)."""

    def __init__(self, in_channels: int, alpha: float = 0.96, smooth_coef: float = 0.04, delta: float = 2.0,
                 root: float = 2.0, floor: float = 1e-6, trainable: bool = False,
                 learn_smooth_coef: bool = False, per_channel_smooth_coef: bool = False):
        super().__init__()
        self._alpha_init, self._delta_init, self._root_init = alpha, delta, root
        self._smooth_coef = smooth_coef
        self._floor = floor
        self._trainable = trainable
        self._learn_smooth_coef = learn_smooth_coef
        self._per_channel_smooth_coef = per_channel_smooth_coef
        self.alpha = nn.Parameter(torch.ones(in_channels) * alpha)
        self.delta = nn.Parameter(torch.ones(in_channels) * delta)
        self.root = nn.Parameter(torch.ones(in_channels) * root)
        if not learn_smooth_coef:
            raise ValueError("SimpleRNN based ema not implemented.")
        self.ema = ExponentialMovingAverage(in_channels, coeff_init=smooth_coef, per_channel=per_channel_smooth_coef)


class Leaf(nn.Module):
    """LEAF frontend: (B,1,T) float32 waveform on a B200 -> (B,n_filters,N) features.

    Same signature as the reference (frontend.py:23-36).  Extra keyword ``algo`` selects the
    correlation kernel: "auto" (tensor cores when the geometry allows), "tc", "tc_full" (no support pruning) or
    "fp32".  ``out_dtype=torch.bfloat16`` makes the PCEN kernel write bf16 features (inference only) and
    ``out_layout="b1fn"`` returns them as (B,1,F,N), the shape the reference's Classifier feeds its 2-D backbone
    (reference models/classifier.py:15-17) -- a view, no copy.
    """

    def __init__(self, n_filters: int = 40, sample_rate: int = 16000, window_len: float = 25.,
                 window_stride: float = 10., preemp: bool = False, init_min_freq=60.0, init_max_freq=7800.0,
                 mean_var_norm: bool = False, pcen_compression: bool = True, use_legacy_complex=False,
                 initializer="default", algo: str = "auto", out_dtype: torch.dtype = torch.float32,
                 out_layout: str = "bfn", sort_filters: bool = False):
        super().__init__()
        window_size = int(sample_rate * window_len // 1000 + 1)
        hop = int(sample_rate * window_stride // 1000)
        # preemp / mean_var_norm: the reference raises NotImplementedError for both; here they follow the original LEAF
        self._preemp = PreEmphasis() if preemp else None
        if initializer == "default":
            initializer = MelGaborInit(sample_rate=sample_rate, min_freq=init_min_freq, max_freq=init_max_freq)
        self._complex_conv = GaborConv1d(filters=2 * n_filters, kernel_size=window_size, strides=1, padding="same",
                                         use_bias=False, initializer=initializer, sort_filters=sort_filters,
                                         use_legacy_complex=use_legacy_complex)
        self._activation = SquaredModulus()
        self._pooling = GaussianLowPass(n_filters, kernel_size=window_size, strides=hop, padding="same")
        self._instance_norm = InstanceNorm() if mean_var_norm else None
        if pcen_compression:
            self._compression = PCENLayer(n_filters, alpha=0.96, smooth_coef=0.04, delta=2.0, floor=1e-12,
                                          trainable=True, learn_smooth_coef=True, per_channel_smooth_coef=True)
        else:
            self._compression = None
        self._maximum_val = torch.tensor(1e-5)
        if out_layout not in ("bfn", "b1fn"):
            raise ValueError("out_layout must be 'bfn' (reference) or 'b1fn'")
        self.algo = algo
        self.out_dtype = out_dtype
        self.out_layout = out_layout
        self._spec = LF.LeafSpec(F=n_filters, K=window_size, H=hop, compression=bool(pcen_compression), algo=algo,
                                 pcen_floor=1e-12, clamp_min=1e-5, out_dtype=out_dtype)

    # ------------------------------------------------------------------ helpers
    @property
    def spec(self) -> LF.LeafSpec:
        if self._spec.algo != self.algo or self._spec.out_dtype != self.out_dtype:
            self._spec = LF.LeafSpec(**{**self._spec.__dict__, "algo": self.algo, "out_dtype": self.out_dtype})
        return self._spec

    def num_frames(self, n_samples: int) -> int:
        return self.spec.num_frames(n_samples)

    def _require_plain(self, what: str) -> None:
        """The host-buffer, chunked and streaming entry points run the fused path only: with the optional stages around
        it (pre-emphasis in front, mean/variance normalisation over all frames behind) they would silently compute
        something else, so they refuse."""
        if self._preemp is not None or self._instance_norm is not None:
            raise NotImplementedError(f"{what} does not run the optional pre-emphasis / mean_var_norm stages; "
                                      "use Leaf.forward for a module built with preemp=True or mean_var_norm=True")

    def _param_tuple(self):
        pc = self._compression
        kernel = self._complex_conv._kernel
        if self._complex_conv._sort_filters:
            # order by the constrained centre frequency (original LEAF: argsort + gather on the clamped kernel)
            order = torch.argsort(kernel.detach()[:, 0].clamp(0.0, math.pi), stable=True)
            kernel = kernel[order]
        return (kernel, self._pooling.weights, self._pooling._bias,
                None if pc is None else pc.alpha, None if pc is None else pc.delta,
                None if pc is None else pc.root, None if pc is None else pc.ema._weights)

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """reference frontend.py:78-89, fused."""
        if isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 3 and x.shape[0] == 0 and x.shape[2] > 0:
            # empty batch: the reference returns an empty (0,F,N) tensor (conv1d accepts B = 0)
            out = torch.zeros((0, self.spec.F, self.spec.num_frames(x.shape[2])), dtype=self.out_dtype, device=x.device)
        else:
            if self._preemp is not None:
                x = LF.pre_emphasis(x, self._preemp.weight)
            out = LF.leaf_forward(self.spec, x, *self._param_tuple())
            if self._instance_norm is not None:
                out = LF.instance_norm(out, self._instance_norm.eps)
        return out.unsqueeze(1) if self.out_layout == "b1fn" else out

    def forward_prepared(self, x_raw: torch.Tensor, n_samples: int, raw_lengths=None, starts="center",
                         pad_mode: str = "edge", peak_normalize: bool = True) -> torch.Tensor:
        """Features of raw, unequal-length clips without a prepared copy of the batch: every clip is cropped
        (``starts`` "center" or per-clip offsets) or padded (``pad_mode`` "edge" / "min" / "wrap" / "zero") to ``n_samples`` and, when
        its peak exceeds 1, peak-normalised -- the reference's per-clip transforms (utilities/data/raw_transforms.py:
        121-160, 334-344; utilities/data/utils.py:8-28) -- inside the kernels' own staging of the waveform.
        ``x_raw`` (B,1,Traw) float32 or int16 PCM on the GPU, ``raw_lengths`` (B,) true lengths."""
        if self._preemp is not None:
            # the clips are staged (cropped / padded / normalised) inside the Gabor kernel; a pre-emphasis belongs
            # between the two, so it needs the prepared batch
            raise NotImplementedError("forward_prepared cannot apply preemp=True on the fly; prepare the batch and call forward")
        prep = LF.prepare_clips(self.spec, x_raw, n_samples, raw_lengths, starts, pad_mode, peak_normalize)
        out = LF.leaf_forward(self.spec, x_raw, *self._param_tuple(), prep=prep)
        if self._instance_norm is not None:
            out = LF.instance_norm(out, self._instance_norm.eps)
        return out.unsqueeze(1) if self.out_layout == "b1fn" else out

    def forward_host(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
                     n_slices: int = 8) -> torch.Tensor:
        """Inference on host buffers: pinned (B,1,T) in -> pinned (B,F,N) out, H2D / kernels / D2H
        pipelined over batch slices inside the library (leafk_forward_host).  No autograd."""
        self._require_plain("forward_host")
        prm = [None if p is None else p.detach() for p in self._param_tuple()]
        return LF.forward_host(self.spec, x_host, *prm, out_host=out_host, n_slices=n_slices,
                               device=self._complex_conv._kernel.device)

    def extra_repr(self) -> str:
        s = self._spec
        return f"n_filters={s.F}, taps={s.K}, hop={s.H}, pcen={s.compression}, algo={self.algo}"
