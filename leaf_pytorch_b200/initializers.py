"""Initial (centre, width) of the Gabor filters from a mel filterbank.

Restates the reference's one-time initialisation (initializers.py:7-18 -> filters.py:27-57):
triangular mel filters on a 512-point FFT grid, square-rooted; the peak bin gives the centre
frequency and the count of bins at or above half the peak gives the full width at half maximum,
which maps to a Gaussian width in samples.  The arithmetic (dtypes and operation order) is kept so
that the resulting ``_kernel`` parameter is bit-identical to the reference's -- checked against
``init_kernel`` in the golden vectors by tests/test_boundary.py.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def same_padding(kernel_size: int) -> Tuple[int, int]:
    """(left, right) zeros of a 'same' correlation (reference utils.py:5-10)."""
    half = kernel_size // 2
    return half + kernel_size % 2 - 1, half


def mel_gabor_parameters(n_filters: int, sample_rate: int = 16000, min_freq: float = 60.0,
                         max_freq: float = 7800.0, n_fft: int = 512) -> torch.Tensor:
    """(n_filters, 2) float32: column 0 centre frequency in rad/sample, column 1 width in samples."""
    import torchaudio  # init-time only dependency, as in the reference (filters.py:4,48)

    bank = torchaudio.functional.melscale_fbanks(n_freqs=n_fft // 2 + 1, f_min=min_freq, f_max=max_freq,
                                                 n_mels=n_filters, sample_rate=sample_rate).transpose(1, 0)
    amp = torch.sqrt(bank)                                   # (F, n_fft/2+1)
    peak_bin = torch.argmax(amp, dim=1)
    peak = torch.max(amp, dim=1, keepdim=True).values
    fwhm_bins = torch.sum((amp >= peak / 2.).float(), dim=1)
    scale = torch.sqrt(2. * torch.log(torch.tensor(2.))) * n_fft
    centre = peak_bin * 2 * math.pi / n_fft
    width = scale / (math.pi * fwhm_bins)
    return torch.stack([centre, width], dim=1)


class MelGaborInit:
    """Callable initializer: ``MelGaborInit(...)((F, 2)) -> (F,2) tensor`` (reference GaborInit,
    initializers.py:7-18; like the reference it always uses a 512-point grid, filters.py:17)."""

    def __init__(self, sample_rate: int = 16000, min_freq: float = 60.0, max_freq: float = 7800.0, **_unused):
        self.sample_rate = sample_rate
        self.min_freq = min_freq
        self.max_freq = max_freq

    def __call__(self, shape, dtype=None) -> torch.Tensor:
        if len(shape) != 2:
            raise NotImplementedError("only (n_filters, 2) parameter shapes are supported")
        return mel_gabor_parameters(shape[0], self.sample_rate, self.min_freq, self.max_freq)


# name used by the reference (initializers.py:7)
GaborInit = MelGaborInit
