"""Caller glue: the step right after the hot path.

The reference wraps frontend and backbone as ``Classifier`` (reference models/classifier.py:7-18): ``features(x)`` ->
``unsqueeze(1)`` -> 2-D CNN.  This module offers the same wrapper for any backbone ``nn.Module`` taking (B,1,F,N):
the frontend emits that shape directly (``out_layout="b1fn"``, a view), optionally as bf16 for a bf16 backbone, and in
eval mode frontend + backbone can be captured into ONE CUDA graph, so a serving step is a single graph launch.  To run
the reference's own ``Classifier`` unchanged on this frontend, see ``integration.install()``."""
from __future__ import annotations

import torch
from torch import nn

from .frontend import Leaf


class Classifier(nn.Module):
    """``features`` (a Leaf) followed by ``model`` (any backbone on (B,1,F,N)); attribute names as in the reference."""

    def __init__(self, features: Leaf, model: nn.Module, feature_dtype: torch.dtype = torch.float32):
        super().__init__()
        self.features = features
        self.model = model
        self.feature_dtype = feature_dtype

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        fe = self.features
        keep = (fe.out_layout, fe.out_dtype)
        # bf16 features are written by the PCEN kernel itself (inference); training keeps float32
        fe.out_layout = "b1fn"
        fe.out_dtype = self.feature_dtype if not (self.training and torch.is_grad_enabled()) else torch.float32
        try:
            out = fe(x)
        finally:
            fe.out_layout, fe.out_dtype = keep
        if out.dtype != self.feature_dtype and self.feature_dtype != torch.float32:
            out = out.to(self.feature_dtype)
        return self.model(out)

    @torch.no_grad()
    def capture(self, example: torch.Tensor) -> "GraphedClassifier":
        """Eval-mode CUDA graph of frontend + backbone for inputs shaped like ``example`` (device resident)."""
        if self.training:
            raise RuntimeError("capture() is an inference feature: call .eval() first")
        return GraphedClassifier(self, example)


class GraphedClassifier:
    """One CUDA-graph launch per batch: static input / output buffers, ``__call__`` copies in and replays."""

    def __init__(self, clf: Classifier, example: torch.Tensor):
        self.static_in = example.clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side):                      # warm-up outside the capture (lazy initialisation, autotune)
            for _ in range(3):
                clf(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = clf(self.static_in)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        self.static_in.copy_(x)
        self.graph.replay()
        return self.static_out
