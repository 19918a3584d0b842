"""``get_frontend(cfg)``: build the frontend from the reference's config dictionary
(keys as read by reference frontend_helper.py:7-54 from cfgs/**.cfg)."""
from __future__ import annotations

import os

import torch

from .frontend import Leaf


def get_frontend(opt, algo: str = "auto") -> Leaf:
    fe_cfg = opt["frontend"]
    audio_cfg = opt["audio_config"]
    if "leaf" not in fe_cfg["name"].lower():
        raise NotImplementedError("Other front ends not implemented yet.")
    common = dict(use_legacy_complex=fe_cfg.get("use_legacy_complex", False),
                  initializer=fe_cfg.get("initializer", "default"), algo=algo)
    if fe_cfg.get("default_args", False):
        fe = Leaf(**common)
    else:
        fe = Leaf(n_filters=int(fe_cfg.get("n_filters", 40.0)),
                  sample_rate=int(audio_cfg.get("sample_rate", 16000)),
                  window_len=float(audio_cfg.get("window_len", 25.)),
                  window_stride=float(audio_cfg.get("window_stride", 10.)),
                  preemp=bool(fe_cfg.get("preemp", False)),
                  init_min_freq=float(fe_cfg.get("min_freq", 60.0)),
                  init_max_freq=float(fe_cfg.get("max_freq", 7800.0)),
                  mean_var_norm=bool(fe_cfg.get("mean_var_norm", False)),
                  pcen_compression=bool(fe_cfg.get("pcen_compress", True)),
                  **common)
    pretrained = fe_cfg.get("pretrained", "")
    if pretrained and os.path.isfile(pretrained):
        fe.load_state_dict(torch.load(pretrained, map_location="cpu"))
    return fe
