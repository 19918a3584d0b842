"""leaf_pytorch_b200: the LEAF audio frontend hot path (leaf_pytorch.frontend.Leaf.forward of
SarthakYadav/leaf-pytorch) as hand-written sm_100a CUDA kernels behind the reference's own
module / factory API.  See DESIGN.md and INTEGRATION.md."""
from .frontend import Leaf
from .frontend_helper import get_frontend
from .functional import LeafSpec, leaf_forward, forward_raw, forward_window, launch_count
from ._native import LeafNativeError, LIB_PATH
from . import integration, streaming, distributed, serving, classifier
from .serving import HostPipeline
from .classifier import Classifier

__all__ = ["Leaf", "get_frontend", "LeafSpec", "leaf_forward", "forward_raw", "forward_window",
           "launch_count", "LeafNativeError", "LIB_PATH", "HostPipeline", "Classifier"]
