"""videopainter_b200 — B200-native (sm_100a) denoising hot path of VideoPainter: CogVideoX-5B-I2V backbone forward plus the
context-encoder branch, behind the reference's model `forward` signatures.  See DESIGN.md / INTEGRATION.md."""
from .models import (CogVideoXTransformer3DModel, CogvideoXBranchModel, install, uninstall, invalidate)  # noqa: F401
from . import engine, graphs, ops, parallel, step_end  # noqa: F401
from .graphs import enable_graphs, graphs_enabled  # noqa: F401

__all__ = ["CogVideoXTransformer3DModel", "CogvideoXBranchModel", "install", "uninstall", "invalidate", "enable_graphs", "graphs_enabled", "engine", "graphs", "ops", "parallel", "step_end"]
