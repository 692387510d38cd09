"""rqae_b200 -- B200 (sm_100a) implementation of the RQAE residual-quantization hot path.

Drop-in for ``rqae.model.RQAE`` of harish-kamath/rqae (same constructor, state-dict
layout, ``forward / decode / hook`` API and code-tensor layout, plus ``encode``); the
compute is hand-written CUDA behind the C ABI of ``librqae_b200.so``
(``include/rqae_b200.h``).  CUDA only: there is no CPU fallback."""
from .model import RQAE  # noqa: F401
from .feature import Feature, FeatureHelper, RQAEFeature, intensity_many  # noqa: F401
from .search import IntensityEngine  # noqa: F401
from . import shard  # noqa: F401
from . import store  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["RQAE", "Feature", "RQAEFeature", "intensity_many", "IntensityEngine", "FeatureHelper"]
__version__ = "0.1.0"
