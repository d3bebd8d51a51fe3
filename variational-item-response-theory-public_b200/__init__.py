"""B200-native VIBO ELBO engine: drop-in VIBO_{1,2,3}PL modules over
hand-written sm_100a CUDA kernels (libvibo_b200.so, C ABI in
include/vibo_b200.h).  Import as ``vibo_b200`` (shim at the repo root)."""
from . import _lib, distributed, functional, kernels  # noqa: F401
from .flows import NormalizingFlows, PlanarFlow  # noqa: F401
from .decoders import DeepIRT, LinkedIRT, ResidualIRT  # noqa: F401
from .models import (VI_1PL, VI_2PL, VI_3PL, VIBO_1PL, VIBO_2PL, VIBO_3PL, AbilityInferenceNetwork,  # noqa: F401
                     ConditionalAbilityInferenceNetwork, ItemInferenceNetwork)

__version__ = "0.1.0"
