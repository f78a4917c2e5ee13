"""superconductor_vae_b200 - B200-native KV-cache decode engine behind the reference's Python API.

Only the hot path of jamesconde/superconductor-vae lives here: the
`EnhancedTransformerDecoder` KV-cache generation conditioned on `FullMaterialsVAE` latents.
CUDA kernels (sm_100a) and the C ABI are under ``csrc/`` / ``include/scvae_b200.h``; there is no CPU fallback.
"""
from ._lib import EngineError, build, launch_count, lib  # noqa: F401
from .decoder import EnhancedTransformerDecoder, END_IDX, PAD_IDX, START_IDX  # noqa: F401
from .encoder import FullMaterialsVAE  # noqa: F401
from .tokenizer import FractionAwareTokenizer  # noqa: F401
from . import latent, parallel  # noqa: F401

__all__ = ["EnhancedTransformerDecoder", "FullMaterialsVAE", "FractionAwareTokenizer", "EngineError", "build",
           "launch_count", "lib", "latent", "parallel", "PAD_IDX", "START_IDX", "END_IDX"]
