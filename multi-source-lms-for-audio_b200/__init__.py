"""vq-b200: the VQ-VAE vector-quantiser bottleneck of deborahdore/multi-source-lms-for-audio, rebuilt for B200.

Only the hot path lives here: `csrc/` (CUDA kernels + C ABI -> libvqb_b200.so), the ctypes binding, the
`VectorQuantizer` drop-in and the small host-side helpers either side of it.
"""
from ._lib import LIB_PATH, VqbError
from .quantizer import VectorQuantizer
from .vqvae_step import VQVAEStep
from .index_export import Quantize, export_windows
from . import functional, distributed, codebook_io

__all__ = ["VectorQuantizer", "VQVAEStep", "Quantize", "export_windows", "functional", "distributed", "codebook_io", "LIB_PATH",
           "VqbError"]
