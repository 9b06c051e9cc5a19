"""gaussdca.jl_b200 -- the gDCA hot path of carlobaldassi/GaussDCA.jl on NVIDIA B200 (sm_100a).

csrc/      hand-written CUDA kernels + the C ABI (include/gdca_b200.h) -> libgdca_b200.so
api.py     host-side mirror of the reference's public surface (gDCA, printrank, staged pieces)
fasta.py   host I/O that stays on the CPU (FASTA parse, gap filter, dedup)
dist.py    one-process-per-GPU driver (torch.distributed / NCCL) for the sharded stages
julia/     the Julia wrapper module (ccall) that replaces src/GaussDCA.jl

There is no CPU fallback: importing works anywhere, calling needs libgdca_b200.so and a B200.
"""
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all
from ._lib import Context, GdcaError, LIB_PATH, RANK_DTYPE, default_context, load  # noqa: F401

__all__ = list(_api_all) + ["Context", "GdcaError", "LIB_PATH", "RANK_DTYPE", "default_context", "load"]
