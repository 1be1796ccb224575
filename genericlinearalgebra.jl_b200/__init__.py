"""B200-native Householder-QR / Cholesky-update hot path of GenericLinearAlgebra.jl.

csrc/   hand-written sm_100a kernels + the C ABI (include/gla_cuda.h) -> lib/libgla_cuda.so
glacuda.py   host-side mirror of the reference's operator interface over that C ABI (ctypes)
julia/  the `ccall` shim a Julia host loads (unexecuted here: no Julia runtime in the image)

The directory name contains a dot, so it is loaded by path: see `load()` in /__graft_entry__.py.
"""
from .glacuda import *  # noqa: F401,F403
from . import glacuda  # noqa: F401
from . import sharding  # noqa: F401
from .sharding import shard_range, tsqr_R_sharded  # noqa: F401
