"""newtonkrylov.jl_b200 — B200-native JFNK inner loop behind the API of vchuravy/NewtonKrylov.jl.

The directory name contains a dot, so import it through the shim at the repository root:

    import newtonkrylov_jl_b200 as nk

`nk.host` mirrors the reference interface (newton_krylov_, JacobianOperator, Krylov k* hooks,
HaloVector, implicit `solve`) on top of the C ABI (`include/ariadne_b200.h`, bound in `_lib`).
Importing the package does not load the CUDA library; the first call does, and raises if the
library is missing (there is no CPU fallback).
"""
from . import _abi, _lib, build as _build  # noqa: F401
from ._lib import AriadneError, LIB_PATH  # noqa: F401
from .host import *  # noqa: F401,F403
from . import host, dist  # noqa: F401

__all__ = [n for n in dir(host) if not n.startswith("_")]
