"""distributedarrays.jl_b200 -- B200-native backend for the DArray map!/broadcast + mapreduce hot path.

Import it as ``darray_b200`` (the directory name is not a Python identifier; ``darray_b200.py`` at the repo root is a
loader shim).  The public names mirror DistributedArrays.jl's for this path: ``DArray``, ``distribute``, ``localpart``,
``localindices``, ``locate``, ``makelocal``, ``procs``, ``dzeros/dones/dfill/drand``, ``map`` (``map_``), ``map!``
(``map_inplace``), broadcast (``broadcast`` / ``broadcast_into``), ``reduce``, ``mapreduce``, ``sum``, ``prod``,
``maximum``, ``minimum``, ``all``, ``any``, ``count``, ``extrema``, ``Array(d)`` (``to_array``), range ``getindex``.

Everything computes on the GPU through ``csrc/libdab200.so`` (C ABI: ``include/dab200.h``).  There is no CPU fallback:
importing works anywhere, but the first op without the built extension or without a B200 raises.
"""
from . import _lib
from ._lib import ArgumentError, DabError, DimensionMismatch, UnsupportedError
from ._broadcast import (Expr, Int128, abs2, broadcast, broadcast_into, ceil, copy, cos, deepcopy, drandn, exp, floor, ifelse, inv, isnan, jl_max, jl_min,
                        log, map_, map_bang, map_inplace, map_localparts, mod, rem, sign, sin, sqrt, tan, tanh, widen)
from ._broadcast import (acos, acosh, acot, acoth, acsc, acsch, asec, asech, asin, asinh, atan, atanh, cbrt, cosh, cospi, cot, coth, csc,  # noqa: F401
                        csch, deg2rad, erf, erfc, erfcinv, erfcx, erfinv, exp10, exp2, expm1, gamma, isfinite, isinf, log10, log1p, log2,
                        loggamma, rad2deg, round_, sec, sech, sinh, sinpi, trunc)
from ._darray import (B200Array, DArray, SubDArray, allowscalar, dab_dtype, np_dtype, copyto, d_closeall, darray, darray_from_chunks, darray_like,
                     dfill, distribute, dones, drand, dzeros, fill_, localindices, localpart, locate, makelocal, pinned_empty, procs,
                     registry_size, reshape, similar, to_array)
from .layout import Layout, chunk_idxs, cuts_for, defaultdist, make_layout, slab_plan
from ._mapreduce import (all, any, axpy_, count, dot, extrema, isequal, mapreduce, mapreducedim, maximum, mean, minimum, nnz, norm,  # noqa: A004
                         prod, reduce, rmul_, sum)
from ._linalg import Adjoint, Transpose, adjoint, copy_transposed, lmul_diag, matmat, matmul, mul_, mul_mat_, rmul_diag, transpose
from ._sort import sort, sort_with_boundaries
from .runtime import Runtime, init, myid, nworkers, runtime, workers

__all__ = [n for n in dir() if not n.startswith("_")]
