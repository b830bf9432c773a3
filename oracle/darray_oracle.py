"""CPU oracle for the DArray map!/broadcast + mapreduce hot path.

*** TEST INFRASTRUCTURE ONLY ***  Nothing under ``oracle/`` is part of the product.
Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py``
(``cpu_baseline`` / ``--impl reference``) may import or execute it, and there
only as the checker / baseline -- never as the thing shipped or measured as the
GPU result.

What this is: a NumPy restatement of what DistributedArrays.jl v0.6.9 executes on
the hot path.  Citations are relative to ``/root/reference``.  The reference is
100 % Julia and Julia is not installed in this image, so the reference itself
cannot be run; the arithmetic of the path lives in Julia ``Base`` (not vendored
in the reference tree).  Consequently:

* integer / index / layout / data-movement / elementwise results are PINNED:
  they are checked against every known-answer vector the reference's own tests
  hold (``tests/test_oracle_golden.py`` lists them with file:line);
* the *order* of floating-point reductions is a model of Julia Base
  (pairwise, block 1024, SIMD base block), validated only against the single
  golden the reference documents (``docs/src/index.md:222-225``,
  ``sum(fill(1.1,(100,100))) == 11000.000000000013``).  Float32 reduction bits
  are "parity unpinned" -- the binding contract is BASELINE's 1e-6 relative
  tolerance, checked against this model AND an exact (integer) ground truth.

Storage convention: Julia arrays are column-major; every chunk here is a NumPy
array in Fortran order so that ``ravel(order="F")`` is Julia's linear order.
Indices are 1-based inclusive ranges ``(lo, hi)`` exactly as in the reference.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

Range = Tuple[int, int]  # 1-based inclusive (lo, hi); empty when hi < lo

# --------------------------------------------------------------------------
# Layout math  (src/darray.jl:251-307, 448-456)
# --------------------------------------------------------------------------


def _prime_factors_desc(n: int) -> List[int]:
    """``sort!(collect(keys(factor(np))), rev=true)`` (src/darray.jl:255)."""
    out, p = [], 2
    while n > 1 and p * p <= n:
        if n % p == 0:
            out.append(p)
            while n % p == 0:
                n //= p
        p += 1
    if n > 1:
        out.append(n)
    return sorted(out, reverse=True)


def defaultdist_grid(dims: Sequence[int], npids: int) -> List[int]:
    """Process-grid shape: ``defaultdist(dims, pids)`` src/darray.jl:251-276.

    Repeatedly hands the largest remaining prime factor of ``np`` to the
    currently largest dimension, ties resolved to the HIGHEST dimension
    (``findlast``, :266-268); a factor is dropped when the dimension is smaller
    than it (:269-272).
    """
    dims = list(dims)
    chunks = [1] * len(dims)
    np_ = npids
    f = _prime_factors_desc(np_)
    k = 0
    while np_ > 1:
        if np_ % f[k] != 0:
            k += 1
            if k >= len(f):
                break
        fac = f[k]
        d = max(dims)
        dno = max(i for i, v in enumerate(dims) if v == d)
        if dims[dno] >= fac:
            dims[dno] //= fac
            chunks[dno] *= fac
        np_ //= fac
    return chunks


def defaultdist_cuts(sz: int, nc: int) -> List[int]:
    """Cut points: ``defaultdist(sz::Int, nc::Int)`` src/darray.jl:279-296.

    ``cuts[i]`` is the first (1-based) index of chunk ``i``; length ``nc+1``.
    The first ``rem(sz,nc)`` chunks get one extra element.
    """
    if sz >= nc:
        chunk_size, remainder = divmod(sz, nc)
        grid = []
        for i in range(1, nc + 2):
            g = (i - 1) * chunk_size + 1
            g += (i - 1) if i <= remainder else remainder
            grid.append(g)
        return grid
    return list(range(1, sz + 2)) + [0] * (nc - sz)


def chunk_idxs(dims: Sequence[int], chunks: Sequence[int]):
    """``chunk_idxs(dims, chunks)`` src/darray.jl:299-307.

    Returns ``(idxs, cuts)``; ``idxs`` is a dict keyed by 1-based grid
    CartesianIndex tuples, value = tuple of (lo,hi) ranges.
    """
    cuts = [defaultdist_cuts(d, c) for d, c in zip(dims, chunks)]
    idxs = {}
    for cidx in grid_iter(chunks):
        # an empty UnitRange a:b with b < a-1 is normalised to a:a-1 by Julia's UnitRange constructor
        idxs[cidx] = tuple((cuts[i][cidx[i] - 1], max(cuts[i][cidx[i] - 1] - 1, cuts[i][cidx[i]] - 1)) for i in range(len(dims)))
    return idxs, cuts


def grid_iter(shape: Sequence[int]):
    """CartesianIndices(shape) in Julia (column-major, first index fastest), 1-based."""
    for rev in itertools.product(*[range(1, s + 1) for s in reversed(shape)]):
        yield tuple(reversed(rev))


def locate(cuts: Sequence[Sequence[int]], I: Sequence[int]) -> Tuple[int, ...]:
    """``locate(d, I...)`` src/darray.jl:448-456 (``searchsortedlast`` on cuts)."""
    out = []
    for c, i in zip(cuts, I):
        fi = sum(1 for v in c if v <= i)  # searchsortedlast on an ascending vector
        # cuts may carry trailing zeros when sz < nc (:294); searchsortedlast assumes
        # sortedness, the reference never locates in that degenerate layout.
        if fi >= len(c):
            raise ValueError("element not contained in array")
        out.append(fi)
    return tuple(out)


def default_nprocs(dims: Sequence[int], nworkers: int) -> int:
    """``workers()[1:min(nworkers(), maximum(dims))]`` src/darray.jl:174,545."""
    return min(nworkers, max(dims))


def rlen(r: Range) -> int:
    return max(0, r[1] - r[0] + 1)


# --------------------------------------------------------------------------
# DArray model  (src/darray.jl:25-31)
# --------------------------------------------------------------------------


@dataclass
class ODArray:
    """Oracle DArray: metadata exactly as the reference struct, chunks on the host."""

    dims: Tuple[int, ...]
    grid: Tuple[int, ...]                      # size(pids)
    pids: List[int]                            # vec(pids): column-major grid order == procs(d)
    indices: List[Tuple[Range, ...]]           # vec(indices), same order
    cuts: List[List[int]]
    chunks: List[np.ndarray] = field(default_factory=list)  # Fortran-ordered localparts

    @property
    def ndim(self):
        return len(self.dims)

    def grid_index(self, lin: int) -> Tuple[int, ...]:
        """1-based Cartesian grid index of the lin-th (0-based) chunk."""
        out, r = [], lin
        for g in self.grid:
            out.append(r % g + 1)
            r //= g
        return tuple(out)


def make_layout(dims, pids: Sequence[int], dist: Optional[Sequence[int]] = None) -> ODArray:
    """``DArray(init, dims, procs, dist)`` src/darray.jl:159-166 (+ :168-173 for default dist)."""
    dims = tuple(int(d) for d in dims)
    if len(pids) == 0:
        raise ValueError("no processors given")  # src/darray.jl:169-171
    dist = list(dist) if dist is not None else defaultdist_grid(dims, len(pids))
    np_ = int(np.prod(dist))
    pids = list(pids)[:np_]
    idxs, cuts = chunk_idxs(dims, dist)
    order = list(grid_iter(dist))
    return ODArray(dims, tuple(dist), pids, [idxs[c] for c in order], cuts)


def distribute(A: np.ndarray, nworkers: int = None, procs: Sequence[int] = None,
               dist: Sequence[int] = None) -> ODArray:
    """``distribute(A; procs, dist)`` src/darray.jl:544-555: slice A by chunk_idxs."""
    A = np.asarray(A)
    if procs is None:
        nw = nworkers if nworkers is not None else 1
        procs = list(range(1, default_nprocs(A.shape, nw) + 1))
    d = make_layout(A.shape, procs, dist)
    d.chunks = [np.asfortranarray(A[tuple(slice(lo - 1, hi) for lo, hi in idx)]) for idx in d.indices]
    return d


def distribute_like(A: np.ndarray, DA: ODArray) -> ODArray:
    """``distribute(A, DA)`` src/darray.jl:563-570."""
    if tuple(A.shape) != DA.dims:
        raise ValueError("DimensionMismatch")
    d = ODArray(DA.dims, DA.grid, list(DA.pids), list(DA.indices), [list(c) for c in DA.cuts])
    d.chunks = [np.asfortranarray(A[tuple(slice(lo - 1, hi) for lo, hi in idx)]) for idx in d.indices]
    return d


def from_chunks(chunk_grid: Sequence[Sequence[np.ndarray]] | np.ndarray, grid: Sequence[int],
                pids: Sequence[int]) -> ODArray:
    """``DArray(refs)`` src/darray.jl:183-216: layout derived from the chunk sizes.

    ``chunk_grid`` is a flat list in column-major grid order.
    """
    grid = tuple(grid)
    order = list(grid_iter(grid))
    sizes = {c: chunk_grid[i].shape for i, c in enumerate(order)}
    indices = []
    for c in order:
        rng = []
        for x in range(len(grid)):
            start = 1
            for j in range(1, c[x]):
                prev = tuple(j if y == x else c[y] for y in range(len(grid)))
                start += sizes[prev][x]
            rng.append((start, start + sizes[c][x] - 1))
        indices.append(tuple(rng))
    cuts = []
    for x in range(len(grid)):
        lasts = sorted({idx[x][1] + 1 for idx in indices})
        cuts.append([1] + lasts)
    dims = tuple(c[-1] - 1 for c in cuts)
    d = ODArray(dims, grid, list(pids), indices, cuts)
    d.chunks = [np.asfortranarray(a) for a in chunk_grid]
    return d


def to_array(d: ODArray) -> np.ndarray:
    """``Array(d)`` src/darray.jl:574-582."""
    a = np.empty(d.dims, dtype=d.chunks[0].dtype, order="F")
    for idx, ch in zip(d.indices, d.chunks):
        if all(rlen(r) > 0 for r in idx):
            a[tuple(slice(lo - 1, hi) for lo, hi in idx)] = ch
    return a


def localindices(d: ODArray, pid: int) -> Tuple[Range, ...]:
    """``localindices(d)`` on worker ``pid`` src/darray.jl:394-400."""
    for p, idx in zip(d.pids, d.indices):
        if p == pid:
            return idx
    return tuple((1, 0) for _ in d.dims)


def makelocal_view_ranges(d: ODArray, pid: int, J: Sequence[Range]):
    """Local branch of ``makelocal`` src/darray.jl:351-360.

    Returns the 1-based local ranges into localpart(d) when J is inside the
    local indices, else ``None`` (the halo-fetch branch :361-366).
    """
    lid = localindices(d, pid)
    ok = all((rlen(j) == 0) or (l[0] <= j[0] and j[1] <= l[1]) for l, j in zip(lid, J))
    if not ok:
        return None
    return tuple((j[0] - (l[0] - 1), j[1] - (l[0] - 1)) for l, j in zip(lid, J))


# --------------------------------------------------------------------------
# Halo read: Array(view(d, I...))  (src/darray.jl:584-602, 798-820)
# --------------------------------------------------------------------------


def slab_plan(d: ODArray, J: Sequence[Range]):
    """Per intersecting chunk: (chunk#, local src ranges, dst ranges, whole_chunk?).

    Follows ``setindex!(a::Array, s::SubDArray, I...)`` src/darray.jl:798-820 for the
    unit-range case: ``K = J ∩ K_c`` (:805), whole chunk when ``K == K_c`` (:809),
    else owner-side ``localpart(d)[K .- (first(K_c)-1)]`` (:814-815); destination
    indices are K's positions inside J (what indexin_mask/restrict_indices compute, :807-808).
    """
    plan = []
    for c, Kc in enumerate(d.indices):
        K = tuple((max(j[0], k[0]), min(j[1], k[1])) for j, k in zip(J, Kc))
        if any(rlen(r) == 0 for r in K):
            continue
        whole = all(a == b for a, b in zip(K, Kc))
        src = tuple((k[0] - (kc[0] - 1), k[1] - (kc[0] - 1)) for k, kc in zip(K, Kc))
        dst = tuple((k[0] - (j[0] - 1), k[1] - (j[0] - 1)) for k, j in zip(K, J))
        plan.append((c, src, dst, whole))
    return plan


def getindex_array(d: ODArray, J: Sequence[Range]) -> np.ndarray:
    """``Array(d[J...])`` for unit ranges: src/darray.jl:661 (view) + :584-602."""
    out = np.empty([rlen(j) for j in J], dtype=d.chunks[0].dtype, order="F")
    for c, src, dst, _ in slab_plan(d, J):
        out[tuple(slice(lo - 1, hi) for lo, hi in dst)] = d.chunks[c][tuple(slice(lo - 1, hi) for lo, hi in src)]
    return out


def getindex_general(d: ODArray, I: Sequence) -> np.ndarray:
    """``Array(d[I...])`` for any mix of Int, UnitRange ``(lo, hi)``, StepRange / Vector{Int} (1-D int arrays of 1-based indices):
    ``setindex!(a::Array, s::SubDArray, I...)`` src/darray.jl:798-820, chunk by chunk --
    ``K_mask = map(indexin_mask, J, K_c)`` (:807, which positions of J fall into the chunk), ``idxs = restrict_indices(Inew, K_mask)``
    (:808, where they land in ``a``), ``localidxs = K .- (first(K_c) - 1)`` (:813) and ``a[idxs...] = localpart(d)[localidxs...]`` (:814).
    Scalar indices drop their dimension (view semantics)."""
    J, drop = [], []
    for ix in I:
        if isinstance(ix, (int, np.integer)):
            J.append(np.array([int(ix)], dtype=np.int64))
            drop.append(True)
        elif isinstance(ix, tuple):
            J.append(np.arange(ix[0], ix[1] + 1, dtype=np.int64))
            drop.append(False)
        else:
            J.append(np.asarray(ix, dtype=np.int64))
            drop.append(False)
    for j, n in zip(J, d.dims):
        if j.size and (j.min() < 1 or j.max() > n):
            raise IndexError("BoundsError")
    a = np.empty([len(j) for j in J], dtype=d.chunks[0].dtype, order="F")
    for c, Kc in enumerate(d.indices):
        masks = [(j >= k[0]) & (j <= k[1]) for j, k in zip(J, Kc)]                # indexin_mask(J, K_c)
        if any(not m.any() for m in masks):
            continue
        idxs = [np.nonzero(m)[0] for m in masks]                                    # restrict_indices: positions inside a
        local = [j[m] - k[0] for j, m, k in zip(J, masks, Kc)]                      # 0-based local indices in the chunk
        a[np.ix_(*idxs)] = d.chunks[c][np.ix_(*local)]
    return a[tuple(0 if dr else slice(None) for dr in drop)]


def darray_from_view(d: ODArray, I: Sequence) -> ODArray:
    """``DArray(SD::SubDArray)`` src/darray.jl:603-609: ``DArray(size(SD), procs(D)) do I; convert(Array, D[reindex(SD.indices, I)...])``
    -- default distribution of the view's size over procs(D); chunk values = the matching block of the gathered view."""
    full = getindex_general(d, I)
    out = make_layout(full.shape, d.pids)
    out.chunks = [np.asfortranarray(full[tuple(slice(lo - 1, hi) for lo, hi in ix)]) for ix in out.indices]
    return out


# --------------------------------------------------------------------------
# Scalar semantics of Julia Base used on the path (Appendix A of SURVEY.md)
# --------------------------------------------------------------------------


def jl_max(x, y):
    """Julia ``max`` on floats: NaN-propagating, ``+0.0 > -0.0``; exact on ints."""
    x, y = np.asarray(x), np.asarray(y)
    if not np.issubdtype(x.dtype, np.floating):
        return np.maximum(x, y)
    r = np.where((y > x) | (np.signbit(x) & ~np.signbit(y) & (x == y)), y, x)
    return np.where(np.isnan(x) | np.isnan(y), np.asarray(np.nan, dtype=x.dtype), r).astype(x.dtype)


def jl_min(x, y):
    x, y = np.asarray(x), np.asarray(y)
    if not np.issubdtype(x.dtype, np.floating):
        return np.minimum(x, y)
    r = np.where((y < x) | (~np.signbit(x) & np.signbit(y) & (x == y)), y, x)
    return np.where(np.isnan(x) | np.isnan(y), np.asarray(np.nan, dtype=x.dtype), r).astype(x.dtype)


def _add(x, y):
    with np.errstate(over="ignore", invalid="ignore"):
        return x + y


def _mul(x, y):
    with np.errstate(over="ignore", invalid="ignore", under="ignore"):
        return x * y


OPS = {"+": _add, "*": _mul, "max": jl_max, "min": jl_min}


def widen_for(op: str, dtype) -> np.dtype:
    """``add_sum`` / ``mul_prod`` widen Int8/16/32 (UInt8/16/32) to Int (UInt); floats stay."""
    dtype = np.dtype(dtype)
    if op in ("+", "*") and dtype.kind == "i" and dtype.itemsize < 8:
        return np.dtype(np.int64)
    if op in ("+", "*") and dtype.kind == "u" and dtype.itemsize < 8:
        return np.dtype(np.uint64)
    if op in ("+", "*") and dtype.kind == "b":
        return np.dtype(np.int64)
    return dtype


def default_simd(dtype) -> Tuple[int, int]:
    """(lanes, interleave) LLVM picks for an ``@simd`` reduction on an AVX2 host."""
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return (32 // dtype.itemsize, 4)
    return (1, 1)  # integer reductions are exact; order is irrelevant


# --------------------------------------------------------------------------
# Whole-array reduction: Base._mapreduce / mapreduce_impl  
# reached from src/mapreduce.jl:23,31
# --------------------------------------------------------------------------

PAIRWISE_BLOCK = 1024


def _base_block(op, v: np.ndarray, lanes: int, interleave: int) -> np.ndarray:
    """Sequential base case of ``mapreduce_impl`` on the LAST axis of ``v``.

    ``v0 = op(a1, a2); @simd for i = 3:n  v0 = op(v0, a_i)``.  ``@simd`` licenses
    reassociation; LLVM's loop vectoriser turns it into W = lanes*interleave
    running accumulators (accumulator 0 seeded with v0, the others with the
    identity, modelled here by seeding them with their first element), combined
    after the loop as ((acc0 + acc1) + acc2) + acc3 lane-wise across the
    interleaved vectors, then a log2 shuffle tree across lanes, then the scalar
    remainder loop.  lanes == interleave == 1 is the strict left fold.
    """
    n = v.shape[-1]
    W = lanes * interleave
    acc0 = op(v[..., 0], v[..., 1])
    m = n - 2
    nvec = m // W if W > 1 else 0
    if nvec == 0:
        r = acc0
        for i in range(2, n):
            r = op(r, v[..., i])
        return r
    body = v[..., 2:2 + nvec * W].reshape(v.shape[:-1] + (nvec, W))
    acc = body[..., 0, :].copy()
    acc[..., 0] = op(acc0, acc[..., 0])
    for j in range(1, nvec):
        acc = op(acc, body[..., j, :])
    # interleaved vectors: acc[..., u*lanes:(u+1)*lanes] is vector u
    vec = acc[..., 0:lanes]
    for u in range(1, interleave):
        vec = op(acc[..., u * lanes:(u + 1) * lanes], vec)
    w = lanes
    while w > 1:  # shuffle tree: x[i] + x[i + w/2]
        h = w // 2
        vec = op(vec[..., :h], vec[..., h:w])
        w = h
    r = vec[..., 0]
    for i in range(2 + nvec * W, n):
        r = op(r, v[..., i])
    return r


def _pairwise(op, v: np.ndarray, lanes: int, interleave: int, blk: int = PAIRWISE_BLOCK):
    """``mapreduce_impl(f, op, A, ifirst, ilast, blksize)`` on the last axis (f pre-applied)."""
    n = v.shape[-1]
    if n == 1:
        return v[..., 0]
    if n - 1 < blk:  # ``ilast - ifirst < blksize``
        return _base_block(op, v, lanes, interleave)
    half = ((n - 1) >> 1) + 1  # imid = ifirst + (ilast-ifirst)>>1, inclusive
    return op(_pairwise(op, v[..., :half], lanes, interleave, blk),
              _pairwise(op, v[..., half:], lanes, interleave, blk))


def julia_mapreduce(f: Optional[Callable], op: str, A: np.ndarray, simd=None, init=None):
    """``mapreduce(f, op, A::Array)`` (IndexLinear): returns a NumPy scalar of the result type.

    n == 0 -> ValueError for max/min, zero/one for +/*;  n == 1 -> f(a1);
    n < 16 -> plain left fold;  else pairwise with block 1024.
    """
    a = np.asarray(A).ravel(order="F")
    v = f(a) if f is not None else a
    v = np.asarray(v)
    v = v.astype(widen_for(op, v.dtype), copy=False)
    fn = OPS[op]
    n = v.shape[0]
    if n == 0:
        if op == "+":
            r = v.dtype.type(0)
        elif op == "*":
            r = v.dtype.type(1)
        else:
            raise ValueError("reducing over an empty collection is not allowed")
    elif n == 1:
        r = v[0]
    elif n < 16:
        r = v[0]
        for i in range(1, n):
            r = fn(r, v[i])
    else:
        lanes, inter = simd if simd is not None else default_simd(v.dtype)
        r = _pairwise(fn, v, lanes, inter)
    r = np.asarray(r, dtype=v.dtype)[()]
    if init is not None:
        r = fn(v.dtype.type(init), r)
    return r


def darray_mapreduce(f, op: str, d: ODArray, simd=None):
    """``Base._mapreduce(f, op, ::IndexCartesian, d::DArray)`` src/mapreduce.jl:29-35.

    One ``mapreduce(f, op, localpart)`` per worker (:31), then ``reduce(op, results)``
    on the caller (:34): P < 16 so a plain left fold in ``procs(d)`` order.
    """
    results = [julia_mapreduce(f, op, ch, simd) for ch in d.chunks]
    fn = OPS[op]
    r = results[0]
    for x in results[1:]:
        r = fn(r, x)
    return np.asarray(r)[()], results


def wrap_int128(v: int) -> int:
    """Two's-complement wrap-around of Julia's Int128 machine arithmetic."""
    v &= (1 << 128) - 1
    return v - (1 << 128) if v >> 127 else v


def darray_mapreduce_int128(f: Callable[[int], int], op: str, d: ODArray) -> int:
    """``mapreduce(f, op, DA)`` for an Int128-valued ``f`` on an integer DArray -- the exactness test of test/darray.jl:286-294
    (``f in (x -> Int128(2x), x -> Int128(x^2), x -> Int128(x^2 + 2x - 1))``, ``op in (+, *)``).  Per worker
    ``mapreduce(f, op, localpart)`` (src/mapreduce.jl:31), then the caller's left fold (:34); + and * in Int128 wrap, so every
    grouping gives the same bits -- which is what makes the reference's ``== 0`` test meaningful.  Python integers, exact."""
    def red(vals):
        acc = vals[0]
        for v in vals[1:]:
            acc = wrap_int128(acc + v if op == "+" else acc * v)
        return acc
    parts = [red([wrap_int128(f(int(x))) for x in c.reshape(-1, order="F")]) for c in d.chunks if c.size]
    return red(parts)


def darray_all(pred, d: ODArray) -> bool:
    """src/mapreduce.jl:97-104."""
    return all(bool(np.all(pred(ch))) for ch in d.chunks)


def darray_any(pred, d: ODArray) -> bool:
    """src/mapreduce.jl:106-113."""
    return any(bool(np.any(pred(ch))) for ch in d.chunks)


def darray_count(pred, d: ODArray) -> int:
    """src/mapreduce.jl:115-122."""
    return int(sum(int(np.count_nonzero(pred(ch))) for ch in d.chunks))


def darray_extrema(d: ODArray):
    """src/mapreduce.jl:124-131: per-chunk extrema, then fold with (min, max)."""
    r = [(julia_mapreduce(None, "min", ch), julia_mapreduce(None, "max", ch)) for ch in d.chunks]
    t = r[0]
    for s in r[1:]:
        t = (jl_min(t[0], s[0])[()], jl_max(t[1], s[1])[()])
    return t


# --------------------------------------------------------------------------
# Dimensional reduction  (src/mapreduce.jl:42-94)
# --------------------------------------------------------------------------


def _init_value(op: str, dtype):
    if op == "+":
        return dtype.type(0)
    if op == "*":
        return dtype.type(1)
    return None


def julia_mapreducedim(f, op: str, A: np.ndarray, region: Sequence[int], R: np.ndarray = None, simd=None):
    """``Base.mapreducedim!(f, op, R, A)`` on a local Array; ``region`` is 1-based dims.

    R (if given) is accumulated onto (that is how ``init`` and the between-phase
    enter, src/mapreduce.jl:77).  Order follows ``_mapreducedim!``:
      * reduced dims are exactly the leading dims 1..k and the slice is longer
        than 16  -> per-slice pairwise ``mapreduce_impl``;
      * else reducing dim 1 -> sequential over dim 1 per trailing index;
      * else sequential sweep over the trailing indices (column by column).
    """
    A = np.asarray(A)
    fn = OPS[op]
    N = A.ndim
    region = sorted({int(r) for r in region if 1 <= int(r) <= N})
    v = f(A) if f is not None else A
    v = np.asfortranarray(np.asarray(v))
    v = v.astype(widen_for(op, v.dtype), copy=False)
    rshape = tuple(1 if (i + 1) in region else A.shape[i] for i in range(N))
    if R is None:
        iv = _init_value(op, v.dtype)
        if iv is not None:
            R = np.full(rshape, iv, dtype=v.dtype, order="F")
        else:
            # max/min: reducedim_init uses the extremum of the first slice; equivalent to
            # seeding with the first element along every reduced dim.
            first = v[tuple(slice(0, 1) if (i + 1) in region else slice(None) for i in range(N))]
            R = np.array(first, dtype=v.dtype, order="F", copy=True)
    else:
        R = np.array(R, dtype=v.dtype, order="F", copy=True).reshape(rshape, order="F")
    if v.size == 0:
        return R
    lanes, inter = simd if simd is not None else default_simd(v.dtype)
    k = len(region)
    if k == 0:
        return fn(R, v)
    # Base.check_reducedims: a dim counts as reduced when size(R, i) == 1; lsiz = product of the leading reduced extents,
    # 0 as soon as a reduced dim (of extent > 1) follows a kept one.
    lsiz, had_nonreduc = 1, False
    for i in range(N):
        if rshape[i] == 1:
            if A.shape[i] > 1:
                lsiz = 0 if had_nonreduc else lsiz * A.shape[i]
        else:
            had_nonreduc = True
    if lsiz > 16:
        flat = v.reshape((lsiz, -1), order="F")           # column = one contiguous slice
        s = _pairwise(fn, np.ascontiguousarray(flat.T), lanes, inter)
        return fn(R, s.reshape(rshape, order="F"))
    # sequential paths: move the reduced dims through a left fold in memory order
    out = R
    red_axes = [r - 1 for r in region]
    # iterate reduced index tuples in column-major order (first reduced dim fastest)
    red_shape = [A.shape[a] for a in red_axes]
    for rev in itertools.product(*[range(s) for s in reversed(red_shape)]):
        ridx = tuple(reversed(rev))
        sl = [slice(None)] * N
        for a, i in zip(red_axes, ridx):
            sl[a] = slice(i, i + 1)
        out = fn(out, v[tuple(sl)])
    return out


def darray_mapreducedim(f, op: str, d: ODArray, region: Sequence[int], init=None, simd=None):
    """``mapreduce(f, op, d::DArray; dims=region[, init])``: src/mapreduce.jl:42-94.

    Returns an ODArray R laid out on the "lowest" pids of each fibre (:44).
    Phase 1 ``mapreducedim_within`` (:54-66): every worker reduces its chunk.
    Phase 2 ``mapreducedim_between!`` (:71-81): each owner accumulates the
    partials of its fibre, in grid order along the reduced dims, onto R.
    """
    N = d.ndim
    for r in region:
        if int(r) <= 0:
            raise ValueError("ArgumentError: region dimension(s) must be ≥ 1, got %r" % (r,))
    region = sorted({int(r) for r in region})
    reg_in = [r for r in region if r <= N]
    # R layout: pids[1:1 along region, : elsewhere]  (:44)
    Rgrid = tuple(1 if (i + 1) in reg_in else g for i, g in enumerate(d.grid))
    order = list(grid_iter(d.grid))
    lin_of = {c: i for i, c in enumerate(order)}
    # phase 1: partials (no init passed, :64)
    f_applied = f
    partials = [julia_mapreducedim(f_applied, op, ch, reg_in, simd=simd) for ch in d.chunks]
    out_dtype = partials[0].dtype
    Rchunks, Rpids, Rindices = [], [], []
    for rc in grid_iter(Rgrid):
        # fibre members: all grid coords equal to rc outside region, any coord inside
        members = [c for c in order if all(c[i] == rc[i] for i in range(N) if (i + 1) not in reg_in)]
        owner = lin_of[rc]
        rshape = partials[owner].shape
        if init is not None:
            R = np.full(rshape, out_dtype.type(init), dtype=out_dtype, order="F")
        else:
            iv = _init_value(op, out_dtype)
            R = np.full(rshape, iv, dtype=out_dtype, order="F") if iv is not None else None
        # phase 2 (:76-77): Bfull = partials stacked along the region dims in grid order,
        # then Base.mapreducedim!(identity, op, localpart(R), Bfull): sequential (size<=16).
        for m in members:
            R = np.array(partials[lin_of[m]], copy=True) if R is None else OPS[op](R, partials[lin_of[m]])
        Rchunks.append(np.asfortranarray(R))
        Rpids.append(d.pids[owner])
        Rindices.append(tuple((1, 1) if (i + 1) in reg_in else d.indices[owner][i] for i in range(N)))
    Rdims = tuple(1 if (i + 1) in reg_in else s for i, s in enumerate(d.dims))
    Rcuts = [[1, 2] if (i + 1) in reg_in else list(d.cuts[i]) for i in range(N)]
    R = ODArray(Rdims, Rgrid, Rpids, Rindices, Rcuts, Rchunks)
    return R


# --------------------------------------------------------------------------
# Elementwise: broadcast / map!  (src/broadcast.jl:65-98, src/mapreduce.jl:5-12)
# --------------------------------------------------------------------------


def affine_unfused(a, x: np.ndarray, b) -> np.ndarray:
    """``y .= a .* x .+ b`` in the element type: two IEEE roundings, never an FMA."""
    t = x.dtype.type
    with np.errstate(over="ignore", invalid="ignore"):
        return (t(a) * x) + t(b)


def darray_map_inplace(fn: Callable[[np.ndarray], np.ndarray], dest: ODArray, src: ODArray) -> ODArray:
    """``map!(f, dest, src)`` src/mapreduce.jl:5-12: per worker
    ``map!(f, localpart(dest), makelocal(src, localindices(dest)...))``."""
    for i, idx in enumerate(dest.indices):
        local = getindex_array(src, idx)  # == makelocal (local branch is a view of the same data)
        dest.chunks[i] = np.asfortranarray(fn(local).astype(dest.chunks[i].dtype, copy=False))
    return dest


def darray_broadcast(fn: Callable[..., np.ndarray], dest_layout: ODArray, *args, out_dtype=None) -> ODArray:
    """``copyto!(dest, bc)`` / ``copy(bc)`` src/broadcast.jl:65-98.

    Each arg is a scalar, an ODArray, or a NumPy array (auto-``distribute``d with
    the default layout by ``bcdistribute``, :124-137).  Per destination chunk the
    args are localised with ``_bcview`` (:103-120): size-1 dims are kept as 1:1
    (extrusion), others are cut to the chunk's ranges, then ``fn`` runs once.
    """
    out = ODArray(dest_layout.dims, dest_layout.grid, list(dest_layout.pids), list(dest_layout.indices),
                  [list(c) for c in dest_layout.cuts])
    N = len(out.dims)
    for idx in out.indices:
        largs = []
        for a in args:
            if isinstance(a, ODArray):
                ad = len(a.dims)
                view = []
                for k in range(ad):
                    if a.dims[k] == 1:
                        view.append((1, 1))
                    elif k < N:
                        view.append(idx[k])
                    else:
                        view.append((1, a.dims[k]))
                loc = getindex_array(a, view)
                largs.append(loc)
            elif isinstance(a, np.ndarray) and a.ndim > 0:
                sl = tuple(slice(0, 1) if a.shape[k] == 1 else slice(idx[k][0] - 1, idx[k][1]) for k in range(a.ndim))
                largs.append(np.asfortranarray(a[sl]))
            else:
                largs.append(a)
        shape = tuple(rlen(r) for r in idx)
        with np.errstate(all="ignore"):
            res = np.broadcast_to(fn(*largs), shape)
        if out_dtype is not None:
            res = res.astype(out_dtype)
        out.chunks.append(np.asfortranarray(res))
    return out


# --------------------------------------------------------------------------
# Synthetic inputs: counter-based RNG shared bit-for-bit with the CUDA side
# (distribution of Julia's rand(Float32): Float32(rand(UInt32) >>> 8) * 2f0^-24)
# --------------------------------------------------------------------------

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def hash_u32(seed: int, idx: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser of ``idx + (seed+1)*golden``; high 32 bits."""
    with np.errstate(over="ignore"):
        z = np.asarray(idx, dtype=np.uint64) + np.uint64((int(seed) + 1) & 0xFFFFFFFFFFFFFFFF) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint32)


def rand_u01(seed: int, start: int, n: int, dtype=np.float32) -> np.ndarray:
    """Element ``i`` (global linear index ``start+i``) = (hash >> 8) * 2^-24 in [0,1)."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    k = (hash_u32(seed, idx) >> np.uint32(8)).astype(np.float64)
    return (k * (2.0 ** -24)).astype(dtype)


def rand_u01_ksum(seed: int, start: int, n: int, block: int = 1 << 24) -> int:
    """Exact integer Σ k_i (so that the exact sum of the array is Σk · 2^-24)."""
    tot = 0
    for s in range(start, start + n, block):
        m = min(block, start + n - s)
        idx = np.arange(s, s + m, dtype=np.uint64)
        tot += int((hash_u32(seed, idx) >> np.uint32(8)).astype(np.uint64).sum())
    return tot


# --------------------------------------------------------------------------
# Level-2 linear algebra and transposes  (src/linalg.jl:1-17, 78-187, 278-311)
# --------------------------------------------------------------------------


def _tile_matvec(A: np.ndarray, x: np.ndarray, trans: bool) -> np.ndarray:
    """``localpart(A)*xj`` / ``localpart(A)'*xj`` (src/linalg.jl:95-97, 141).  Julia dispatches Float32/Float64 to BLAS gemv
    (summation order unspecified -> parity unpinned, tolerance contract) and integers to the generic wrap-around loop.  The
    restatement: floats summed in fp64 and rounded once; integers exactly, modulo 2^bits."""
    M = A.T if trans else A
    if A.dtype.kind == "f":
        return (M.astype(np.float64) @ x.astype(np.float64)).astype(A.dtype)
    with np.errstate(over="ignore"):
        return (M * x[None, :].astype(A.dtype)).sum(axis=1, dtype=A.dtype)


def _add_scaled(dest: np.ndarray, src: np.ndarray, scale) -> np.ndarray:
    """``add!(dest, src, scale)`` src/linalg.jl:62-76: ``dest[i] += src[i]`` when scale == 1, else ``dest[i] += scale*src[i]``
    (two roundings, no FMA)."""
    dt = dest.dtype
    if scale == 1:
        return (dest + src).astype(dt)
    return (dest + (dt.type(scale) * src).astype(dt)).astype(dt)


def darray_mul_vec(y: ODArray, A: ODArray, x: np.ndarray, alpha=1, beta=0, trans: bool = False) -> ODArray:
    """``mul!(y::DVector, A::DMatrix, x, α, β)`` src/linalg.jl:78-118 and the adjoint/transpose form :120-167.

    Tile products R[i,j]; y scaled by β (``fill!(0)`` when β == 0, untouched when β == 1); then ``add!(localpart(y), R[i,j], α)``
    for j = 1..  The reference issues those adds as @async tasks, so their order is not fixed; j order is used here."""
    rd, cd = (1, 0) if trans else (0, 1)          # A's dim that indexes y / that is contracted with x
    if A.dims[cd] != len(x):
        raise ValueError("DimensionMismatch")
    if list(y.cuts[0]) != list(A.cuts[rd]):
        raise ValueError("ArgumentError: cuts of output vector must match cuts of matrix")
    gi, gj = A.grid[rd], A.grid[cd]
    dt = y.chunks[0].dtype if y.chunks else A.chunks[0].dtype
    out = []
    for i in range(gi):
        yi = y.chunks[i].copy()
        if beta != 1:
            yi = (yi * dt.type(beta)).astype(dt) if beta != 0 else np.zeros_like(yi)
        for j in range(gj):
            lin = (j + i * A.grid[0]) if trans else (i + j * A.grid[0])   # procs(A)[j,i] / procs(A)[i,j]
            lo, hi = A.cuts[cd][j], A.cuts[cd][j + 1] - 1
            r = _tile_matvec(A.chunks[lin], np.asarray(x[lo - 1:hi]), trans)
            yi = _add_scaled(yi, r.astype(dt), alpha)
        out.append(yi)
    return ODArray(y.dims, y.grid, y.pids, y.indices, y.cuts, out)


def darray_matvec(A: ODArray, x: np.ndarray, trans: bool = False) -> ODArray:
    """``A*x`` src/linalg.jl:280-284 / ``A'*x`` :293-301: y over procs(A)[:,1] (resp. procs(A)[1,:]), one chunk per grid row
    (resp. column), then ``mul!(y, A, x)``."""
    rd = 1 if trans else 0
    g0 = A.grid[0]
    pids = [A.pids[j * g0] for j in range(A.grid[1])] if trans else [A.pids[i] for i in range(g0)]
    dt = np.result_type(A.chunks[0].dtype, np.asarray(x).dtype)
    y = make_layout((A.dims[rd],), pids, [A.grid[rd]])
    y.chunks = [np.zeros(rlen(ix[0]), dtype=dt) for ix in y.indices]   # uninitialised in the reference; β = 0 overwrites
    return darray_mul_vec(y, A, np.asarray(x), 1, 0, trans)


def _tile_matmat(A: np.ndarray, B: np.ndarray, trans: bool) -> np.ndarray:
    """``localpart(A)*Bjk`` / ``transpose(localpart(A))*Bjk`` (src/linalg.jl:218-226): BLAS gemm for floats (order unspecified ->
    tolerance contract; restated in fp64, rounded once), the generic wrap-around loop for integers."""
    M = A.T if trans else A
    if A.dtype.kind == "f":
        return np.asfortranarray((M.astype(np.float64) @ B.astype(np.float64)).astype(A.dtype))
    with np.errstate(over="ignore"):
        return np.asfortranarray(M @ B.astype(A.dtype))


def darray_mul_mat(C: ODArray, A: ODArray, B: np.ndarray, alpha=1, beta=0, trans: bool = False) -> ODArray:
    """``_matmatmul!(C::DMatrix, A::DMatrix, B::AbstractMatrix, α, β, tA)`` src/linalg.jl:189-257: tile products R[i,j,k] on
    procs(A)[i,j] (``[j,i]`` for tA in 'T','C'), C scaled by β, then ``add!(localpart(C), R[i,j,k], α)`` on C.pids[i,k] (j order here;
    the reference issues those adds as @async tasks)."""
    B = np.asarray(B)
    rd, cd = (1, 0) if trans else (0, 1)
    mA, nA = A.dims[rd], A.dims[cd]
    if B.shape[0] != nA:
        raise ValueError("DimensionMismatch: matrix A has dimensions (%d, %d), matrix B has dimensions %s" % (mA, nA, B.shape))
    if C.dims != (mA, B.shape[1]):
        raise ValueError("DimensionMismatch: result C has dimensions %s, needs (%d, %d)" % (C.dims, mA, B.shape[1]))
    if list(C.cuts[0]) != list(A.cuts[rd]):
        raise ValueError("ArgumentError: cuts of the first dimension of the output matrix must match cuts of the input matrix")
    gi, gj, gk = A.grid[rd], A.grid[cd], C.grid[1]
    dt = C.chunks[0].dtype
    out = [None] * len(C.chunks)
    for k in range(gk):
        clo, chi = C.cuts[1][k], C.cuts[1][k + 1] - 1
        for i in range(gi):
            lin_c = i + k * C.grid[0]
            ci = C.chunks[lin_c].copy()
            if beta != 1:
                ci = (ci * dt.type(beta)).astype(dt) if beta != 0 else np.zeros_like(ci)
            for j in range(gj):
                lin = (j + i * A.grid[0]) if trans else (i + j * A.grid[0])
                lo, hi = A.cuts[cd][j], A.cuts[cd][j + 1] - 1
                r = _tile_matmat(A.chunks[lin], B[lo - 1:hi, clo - 1:chi], trans)
                ci = _add_scaled(ci, r.astype(dt), alpha)
            out[lin_c] = np.asfortranarray(ci)
    return ODArray(C.dims, C.grid, C.pids, C.indices, C.cuts, out)


def darray_matmat(A: ODArray, B: ODArray, trans: bool = False) -> ODArray:
    """``A*B`` src/linalg.jl:285-292 and ``A'*B`` / ``transpose(A)*B`` :302-311 for DMatrix B: C over
    ``procs(A)[:, 1:min(size(procs(A),2), size(procs(B),2))]`` with grid ``(size(procs(A),1), that min)`` -- for the transposed forms
    ``procs(A)[1:min(size(procs(A),1), size(procs(B),2)), :]`` with grid ``(size(procs(A),2), that min)`` -- then ``mul!(C, A, B)``."""
    g0, g1 = A.grid
    pg = np.asarray(A.pids).reshape((g0, g1), order="F")
    Bfull = to_array(B)
    if not trans:
        nc = min(g1, B.grid[1])
        pids = list(pg[:, :nc].reshape(-1, order="F"))
        Cl = make_layout((A.dims[0], B.dims[1]), pids, [g0, nc])
    else:
        nr = min(g0, B.grid[1])
        pids = list(pg[:nr, :].reshape(-1, order="F"))
        Cl = make_layout((A.dims[1], B.dims[1]), pids, [g1, nr])
    dt = np.result_type(A.chunks[0].dtype, B.chunks[0].dtype)
    Cl.chunks = [np.zeros(tuple(rlen(r) for r in ix), dtype=dt, order="F") for ix in Cl.indices]
    return darray_mul_mat(Cl, A, Bfull, 1, 0, trans)


def darray_transpose(D: ODArray) -> ODArray:
    """``copy(::Transpose{T,<:DArray{T,2}})`` / Adjoint for real T, src/linalg.jl:1-17:
    ``DArray(reverse(size(D)), procs(D)) do I; transpose!(lp, Array(D[reverse(I)...]))``."""
    full = to_array(D)
    R = make_layout((D.dims[1], D.dims[0]), D.pids)
    R.chunks = [np.asfortranarray(full[I[1][0] - 1:I[1][1], I[0][0] - 1:I[0][1]].T) for I in R.indices]
    return R


def darray_scale_diag(DA: ODArray, d: np.ndarray, side: str) -> ODArray:
    """``lmul!(D::Diagonal, DA)`` (side 'l': rows scaled, ``d[i]*A[i,j]``) and ``rmul!(DA, D::Diagonal)`` (side 'r':
    ``A[i,j]*d[j]``), src/linalg.jl:169-187."""
    out = []
    for ch, I in zip(DA.chunks, DA.indices):
        if side == "l":
            s = np.asarray(d[I[0][0] - 1:I[0][1]]).astype(ch.dtype)[:, None]
            out.append(np.asfortranarray((s * ch).astype(ch.dtype)))
        else:
            s = np.asarray(d[I[1][0] - 1:I[1][1]]).astype(ch.dtype)[None, :]
            out.append(np.asfortranarray((ch * s).astype(ch.dtype)))
    return ODArray(DA.dims, DA.grid, DA.pids, DA.indices, DA.cuts, out)


# --------------------------------------------------------------------------
# Samplesort of a DVector  (src/sort.jl)
# --------------------------------------------------------------------------


def jl_sort(v: np.ndarray) -> np.ndarray:
    """``sort(v)`` in Julia's ``isless`` order: -0.0 before +0.0, NaNs last (Julia keeps the NaNs in input order)."""
    v = np.asarray(v)
    if v.dtype.kind != "f":
        return np.sort(v, kind="stable")
    nan = np.isnan(v)
    body = v[~nan]
    order = np.lexsort((~np.signbit(body), body))
    return np.concatenate([body[order], v[nan]])


def _typemin(dt):
    return -np.inf if np.dtype(dt).kind == "f" else np.iinfo(dt).min


def _typemax(dt):
    return np.inf if np.dtype(dt).kind == "f" else np.iinfo(dt).max


def sort_sample_indices(llp: int, sample_size: int = 512) -> range:
    """0-based indices of ``sorted[collect(1:div(llp,sample_size):llp)]`` with ``sample_size = min(sample_size, llp)``
    (src/sort.jl:3-9).  An empty localpart divides by zero in the reference."""
    ss = sample_size if llp > sample_size else llp
    if ss == 0:
        raise ZeroDivisionError("DivideError: integer division error")
    return range(0, llp, llp // ss)


def sort_boundaries_from_samples(samples: np.ndarray, nparts: int, dt) -> np.ndarray:
    """``sort!(samples); samples[1] = typemin(T); boundaries = samples[[1+(x-1)*div(length(samples), np) for x in 1:np]];
    push!(boundaries, typemax(T))`` (src/sort.jl:78-85, 149-153)."""
    s = jl_sort(np.asarray(samples, dtype=dt)).copy()
    s[0] = _typemin(dt)
    step = len(s) // nparts
    b = [s[(x - 1) * step] for x in range(1, nparts + 1)]
    b.append(_typemax(dt))
    return np.asarray(b, dtype=dt)


def sort_uniform_sample(lb, ub, nparts: int, dt) -> np.ndarray:
    """The ``sample::Tuple`` branch (src/sort.jl:127-145): ``s[n] = lb + (n-1)*abs(ub-lb)/np`` (rounded for integer T)."""
    assert lb <= ub
    if isinstance(lb, np.float32) and isinstance(ub, np.float32):
        part = np.float32(abs(ub - lb)) / np.float32(nparts)
        vals = [np.float32(lb + np.float32(n) * part) for n in range(nparts)]
    else:
        if np.dtype(dt).kind == "f" or not (isinstance(lb, (int, np.integer)) and isinstance(ub, (int, np.integer))):
            part = abs(float(ub) - float(lb)) / nparts
        else:  # abs(ub - lb) in T's wrap-around machine arithmetic (a full-range Int sample overflows, as in the reference)
            bits = 8 * np.dtype(dt).itemsize
            diff = (int(ub) - int(lb) + (1 << (bits - 1))) % (1 << bits) - (1 << (bits - 1))
            part = float(diff if diff == -(1 << (bits - 1)) else abs(diff)) / nparts
        vals = [float(lb) + n * part for n in range(nparts)]
    if np.isnan(part) or np.isinf(part):
        raise ValueError("ArgumentError: lower and upper bounds must not be infinities")
    if np.dtype(dt).kind != "f":
        vals = [np.rint(v) for v in vals]
    return np.asarray(vals).astype(dt)


def jl_sortperm_stable(keys: np.ndarray) -> np.ndarray:
    """Permutation of a STABLE sort of ``keys`` in ``isless`` order (-0.0 before +0.0; NaNs last and equal to each other, so
    they keep their input order): what ``sort(v; by = f)`` applies to ``v`` with ``keys = f.(v)`` (Julia's default algorithm for a
    keyed sort is stable)."""
    k = np.asarray(keys)
    if k.dtype.kind != "f":
        return np.argsort(k, kind="stable")
    nan = np.isnan(k)
    return np.lexsort((~np.signbit(k) & ~nan, np.where(nan, 0, k), nan))        # last key is the primary one; lexsort is stable


def jl_sort_by(v: np.ndarray, by: Callable) -> np.ndarray:
    """``sort(v; by = by)``."""
    v = np.asarray(v)
    return v[jl_sortperm_stable(by(v))] if len(v) else v.copy()


def sort_split_points(sorted_lp: np.ndarray, boundaries: np.ndarray, by: Optional[Callable] = None) -> List[int]:
    """The scan of scatter_n_sort_localparts (src/sort.jl:26-50): piece i = sorted[p_sorted : first x with by(x) > by(boundaries[i+1]))."""
    ends, p = [], 0
    n = len(sorted_lp)
    keys = sorted_lp if by is None else by(np.asarray(sorted_lp))
    bkeys = boundaries if by is None else by(np.asarray(boundaries))
    for i in range(len(boundaries) - 1):
        with np.errstate(invalid="ignore"):
            gt = keys[p:] > bkeys[i + 1]
        p_till = p + int(np.argmax(gt)) if gt.any() else n
        ends.append(p_till)
        p = p_till
    return ends


def darray_sort(d: ODArray, sample=True, by: Optional[Callable] = None):
    """``sort(d::DVector; sample, by)`` (src/sort.jl:107-170).  Returns (ODArray of the sorted vector, boundaries).  ``by`` is a
    NumPy-vectorised key function: local sorts and the sort of the gathered samples order by ``by(x)`` (``kwargs...`` at :8, :22, :61,
    :77), an explicit ``sample`` array is sorted WITHOUT it (:148), and the split compares ``by(x) > by(boundaries[i+1])`` (:32).
    Pieces are appended in source order (the reference appends in arrival order, which is not deterministic).

    Reference behaviour kept as it is: the scan hands out ``np`` pieces and whatever follows the last split point is sent to NOBODY
    (:26-50).  Without ``by`` the last boundary is ``typemax(T)``, nothing exceeds it and the last piece runs to the end; with a key
    function for which ``by(typemax(T))`` is not the largest key (``x -> -x``, ``x -> rem(x, 7)``) the elements whose key exceeds
    ``by(typemax(T))`` are dropped from the result -- the restatement (and the product) drop exactly the same elements."""
    nparts = len(d.pids)
    dt = d.chunks[0].dtype
    lsort = jl_sort if by is None else (lambda v: jl_sort_by(v, by))
    srt = [lsort(c) for c in d.chunks]
    if sample is True:
        samples = np.concatenate([s[list(sort_sample_indices(len(s)))] for s in srt])
        if by is None:
            boundaries = sort_boundaries_from_samples(samples, nparts, dt)
        else:
            s = jl_sort_by(samples.astype(dt), by).copy()
            s[0] = _typemin(dt)
            step = len(s) // nparts
            boundaries = np.asarray([s[(x - 1) * step] for x in range(1, nparts + 1)] + [_typemax(dt)], dtype=dt)
    else:
        if sample is False:
            lo = min(c.min() for c in d.chunks)
            hi = max(c.max() for c in d.chunks)
            sample = (lo, hi)
        if isinstance(sample, tuple):
            sample = sort_uniform_sample(sample[0], sample[1], nparts, dt)
        boundaries = sort_boundaries_from_samples(np.asarray(sample), nparts, dt)
    recv = [[] for _ in range(nparts)]
    for s in srt:
        p = 0
        for i, e in enumerate(sort_split_points(s, boundaries, by)):
            recv[i].append(s[p:e])
            p = e
    parts = [lsort(np.concatenate(r)) for r in recv]
    keep = [i for i, p in enumerate(parts) if len(p) > 0]       # zero-length parts are dropped (src/sort.jl:163-168)
    if not keep:
        raise ValueError("ArgumentError: sort left no non-empty part (DArray(refs) of an empty list of refs)")
    sizes = [len(parts[i]) for i in keep]
    starts = np.concatenate([[1], 1 + np.cumsum(sizes)]).astype(int)
    out = ODArray((int(sum(sizes)),), (len(keep),), [d.pids[i] for i in keep],
                  [((int(starts[k]), int(starts[k + 1] - 1)),) for k in range(len(keep))], [list(map(int, starts))],
                  [parts[i] for i in keep])
    return out, boundaries
