// dab_sort.cu -- K11: sort of one chunk, the `sort(localpart(d))` / `sort!(lp_sorting)` steps of the reference's samplesort
// (src/sort.jl:3-15, 18-63), and the split-point search of scatter_n_sort_localparts (:28-40).
//
// Keys only, ascending in Julia's `isless` order: integers by value; floats with -0.0 < +0.0 and every NaN after +Inf (bit
// patterns preserved; the reference keeps NaNs in their original relative order, here they are ordered by payload).
//
// Algorithm: least-significant-digit radix sort, 8-bit digits, hand-written for sm_100a.  HBM-bound integer work:
//   sort_hist_kernel    one read of the keys -> the 256-bin histogram of EVERY digit position (global atomics on per-CTA shared
//                       histograms).  The host reads it back (16 KiB) and drops the passes whose digit is constant -- Int64 data in a
//                       small range needs 2-3 of 8 passes -- and derives each pass's bucket bases.
//   per remaining pass:
//   sort_count_kernel   per-tile digit counts                               (read n keys)
//   sort_scan_kernel    exclusive scan of the counts along the tiles, one CTA per digit value
//   sort_scatter_kernel stable multi-split of each tile and scatter          (read n keys, write n keys)
//                       a warp owns a contiguous run of the tile; ballots group equal digits inside each 32-key step,
//                       per-warp shared counters carry the running rank, so equal digits keep their input order (LSD needs it).
// Algorithmic bytes: elem * (1 + 3 * passes) per key.  Tiles are 256 threads x 8 keys.
#include <type_traits>

#include "dab_common.cuh"

namespace {

constexpr int ST_THREADS = 256;
constexpr int ST_KPT = 8;                        // keys per thread (16 left the scatter at 111 registers = 2 CTAs per SM, latency-bound)
constexpr int ST_TILE = ST_THREADS * ST_KPT;     // 2048 keys per CTA
constexpr int ST_WARPS = ST_THREADS / 32;

// ---- order-preserving bijection raw bits <-> unsigned key ----------------------------------------------------------------------
template <typename T> struct SortKey;
template <> struct SortKey<int32_t> {
    using U = uint32_t;
    static constexpr int DIGITS = 4;
    __host__ __device__ static U enc(U u) { return u ^ 0x80000000u; }
    __host__ __device__ static U dec(U k) { return k ^ 0x80000000u; }
};
template <> struct SortKey<int64_t> {
    using U = uint64_t;
    static constexpr int DIGITS = 8;
    __host__ __device__ static U enc(U u) { return u ^ 0x8000000000000000ull; }
    __host__ __device__ static U dec(U k) { return k ^ 0x8000000000000000ull; }
};
// floats: negatives reversed below the positives (so -0.0 < +0.0), then rotated down by C so that -Inf is key 0 and the
// sign-bit NaNs (which the reversal put below -Inf) wrap around to the very top, above the positive NaNs: NaNs last, bijective.
template <> struct SortKey<float> {
    using U = uint32_t;
    static constexpr int DIGITS = 4;
    static constexpr U C = 0x007FFFFFu;
    __host__ __device__ static U enc(U u) { return ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) - C; }
    __host__ __device__ static U dec(U k) {
        k += C;
        return (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    }
};
template <> struct SortKey<double> {
    using U = uint64_t;
    static constexpr int DIGITS = 8;
    static constexpr U C = 0x000FFFFFFFFFFFFFull;
    __host__ __device__ static U enc(U u) { return ((u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull)) - C; }
    __host__ __device__ static U dec(U k) {
        k += C;
        return (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
    }
};

struct SortBases { uint32_t b[256]; };

// Lanes of the warp whose 8-bit digit equals mine (dg = 256 marks "no key"; those lanes group together): 9 ballots.  On sm_100a
// __match_any_sync costs one round per DISTINCT value in the warp (measured ~45 clk per warp-step on random digits); the bitwise
// form is flat.
__device__ __forceinline__ unsigned int match_digit(unsigned int dg) {
    unsigned int m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 9; ++b) {
        const bool bit = (dg >> b) & 1u;
        const unsigned int bal = __ballot_sync(0xffffffffu, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}

// A thread's ST_KPT CONSECUTIVE keys (blocked arrangement: 128 contiguous bytes of 8-byte keys), 16-byte loads when the tile is full
// and aligned.  Returns how many of them exist.  Used by the counting kernels: runs of equal digits (sorted or narrow-range input)
// collapse into one shared-memory atomic per run instead of serialising a whole warp on one bin.
template <typename U>
__device__ __forceinline__ int load_blocked(const U* __restrict__ in, size_t first, size_t n, U (&key)[ST_KPT]) {
    if (first + ST_KPT <= n && (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
        constexpr int PER = 16 / sizeof(U);
        const int4* p = reinterpret_cast<const int4*>(in + first);
#pragma unroll
        for (int q = 0; q < ST_KPT / PER; ++q) {
            const int4 v = __ldcs(p + q);
            memcpy(&key[q * PER], &v, 16);
        }
        return ST_KPT;
    }
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < ST_KPT; ++k)
        if (first + k < n) {
            key[k] = __ldcs(in + first + k);
            cnt = k + 1;
        }
    return cnt;
}

// ---- all-digit histogram ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(ST_THREADS) sort_hist_kernel(const typename SortKey<T>::U* __restrict__ in, size_t n,
                                                               unsigned long long* __restrict__ ghist) {
    using K = SortKey<T>;
    using U = typename K::U;
    __shared__ unsigned int sh[K::DIGITS][256];
    for (int i = threadIdx.x; i < K::DIGITS * 256; i += ST_THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();
    // grid-stride over tiles: a few CTAs per SM accumulate privately, so the global histogram sees gridDim.x flushes, not n / 4096
    for (size_t base = (size_t)blockIdx.x * ST_TILE; base < n; base += (size_t)gridDim.x * ST_TILE) {
        U key[ST_KPT];
        const int cnt = load_blocked<U>(in, base + (size_t)threadIdx.x * ST_KPT, n, key);
        if (cnt > 0) {
            unsigned int cur[K::DIGITS], run[K::DIGITS];
            const U k0 = K::enc(key[0]);
#pragma unroll
            for (int d = 0; d < K::DIGITS; ++d) {
                cur[d] = (unsigned)(k0 >> (8 * d)) & 255u;
                run[d] = 1;
            }
#pragma unroll
            for (int k = 1; k < ST_KPT; ++k)
                if (k < cnt) {
                    const U kk = K::enc(key[k]);
#pragma unroll
                    for (int d = 0; d < K::DIGITS; ++d) {
                        const unsigned int dg = (unsigned)(kk >> (8 * d)) & 255u;
                        if (dg == cur[d]) {
                            ++run[d];
                        } else {
                            atomicAdd(&sh[d][cur[d]], run[d]);
                            cur[d] = dg;
                            run[d] = 1;
                        }
                    }
                }
#pragma unroll
            for (int d = 0; d < K::DIGITS; ++d) atomicAdd(&sh[d][cur[d]], run[d]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K::DIGITS * 256; i += ST_THREADS) {
        const unsigned int c = (&sh[0][0])[i];
        if (c) atomicAdd(ghist + i, (unsigned long long)c);
    }
}

// ---- per-tile counts of one digit --------------------------------------------------------------------------------------------------
template <typename T, bool RAW>
__global__ void __launch_bounds__(ST_THREADS) sort_count_kernel(const typename SortKey<T>::U* __restrict__ in, size_t n, int shift,
                                                                unsigned int nblocks, unsigned int* __restrict__ counts) {
    using K = SortKey<T>;
    using U = typename K::U;
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    U key[ST_KPT];
    const int cnt = load_blocked<U>(in, (size_t)blockIdx.x * ST_TILE + (size_t)threadIdx.x * ST_KPT, n, key);
    if (cnt > 0) {
        unsigned int cur = (unsigned)((RAW ? K::enc(key[0]) : key[0]) >> shift) & 255u, run = 1;
#pragma unroll
        for (int k = 1; k < ST_KPT; ++k)
            if (k < cnt) {
                const unsigned int dg = (unsigned)((RAW ? K::enc(key[k]) : key[k]) >> shift) & 255u;
                if (dg == cur) {
                    ++run;
                } else {
                    atomicAdd(&sh[cur], run);
                    cur = dg;
                    run = 1;
                }
            }
        atomicAdd(&sh[cur], run);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];   // digit-major: row d holds the tiles' counts of digit d
}

// ---- exclusive scan of each digit's row over the tiles (one CTA per digit value) -----------------------------------------------------
__global__ void __launch_bounds__(1024) sort_scan_kernel(unsigned int* __restrict__ counts, unsigned int nblocks) {
    __shared__ unsigned int wsum[32];
    __shared__ unsigned int carry_s;
    unsigned int* row = counts + (size_t)blockIdx.x * nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (unsigned int base = 0; base < nblocks; base += 1024) {
        const unsigned int i = base + threadIdx.x;
        const unsigned int v = i < nblocks ? row[i] : 0;
        unsigned int inc = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, s);
            if (lane >= s) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned int w = wsum[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, w, s);
                if (lane >= s) w += t;
            }
            wsum[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const unsigned int carry = carry_s;
        const unsigned int excl = carry + (warp ? wsum[warp - 1] : 0) + inc - v;
        if (i < nblocks) row[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
}

// ---- stable scatter of one digit -------------------------------------------------------------------------------------------------------
template <typename T, bool RAW_IN, bool RAW_OUT>
__global__ void __launch_bounds__(ST_THREADS, 4) sort_scatter_kernel(const typename SortKey<T>::U* __restrict__ in,
                                                                  typename SortKey<T>::U* __restrict__ out, size_t n, int shift,
                                                                  unsigned int nblocks, const unsigned int* __restrict__ offsets,
                                                                  SortBases bases) {
    using K = SortKey<T>;
    using U = typename K::U;
    __shared__ unsigned int wc[ST_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ST_WARPS * 256; i += ST_THREADS) (&wc[0][0])[i] = 0;
    __syncthreads();
    // warp `warp` owns keys [base + warp*KPT*32, +KPT*32) of the tile; step k takes 32 consecutive keys
    const size_t wbase = (size_t)blockIdx.x * ST_TILE + (size_t)warp * (ST_KPT * 32);
    const unsigned int lt = (1u << lane) - 1u;
    U key[ST_KPT];
    unsigned short rank[ST_KPT];
#pragma unroll
    for (int k = 0; k < ST_KPT; ++k) {
        const size_t i = wbase + (size_t)k * 32 + lane;
        key[k] = i < n ? __ldcs(in + i) : U(0);
    }
    // Running per-warp digit counters.  Equal digits inside a 32-key step are grouped by ballots; the group's lowest lane bumps the
    // counter once and hands the old value round.  Steps are issued in order by the one warp that owns this counter row, so equal
    // digits keep their input order.
#pragma unroll
    for (int k = 0; k < ST_KPT; ++k) {
        const size_t i = wbase + (size_t)k * 32 + lane;
        const bool valid = i < n;
        if (RAW_IN) key[k] = K::enc(key[k]);
        const unsigned int dg = valid ? ((unsigned)(key[k] >> shift) & 255u) : 256u;   // invalid lanes form their own group
        const unsigned int grp = match_digit(dg);
        const unsigned int before = __popc(grp & lt);
        const int leader = __ffs(grp) - 1;
        unsigned int old = 0;
        if (valid && before == 0) old = atomicAdd(&wc[warp][dg], (unsigned int)__popc(grp));
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = (unsigned short)(old + before);
    }
    __syncthreads();
    {   // digit d: global base of the bucket + this tile's offset inside it, then running prefix over the warps
        const int d = threadIdx.x;
        unsigned int run = bases.b[d] + offsets[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < ST_WARPS; ++w) {
            const unsigned int c = wc[w][d];
            wc[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ST_KPT; ++k) {
        const size_t i = wbase + (size_t)k * 32 + lane;
        if (i < n) {
            const unsigned int dg = (unsigned)(key[k] >> shift) & 255u;
            const unsigned int pos = wc[warp][dg] + rank[k];
            out[pos] = RAW_OUT ? K::dec(key[k]) : key[k];
        }
    }
}

// encode / decode a whole buffer (only when no digit pass runs at all, or as the odd-parity fix-up never needed: kept for n small)
template <typename T>
__global__ void sort_small_kernel(const typename SortKey<T>::U* __restrict__ in, typename SortKey<T>::U* __restrict__ out, unsigned int n) {
    // n <= 1024: rank sort in shared memory by one CTA (stable: ties broken by index)
    using K = SortKey<T>;
    using U = typename K::U;
    __shared__ U sk[1024];
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) sk[i] = K::enc(in[i]);
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
        const U me = sk[i];
        unsigned int r = 0;
        for (unsigned int j = 0; j < n; ++j) {
            const U o = sk[j];
            r += (o < me) || (o == me && j < i);
        }
        out[r] = K::dec(me);
    }
}

int32_t sort_scratch(dab_ctx* ctx, size_t dev_bytes) {
    if (!ctx->sort_host) DAB_CUDA(ctx, cudaMallocHost(&ctx->sort_host, 8 * 256 * sizeof(unsigned long long) + 4096));
    if (ctx->sort_dev_bytes < dev_bytes) {
        if (ctx->sort_dev) {
            DAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            DAB_CUDA(ctx, cudaFree(ctx->sort_dev));
            ctx->sort_dev = nullptr;
            ctx->sort_dev_bytes = 0;
        }
        DAB_CUDA(ctx, cudaMalloc(&ctx->sort_dev, dev_bytes));
        ctx->sort_dev_bytes = dev_bytes;
    }
    return DAB_OK;
}

template <typename T>
int32_t sort_t(dab_ctx* ctx, const void* in_v, void* out_v, void* tmp_v, size_t n) {
    using K = SortKey<T>;
    using U = typename K::U;
    const U* in = (const U*)in_v;
    U* out = (U*)out_v;
    U* tmp = (U*)tmp_v;
    if (n == 0) return DAB_OK;
    if (n <= 1024) {
        const U* src = in;
        if (in == out) {  // the rank sort is not in-place
            DAB_REQUIRE(ctx, tmp != nullptr, DAB_ERR_ARG, "dab_sort: in-place sort needs tmp");
            DAB_CUDA(ctx, cudaMemcpyAsync(tmp, in, n * sizeof(U), cudaMemcpyDeviceToDevice, ctx->stream));
            src = tmp;
        }
        sort_small_kernel<T><<<1, 256, 0, ctx->stream>>>(src, out, (unsigned)n);
        DAB_LAUNCHED(ctx);
        return DAB_OK;
    }
    DAB_REQUIRE(ctx, tmp != nullptr && tmp != out && tmp != in, DAB_ERR_ARG, "dab_sort: tmp must be a distinct buffer of n elements");
    DAB_REQUIRE(ctx, n < 0xFFFFF000ull, DAB_ERR_UNSUPPORTED, "dab_sort: chunks of 2^32 or more elements are not served");
    const unsigned int nblocks = (unsigned int)((n + ST_TILE - 1) / ST_TILE);
    const size_t hist_bytes = (size_t)K::DIGITS * 256 * sizeof(unsigned long long);
    const size_t counts_bytes = (size_t)256 * nblocks * sizeof(unsigned int);
    {
        int32_t st = sort_scratch(ctx, 16384 + counts_bytes);
        if (st != DAB_OK) return st;
    }
    unsigned long long* ghist = (unsigned long long*)ctx->sort_dev;
    unsigned int* counts = (unsigned int*)((char*)ctx->sort_dev + 16384);
    unsigned long long* hhist = (unsigned long long*)ctx->sort_host;
    DAB_CUDA(ctx, cudaMemsetAsync(ghist, 0, hist_bytes, ctx->stream));
    {
        const unsigned int hgrid = nblocks < (unsigned)ctx->sm_count * 4u ? nblocks : (unsigned)ctx->sm_count * 4u;
        sort_hist_kernel<T><<<hgrid, ST_THREADS, 0, ctx->stream>>>(in, n, ghist);
    }
    DAB_LAUNCHED(ctx);
    DAB_CUDA(ctx, cudaMemcpyAsync(hhist, ghist, hist_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int active[8], na = 0;
    for (int d = 0; d < K::DIGITS; ++d) {
        bool constant = false;
        for (int b = 0; b < 256; ++b)
            if (hhist[d * 256 + b] == n) constant = true;
        if (!constant) active[na++] = d;
    }
    if (na == 0) {  // every key equal
        if (in != out) DAB_CUDA(ctx, cudaMemcpyAsync(out, in, n * sizeof(U), cudaMemcpyDeviceToDevice, ctx->stream));
        return DAB_OK;
    }
    // ping-pong so that the LAST pass writes `out`; `in` is never written unless it is `out`
    const U* src = in;
    for (int p = 0; p < na; ++p) {
        const int d = active[p];
        const bool first = (p == 0), last = (p == na - 1);
        U* dst = ((na - 1 - p) % 2 == 0) ? out : tmp;
        if (dst == (U*)src) {
            // only possible on the first pass of an in-place sort with an odd number of passes: stage the input in tmp
            DAB_CUDA(ctx, cudaMemcpyAsync(tmp, src, n * sizeof(U), cudaMemcpyDeviceToDevice, ctx->stream));
            src = tmp;
        }
        SortBases bases;
        unsigned long long run = 0;
        for (int b = 0; b < 256; ++b) {
            bases.b[b] = (uint32_t)run;
            run += hhist[d * 256 + b];
        }
        const int shift = 8 * d;
        if (first) sort_count_kernel<T, true><<<nblocks, ST_THREADS, 0, ctx->stream>>>(src, n, shift, nblocks, counts);
        else sort_count_kernel<T, false><<<nblocks, ST_THREADS, 0, ctx->stream>>>(src, n, shift, nblocks, counts);
        DAB_LAUNCHED(ctx);
        sort_scan_kernel<<<256, 1024, 0, ctx->stream>>>(counts, nblocks);
        DAB_LAUNCHED(ctx);
        if (first && last) sort_scatter_kernel<T, true, true><<<nblocks, ST_THREADS, 0, ctx->stream>>>(src, dst, n, shift, nblocks, counts, bases);
        else if (first) sort_scatter_kernel<T, true, false><<<nblocks, ST_THREADS, 0, ctx->stream>>>(src, dst, n, shift, nblocks, counts, bases);
        else if (last) sort_scatter_kernel<T, false, true><<<nblocks, ST_THREADS, 0, ctx->stream>>>(src, dst, n, shift, nblocks, counts, bases);
        else sort_scatter_kernel<T, false, false><<<nblocks, ST_THREADS, 0, ctx->stream>>>(src, dst, n, shift, nblocks, counts, bases);
        DAB_LAUNCHED(ctx);
        src = dst;
    }
    return DAB_OK;
}

// ---- split points in a sorted chunk ----------------------------------------------------------------------------------------------------
// counts[i] = number of elements x with NOT (x > bounds[i]) counted from the front of the sorted chunk up to the first x > bounds[i]
// -- i.e. the p_till - 1 of the reference's scan (src/sort.jl:31-38) had it started at element 1: the number of non-NaN elements
// <= bounds[i] (IEEE compare, so -0.0 == +0.0), or n when no element exceeds the bound (NaNs compare false and stay in the piece).
template <typename T>
__global__ void sort_bounds_kernel(const typename SortKey<T>::U* __restrict__ sorted, size_t n, const typename SortKey<T>::U* __restrict__ bounds,
                                   int nb, unsigned long long* __restrict__ counts) {
    using K = SortKey<T>;
    using U = typename K::U;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb) return;
    U braw = bounds[t];
    bool nan_bound = false;
    if constexpr (sizeof(T) == 4 && !std::is_integral<T>::value) {
        nan_bound = (braw & 0x7FFFFFFFu) > 0x7F800000u;
        if ((braw & 0x7FFFFFFFu) == 0) braw = 0;  // -0.0 bounds like +0.0
    }
    if constexpr (sizeof(T) == 8 && !std::is_integral<T>::value) {
        nan_bound = (braw & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull;
        if ((braw & 0x7FFFFFFFFFFFFFFFull) == 0) braw = 0;
    }
    if (nan_bound) {  // x > NaN is never true
        counts[t] = n;
        return;
    }
    const U kb = K::enc(braw);
    size_t lo = 0, hi = n;  // first index whose key > kb
    while (lo < hi) {
        const size_t mid = lo + ((hi - lo) >> 1);
        if (K::enc(sorted[mid]) <= kb) lo = mid + 1;
        else hi = mid;
    }
    // everything from lo on is either > bound or NaN; if it is all NaN no element exceeds the bound and the scan runs to the end
    bool rest_nan = false;
    if constexpr (!std::is_integral<T>::value) {
        if (lo < n) {
            const U u = sorted[lo];  // smallest remaining key: NaN iff all remaining are NaN (NaNs sort last)
            if constexpr (sizeof(T) == 4) rest_nan = (u & 0x7FFFFFFFu) > 0x7F800000u;
            else rest_nan = (u & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull;
        }
    }
    counts[t] = rest_nan ? n : lo;
}

template <typename T>
int32_t bounds_t(dab_ctx* ctx, const void* sorted, size_t n, const void* bounds_host, int nb, unsigned long long* counts_host) {
    using U = typename SortKey<T>::U;
    DAB_REQUIRE(ctx, nb >= 1 && nb <= 256, DAB_ERR_ARG, "dab_sorted_split: 1..256 bounds");
    int32_t st = sort_scratch(ctx, 16384);
    if (st != DAB_OK) return st;
    U* dbounds = (U*)ctx->sort_dev;                                             // [0, 2 KiB)
    unsigned long long* dcounts = (unsigned long long*)((char*)ctx->sort_dev + 4096);
    unsigned long long* hc = (unsigned long long*)((char*)ctx->sort_host + 8 * 256 * sizeof(unsigned long long));
    DAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                         // the staging areas may still be in use by a sort
    memcpy(hc, bounds_host, (size_t)nb * sizeof(U));
    DAB_CUDA(ctx, cudaMemcpyAsync(dbounds, hc, (size_t)nb * sizeof(U), cudaMemcpyHostToDevice, ctx->stream));
    sort_bounds_kernel<T><<<(nb + 63) / 64, 64, 0, ctx->stream>>>((const U*)sorted, n, dbounds, nb, dcounts);
    DAB_LAUNCHED(ctx);
    DAB_CUDA(ctx, cudaMemcpyAsync(hc, dcounts, (size_t)nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    DAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(counts_host, hc, (size_t)nb * sizeof(unsigned long long));
    return DAB_OK;
}

}  // namespace

extern "C" {

int32_t dab_sort(dab_ctx* ctx, int32_t dtype, const void* in, void* out, void* tmp, size_t n) {
    DAB_ENTER(ctx);
    DAB_REQUIRE(ctx, n == 0 || (in && out), DAB_ERR_ARG, "dab_sort: null pointer");
    switch (dtype) {
        case DAB_F32: return sort_t<float>(ctx, in, out, tmp, n);
        case DAB_F64: return sort_t<double>(ctx, in, out, tmp, n);
        case DAB_I32: return sort_t<int32_t>(ctx, in, out, tmp, n);
        case DAB_I64: return sort_t<int64_t>(ctx, in, out, tmp, n);
        default: return dab_fail(ctx, DAB_ERR_UNSUPPORTED, "dab_sort: dtype %d", dtype);
    }
}

int32_t dab_sorted_split(dab_ctx* ctx, int32_t dtype, const void* sorted, size_t n, const void* bounds_host, int32_t nb,
                         unsigned long long* counts_host) {
    DAB_ENTER(ctx);
    DAB_REQUIRE(ctx, bounds_host && counts_host && (sorted || n == 0), DAB_ERR_ARG, "dab_sorted_split: null pointer");
    if (n == 0) {
        for (int i = 0; i < nb; ++i) counts_host[i] = 0;
        return DAB_OK;
    }
    switch (dtype) {
        case DAB_F32: return bounds_t<float>(ctx, sorted, n, bounds_host, nb, counts_host);
        case DAB_F64: return bounds_t<double>(ctx, sorted, n, bounds_host, nb, counts_host);
        case DAB_I32: return bounds_t<int32_t>(ctx, sorted, n, bounds_host, nb, counts_host);
        case DAB_I64: return bounds_t<int64_t>(ctx, sorted, n, bounds_host, nb, counts_host);
        default: return dab_fail(ctx, DAB_ERR_UNSUPPORTED, "dab_sorted_split: dtype %d", dtype);
    }
}

}  // extern "C"
