// dab_sortby.cu -- `sort(lp; by = f)` of one chunk: the keyed form of the two local sorts of the reference's samplesort
// (`sort(lp; kwargs...)` src/sort.jl:8, `sort(localpart(d); by = by, kwargs...)` :22, `sort!(lp_sorting; by = by, kwargs...)` :61).
//
// The caller evaluates keys = f.(lp) with the fused broadcast kernel (any traced closure); this file orders the VALUES by those keys,
// stably (Julia's default algorithm for sort(v; by) is stable, so elements with equal keys keep their input order), in Julia's `isless`
// order of the keys (-0.0 < +0.0, every NaN key equal to every other NaN key and after +Inf).
//
// Composition instead of a second radix-sort kernel: a 32-bit radix key and the 32-bit position of the element form ONE 64-bit word
//     word[j] = radix_key(keys[j]) << 32 | j
// and K11 (dab_sort on Int64, the onesweep kernel that is measured and tuned) sorts the words: the position in the low half makes every
// word unique, keeps the sort stable whatever the kernel does with ties, and IS the permutation afterwards.  64-bit keys take two rounds,
// least-significant half first (the textbook LSD argument: round 2 is stable with respect to the order round 1 left):
//     round 1   word[j] = lo32(key[j])      << 32 | j          sort -> S1      p1[j] = lo32(S1[j])
//     round 2   word[j] = hi32(key[p1[j]])  << 32 | j          sort -> S2      perm[j] = p1[lo32(S2[j])]
// Algorithmic bytes per element (k = key bytes, v = value bytes, w = 8-byte word, P = digit passes of the word sort, 8 unless constant
// digits are skipped):  k + w  (pack)  +  w * (1 + 2 P)  (K11)  +  w + 2 v  (gather; the read of vals is a random 4/8-byte access);
// twice the first two terms for 64-bit keys.  Data movement for a by-key order, not a bandwidth path: not tuned beyond coalesced streams.
#include <type_traits>

#include "dab_common.cuh"
#include "dab_sortby_core.cuh"

namespace {

template <typename KT>
__global__ void __launch_bounds__(256) sortby_pack_kernel(const typename SortKey<KT>::U* __restrict__ keys, const unsigned long long* __restrict__ prev,
                                                          int half, unsigned long long* __restrict__ words, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) words[j] = sortby_word<KT>(keys, prev, half, j);
}

// out[j] = vals[perm[j]]
template <typename V>
__global__ void __launch_bounds__(256) sortby_gather_kernel(const V* __restrict__ vals, const unsigned long long* __restrict__ last,
                                                            const unsigned long long* __restrict__ first, V* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) out[j] = vals[sortby_source(last, first, j)];
}

inline size_t word_stride_bytes(size_t n) { return (n * 8 + 255) & ~(size_t)255; }   // every word buffer starts 256-byte aligned

inline int rounds_of(int32_t key_dtype) { return sortby_rounds((key_dtype == DAB_F64 || key_dtype == DAB_I64) ? 8 : 4); }

template <typename KT>
int32_t pack_t(dab_ctx* ctx, const void* keys, const unsigned long long* prev, int half, unsigned long long* words, size_t n) {
    using U = typename SortKey<KT>::U;
    const int grid = dab_grid_for(ctx, (n + 255) / 256, 8);
    sortby_pack_kernel<KT><<<grid, 256, 0, ctx->stream>>>((const U*)keys, prev, half, words, n);
    DAB_LAUNCHED(ctx);
    return DAB_OK;
}

int32_t pack(dab_ctx* ctx, int32_t key_dtype, const void* keys, const unsigned long long* prev, int half, unsigned long long* words, size_t n) {
    switch (key_dtype) {
        case DAB_F32: return pack_t<float>(ctx, keys, prev, half, words, n);
        case DAB_F64: return pack_t<double>(ctx, keys, prev, half, words, n);
        case DAB_I32: return pack_t<int32_t>(ctx, keys, prev, half, words, n);
        case DAB_I64: return pack_t<int64_t>(ctx, keys, prev, half, words, n);
        default: return dab_fail(ctx, DAB_ERR_UNSUPPORTED, "dab_sort_by_key: key dtype %d", key_dtype);
    }
}

template <typename V>
int32_t gather_t(dab_ctx* ctx, const void* vals, const unsigned long long* last, const unsigned long long* first, void* out, size_t n) {
    const int grid = dab_grid_for(ctx, (n + 255) / 256, 8);
    sortby_gather_kernel<V><<<grid, 256, 0, ctx->stream>>>((const V*)vals, last, first, (V*)out, n);
    DAB_LAUNCHED(ctx);
    return DAB_OK;
}

}  // namespace

extern "C" {

int32_t dab_sort_by_key_scratch_bytes(int32_t key_dtype, size_t n, size_t* bytes) {
    if (bytes == nullptr) return dab_fail(nullptr, DAB_ERR_ARG, "dab_sort_by_key_scratch_bytes: null pointer");
    if (key_dtype != DAB_F32 && key_dtype != DAB_F64 && key_dtype != DAB_I32 && key_dtype != DAB_I64)
        return dab_fail(nullptr, DAB_ERR_UNSUPPORTED, "dab_sort_by_key: key dtype %d", key_dtype);
    *bytes = (size_t)(2 + rounds_of(key_dtype)) * word_stride_bytes(n);   // packed words, the radix sort's scratch, one sorted buffer per round
    return DAB_OK;
}

int32_t dab_sort_by_key(dab_ctx* ctx, int32_t key_dtype, const void* keys, int32_t val_bytes, const void* vals, void* vals_out, void* scratch,
                        size_t scratch_bytes, size_t n) {
    DAB_ENTER(ctx);
    if (n == 0) return DAB_OK;
    DAB_REQUIRE(ctx, keys && vals && vals_out && scratch, DAB_ERR_ARG, "dab_sort_by_key: null pointer");
    DAB_REQUIRE(ctx, vals != vals_out, DAB_ERR_ARG, "dab_sort_by_key: vals_out must not alias vals");
    DAB_REQUIRE(ctx, val_bytes == 4 || val_bytes == 8, DAB_ERR_UNSUPPORTED, "dab_sort_by_key: values of %d bytes", val_bytes);
    DAB_REQUIRE(ctx, n < 0xFFFFF000ull, DAB_ERR_UNSUPPORTED, "dab_sort_by_key: chunks of 2^32 or more elements are not served");
    size_t need = 0;
    int32_t st = dab_sort_by_key_scratch_bytes(key_dtype, n, &need);
    if (st != DAB_OK) return dab_fail(ctx, st, "dab_sort_by_key: key dtype %d", key_dtype);
    DAB_REQUIRE(ctx, scratch_bytes >= need, DAB_ERR_ARG, "dab_sort_by_key: scratch of %zu bytes, %zu needed", scratch_bytes, need);
    DAB_REQUIRE(ctx, ((uintptr_t)scratch & 15) == 0, DAB_ERR_ARG, "dab_sort_by_key: scratch must be 16-byte aligned");

    const size_t ws = word_stride_bytes(n);
    unsigned long long* words = (unsigned long long*)scratch;
    unsigned long long* tmp = (unsigned long long*)((char*)scratch + ws);
    unsigned long long* s1 = (unsigned long long*)((char*)scratch + 2 * ws);
    unsigned long long* s2 = (unsigned long long*)((char*)scratch + 3 * ws);   // only touched for 64-bit keys
    const int rounds = rounds_of(key_dtype);

    st = pack(ctx, key_dtype, keys, nullptr, 0, words, n);
    if (st != DAB_OK) return st;
    st = dab_sort(ctx, DAB_I64, words, s1, tmp, n);
    if (st != DAB_OK) return st;
    const unsigned long long* last = s1;
    const unsigned long long* first = nullptr;
    if (rounds == 2) {
        st = pack(ctx, key_dtype, keys, s1, 1, words, n);
        if (st != DAB_OK) return st;
        st = dab_sort(ctx, DAB_I64, words, s2, tmp, n);
        if (st != DAB_OK) return st;
        last = s2;
        first = s1;
    }
    return val_bytes == 4 ? gather_t<uint32_t>(ctx, vals, last, first, vals_out, n) : gather_t<uint64_t>(ctx, vals, last, first, vals_out, n);
}

}  // extern "C"
