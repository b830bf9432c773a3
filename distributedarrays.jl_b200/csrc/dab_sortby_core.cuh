// dab_sortby_core.cuh -- the per-element arithmetic of dab_sort_by_key (dab_sortby.cu) as __host__ __device__ functions, so that the very
// same code runs inside the kernels and inside tools/sortby_host_check.cu (a host-only replay of the whole composition against
// std::stable_sort; built and run by the CPU test tier).
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>

#include "dab_sort_key.cuh"

constexpr unsigned long long DAB_SORTBY_SIGN64 = 0x8000000000000000ull;

// radix key of one by-value: the keys-only bijection, except that all NaNs collapse to the largest key (isless(NaN, NaN) is false both
// ways: NaN keys are ties, and ties keep input order).  The largest key of the bijection is itself a NaN, so nothing else maps there.
template <typename KT>
__host__ __device__ inline typename SortKey<KT>::U sortby_radix_key(typename SortKey<KT>::U raw) {
    using U = typename SortKey<KT>::U;
    if constexpr (std::is_floating_point<KT>::value) {
        constexpr U ABS = (U)~((U)1 << (8 * sizeof(U) - 1));
        constexpr U INF = sizeof(U) == 4 ? (U)0x7F800000u : (U)0x7FF0000000000000ull;
        if ((raw & ABS) > INF) return (U)~(U)0;
    }
    return SortKey<KT>::enc(raw);
}

// word j of a round: half of radix_key(keys[i]) << 32 | j, with i = j (round 1) or i = lo32(prev[j]) (round 2: the order round 1 left).
// Stored with the top bit flipped: dab_sort orders Int64 words as SIGNED integers.
template <typename KT>
__host__ __device__ inline unsigned long long sortby_word(const typename SortKey<KT>::U* keys, const unsigned long long* prev, int half, size_t j) {
    using U = typename SortKey<KT>::U;
    const size_t i = prev ? (size_t)(unsigned int)prev[j] : j;
    const U e = sortby_radix_key<KT>(keys[i]);
    unsigned int h;
    if constexpr (sizeof(U) == 4) h = (unsigned int)e;
    else h = half ? (unsigned int)(e >> 32) : (unsigned int)e;
    return (((unsigned long long)h << 32) | (unsigned long long)j) ^ DAB_SORTBY_SIGN64;
}

// position in vals of the element that ends up at j: lo32(last[j]) for one round, lo32(first[lo32(last[j])]) for two
__host__ __device__ inline unsigned int sortby_source(const unsigned long long* last, const unsigned long long* first, size_t j) {
    const unsigned int a = (unsigned int)last[j];
    return first ? (unsigned int)first[a] : a;
}

// 64-bit keys take two rounds (least-significant half first), 32-bit keys one
inline int sortby_rounds(int32_t key_bytes) { return key_bytes == 8 ? 2 : 1; }
