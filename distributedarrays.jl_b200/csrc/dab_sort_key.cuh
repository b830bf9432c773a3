// dab_sort_key.cuh -- the order-preserving bijection "raw bits of a key <-> unsigned radix key" shared by the keys-only sort
// (dab_sort.cu) and the sort-by-key composition (dab_sortby.cu).  Julia's `isless` order: integers by value; floats with
// -0.0 < +0.0 and every NaN after +Inf.
#pragma once
#include <cstdint>

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
