// sortby_host_check.cu -- host-only replay of dab_sort_by_key (distributedarrays.jl_b200/csrc/dab_sortby.cu) with the SAME per-element code
// the kernels run (dab_sortby_core.cuh: radix key with NaN collapse, key|position word, source index), std::sort on the Int64 words in
// place of K11, against std::stable_sort in Julia's isless order.  No GPU, no kernel launch: test infrastructure for the CPU tier
// (tests/test_cpu_sort.py builds and runs it).  Exit code 0 = every case identical.
//   nvcc -std=c++17 -O2 -I distributedarrays.jl_b200/csrc -o /tmp/sortby_host_check tools/sortby_host_check.cu
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "dab_sortby_core.cuh"

template <typename KT>
static bool isless_key(KT a, KT b) {  // Julia isless: -0.0 < +0.0, NaN after everything and equal to itself
    if constexpr (std::is_floating_point<KT>::value) {
        const bool an = std::isnan(a), bn = std::isnan(b);
        if (an || bn) return !an && bn;
        if (a == b) return std::signbit(a) && !std::signbit(b);
        return a < b;
    } else {
        return a < b;
    }
}

template <typename KT>
static std::vector<uint32_t> replay(const std::vector<KT>& keys) {
    using U = typename SortKey<KT>::U;
    const size_t n = keys.size();
    const U* raw = reinterpret_cast<const U*>(keys.data());
    std::vector<unsigned long long> words(n), s1(n), s2(n);
    auto sort_words = [&](std::vector<unsigned long long>& out) {   // dab_sort(DAB_I64): ascending as SIGNED integers
        std::vector<int64_t> w(n);
        std::memcpy(w.data(), words.data(), n * 8);
        std::sort(w.begin(), w.end());
        std::memcpy(out.data(), w.data(), n * 8);
    };
    for (size_t j = 0; j < n; ++j) words[j] = sortby_word<KT>(raw, nullptr, 0, j);
    sort_words(s1);
    const unsigned long long *last = s1.data(), *first = nullptr;
    if (sortby_rounds((int32_t)sizeof(KT)) == 2) {
        for (size_t j = 0; j < n; ++j) words[j] = sortby_word<KT>(raw, s1.data(), 1, j);
        sort_words(s2);
        last = s2.data();
        first = s1.data();
    }
    std::vector<uint32_t> perm(n);
    for (size_t j = 0; j < n; ++j) perm[j] = sortby_source(last, first, j);
    return perm;
}

template <typename KT>
static int check(const char* name, std::vector<KT> keys) {
    const size_t n = keys.size();
    std::vector<uint32_t> want(n);
    std::iota(want.begin(), want.end(), 0u);
    std::stable_sort(want.begin(), want.end(), [&](uint32_t a, uint32_t b) { return isless_key<KT>(keys[a], keys[b]); });
    const std::vector<uint32_t> got = replay<KT>(keys);
    if (got != want) {
        size_t j = 0;
        while (got[j] == want[j]) ++j;
        std::printf("FAIL %s n=%zu first difference at %zu: got %u want %u\n", name, n, j, got[j], want[j]);
        return 1;
    }
    return 0;
}

template <typename KT>
static int run(const char* name, std::mt19937_64& rng) {
    int bad = 0;
    for (size_t n : {1ul, 2ul, 33ul, 1024ul, 4097ul, 100003ul}) {
        std::vector<KT> k(n);
        if constexpr (std::is_floating_point<KT>::value) {
            std::normal_distribution<double> g(0.0, 1.0);
            for (auto& v : k) v = (KT)(std::round(g(rng) * 10.0) / 10.0);                    // many ties
            const KT special[] = {(KT)NAN, -(KT)NAN, (KT)0.0, (KT)-0.0, (KT)INFINITY, -(KT)INFINITY};
            for (size_t t = 0; t < n / 8 + 1; ++t) k[rng() % n] = special[rng() % 6];
            if (n > 30) {                                                                 // NaN payloads and signs: still ONE key
                using U = typename SortKey<KT>::U;
                U a = sizeof(U) == 4 ? (U)0x7FC00123u : (U)0x7FF8000000000123ull, b = sizeof(U) == 4 ? (U)0xFFC00001u : (U)0xFFF8000000000001ull;
                std::memcpy(&k[5], &a, sizeof(U));
                std::memcpy(&k[9], &b, sizeof(U));
            }
        } else {
            for (auto& v : k) v = (KT)rng();                                               // full range
            for (size_t t = 0; t < n / 2 + 1; ++t) k[rng() % n] = (KT)7;
            if (n > 30) {
                k[0] = std::numeric_limits<KT>::min();
                k[1] = std::numeric_limits<KT>::max();
                k[2] = (KT)-1;
                k[3] = 0;
            }
        }
        bad += check<KT>(name, k);
    }
    return bad;
}

int main() {
    std::mt19937_64 rng(20260923);
    int bad = run<float>("Float32", rng) + run<double>("Float64", rng) + run<int32_t>("Int32", rng) + run<int64_t>("Int64", rng);
    // the radix key is strictly monotone over the isless order and collapses exactly the NaNs
    const float f[] = {-INFINITY, -1.5f, -0.0f, 0.0f, 1e-30f, 2.0f, INFINITY};
    for (int i = 0; i + 1 < 7; ++i) {
        uint32_t a, b;
        std::memcpy(&a, &f[i], 4);
        std::memcpy(&b, &f[i + 1], 4);
        if (!(sortby_radix_key<float>(a) < sortby_radix_key<float>(b))) { std::printf("FAIL monotone %d\n", i); ++bad; }
    }
    if (sortby_radix_key<float>(0x7FC00000u) != 0xFFFFFFFFu || sortby_radix_key<float>(0xFFC12345u) != 0xFFFFFFFFu ||
        sortby_radix_key<double>(0x7FF8000000000001ull) != ~0ull || !(sortby_radix_key<float>(0x7F800000u) < 0xFFFFFFFFu)) {
        std::printf("FAIL NaN collapse\n");
        ++bad;
    }
    std::printf(bad ? "sortby_host_check: %d FAILED\n" : "sortby_host_check: ok\n", bad);
    return bad ? 1 : 0;
}
