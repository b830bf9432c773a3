// dab_gemm.cu -- K12: the tile product of the matrix-matrix mul! (widening row f4, the one contraction on the scope list).
//
// Replaces   localpart(A) * convert(localtype(B), Bjk)   /   transpose(localpart(A)) * ...   inside
// _matmatmul!(C::DMatrix, A::DMatrix, B::AbstractMatrix, alpha, beta, tA)  (reference src/linalg.jl:189-257, the remotecall at
// :218-226).  Column-major (Julia) operands:  R[m x n] = op(A) * B,  op(A) = A (m x k) or A^T (A stored k x m),  B is k x n.
// The caller combines the tile results exactly as the reference does (scale C by beta, add!(localpart(C), R, alpha) per tile).
//
// Float32 -> gemm_tf32x3_kernel, hand-written for sm_100a:
//   * operands arrive by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B tensor maps; SASS UTMALDG) into a 3-stage shared-memory ring,
//     completion on mbarriers; out-of-range rows / columns / k are zero-filled by the TMA unit, so ragged edges need no special code
//   * four converter warps split every fp32 tile into tf32 "hi" (round-to-nearest) and "lo" (the rounded remainder) in place
//   * ONE thread issues tcgen05.mma.kind::tf32 (SASS UTCHMMA... ) with the accumulators in TENSOR MEMORY: per 8-deep k-step the three
//     products a_lo*b_hi + a_hi*b_lo + a_hi*b_hi ("3xTF32": the dropped a_lo*b_lo term is 2^-22 relative), M = N = 128
//   * two-level accumulation, because the tensor core TRUNCATES its fp32 accumulator on every tcgen05.mma (measured: ~2^-25 relative
//     bias per accumulating instruction, linear in their number -- 3e-6 after k = 256 on same-sign data): a TMEM accumulator only sums
//     KC consecutive k (default 64 = 8 instructions), the two correction products go to a SEPARATE accumulator (their truncation is
//     2^-11 smaller), and eight drain warps pull each finished pair of partial tiles out of TMEM (tcgen05.ld, SASS LDTM) and add them to
//     fp32 registers with round-to-nearest while the tensor core already works on the other accumulator set (all 512 TMEM columns)
//   * epilogue: the drain warps hold one output row per thread (TMEM lane == row), so consecutive lanes store consecutive rows of a
//     column-major C: coalesced 128-byte stores
// Everything else (Float64, Int32, Int64; Float32 operands whose base / leading dimension are not 16-byte aligned, which TMA cannot
// address) -> gemm_simt_kernel: shared-memory tiled FMA kernel, fp64 with DFMA, integers wrap like Julia's.
#include <cuda.h>

#include <map>
#include <mutex>
#include <utility>

#include "dab_common.cuh"

namespace {

// ======================================================================= generic SIMT tile kernel =====================================
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <typename T> __device__ __forceinline__ T gemm_fma(T a, T b, T c);
template <> __device__ __forceinline__ float gemm_fma<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double gemm_fma<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
template <> __device__ __forceinline__ int32_t gemm_fma<int32_t>(int32_t a, int32_t b, int32_t c) {
    return (int32_t)((uint32_t)a * (uint32_t)b + (uint32_t)c);
}
template <> __device__ __forceinline__ long long gemm_fma<long long>(long long a, long long b, long long c) {
    return (long long)((unsigned long long)a * (unsigned long long)b + (unsigned long long)c);
}

template <typename T, bool TA>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, size_t lda, const T* __restrict__ B, size_t ldb, T* __restrict__ C,
                                                        size_t ldc, size_t m, size_t n, size_t k) {
    __shared__ T As[SG_BK][SG_BM + 1];
    __shared__ T Bs[SG_BK][SG_BN + 1];
    const size_t m0 = (size_t)blockIdx.x * SG_BM, n0 = (size_t)blockIdx.y * SG_BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    T acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = T(0);
    for (size_t k0 = 0; k0 < k; k0 += SG_BK) {
#pragma unroll
        for (int q = 0; q < (SG_BM * SG_BK) / 256; ++q) {
            const int idx = tid + q * 256;
            int i, kk;
            if (TA) { kk = idx % SG_BK; i = idx / SG_BK; } else { i = idx % SG_BM; kk = idx / SG_BM; }
            const size_t gi = m0 + i, gk = k0 + kk;
            T v = T(0);
            if (gi < m && gk < k) v = TA ? A[gk + gi * lda] : A[gi + gk * lda];
            As[kk][i] = v;
        }
#pragma unroll
        for (int q = 0; q < (SG_BN * SG_BK) / 256; ++q) {
            const int idx = tid + q * 256;
            const int kk = idx % SG_BK, j = idx / SG_BK;
            const size_t gj = n0 + j, gk = k0 + kk;
            Bs[kk][j] = (gj < n && gk < k) ? B[gk + gj * ldb] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[kk][tx + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = Bs[kk][ty + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = gemm_fma<T>(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const size_t gj = n0 + ty + 16 * c;
        if (gj >= n) continue;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const size_t gi = m0 + tx + 16 * r;
            if (gi < m) C[gi + gj * ldc] = acc[r][c];
        }
    }
}

template <typename T>
int32_t launch_simt(dab_ctx* ctx, int transA, size_t m, size_t n, size_t k, const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc) {
    const size_t gx = (m + SG_BM - 1) / SG_BM, gy = (n + SG_BN - 1) / SG_BN;
    if (gx > 0x7fffffffull || gy > 65535ull) return dab_fail(ctx, DAB_ERR_UNSUPPORTED, "dab_gemm: tile grid %zu x %zu too large", gx, gy);
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (transA) gemm_simt_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, C, ldc, m, n, k);
    else gemm_simt_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, C, ldc, m, n, k);
    DAB_LAUNCHED(ctx);
    return DAB_OK;
}

// ======================================================================= tcgen05 3xTF32 kernel =======================================
constexpr int TG_M = 128, TG_N = 128, TG_K = 32, TG_STAGES = 3;
constexpr int TG_TILE_BYTES = TG_M * TG_K * 4;                   // 16 KiB: one operand tile (128 x 32 fp32)
constexpr int TG_STAGE_BYTES = 4 * TG_TILE_BYTES;                // A_hi, A_lo, B_hi, B_lo
constexpr int TG_THREADS = 448;                                  // warp 0 TMA, warp 1 MMA, warps 2-5 convert, warps 6-13 drain/epilogue
constexpr int TG_DRAIN_THREADS = 256;                            // two warps per TMEM lane quarter, 64 accumulator columns each
constexpr int TG_BAR_OFFSET = TG_STAGES * TG_STAGE_BYTES;
constexpr int TG_SMEM_BYTES = TG_BAR_OFFSET + 256 + 1024;        // + barriers + slack for the 1024-byte alignment of the swizzle atoms
constexpr int TG_TMEM_COLS = 512;                                // two sets of {main, correction} 128-column fp32 accumulators

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
// UMMA shared-memory matrix descriptor (sm_100 layout: start address, leading / stride byte offsets in 16-byte units, version 1,
// layout type 2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2u) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B (16-byte atoms), 1 = SWIZZLE_128B with 32-byte atoms
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {   // 32 lanes x 16 consecutive fp32 columns (SASS LDTM.x16)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// fp32 -> (tf32 hi, tf32 lo) with INTEGER arithmetic: round-to-nearest (ties away, the semantics of cvt.rna.tf32.f32) is "add half an ulp of
// the 13 dropped bits, clear them".  cvt.rna.tf32.f32 issues at a fraction of the ALU rate on sm_100a and made the four converter warps the
// limiter of the whole kernel (2 conversions per element); IADD / LOP3 / FSUB run at full rate.  The tensor core ignores the low 13 bits of
// a tf32 operand, so "lo" only needs the rounding add, not the mask.
__device__ __forceinline__ float tf32_rn_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_rn_lo(float r) { return __uint_as_float(__float_as_uint(r) + 0x1000u); }
__device__ __forceinline__ float4 split_tf32(float4& v) {   // v <- hi (tf32, round to nearest), returns lo = tf32_rn(v - hi)
    float4 lo;
    const float hx = tf32_rn_hi(v.x), hy = tf32_rn_hi(v.y), hz = tf32_rn_hi(v.z), hw = tf32_rn_hi(v.w);
    lo.x = tf32_rn_lo(__fsub_rn(v.x, hx));
    lo.y = tf32_rn_lo(__fsub_rn(v.y, hy));
    lo.z = tf32_rn_lo(__fsub_rn(v.z, hz));
    lo.w = tf32_rn_lo(__fsub_rn(v.w, hw));
    v = make_float4(hx, hy, hz, hw);
    return lo;
}

// RAWHI (opt-in, dab_set_option "gemm_rawhi"): the tensor core reads an fp32 word as tf32 by IGNORING its low 13 mantissa bits (verified on
// B200: results identical to an explicit truncation), so the raw tile can serve as the "hi" operand as it is; the converters then only write
// "lo" = tf32_rn(v - trunc_tf32(v)) and the shared-memory traffic of the conversion drops from 3 to 2 tile passes per operand (the kernel is
// shared-memory-bandwidth bound: tensor-core operand reads + conversion = ~190 KiB per 32-deep k-block against 128 B/clk).  Measured at
// 8192^3: +2-5 % speed, but the always-positive remainder of a truncation makes the dropped a_lo*b_lo term a bias: max error 1.02e-6 instead of
// 0.93e-6 at k = 8192 on same-sign data.  The default keeps the round-to-nearest split (hi rewritten in place), which stays inside 1e-6.
__device__ __forceinline__ float4 lo_of_trunc(const float4& v) {   // tf32_rn(v - trunc_tf32(v)), the remainder of the hardware's truncation
    float4 lo;
    lo.x = tf32_rn_lo(__fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xffffe000u)));
    lo.y = tf32_rn_lo(__fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)));
    lo.z = tf32_rn_lo(__fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xffffe000u)));
    lo.w = tf32_rn_lo(__fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xffffe000u)));
    return lo;
}

template <bool TA, bool RAWHI>
__global__ void __launch_bounds__(TG_THREADS, 1) gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                    float* __restrict__ C, size_t ldc, uint32_t m, uint32_t n, uint32_t k, uint32_t kc_blocks) {
    extern __shared__ unsigned char tg_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tg_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TG_BAR_OFFSET);
    // barrier indices: full[s] = s, conv[s] = 3 + s, empty[s] = 6 + s, acc_full[a] = 9 + a, acc_empty[a] = 11 + a
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const uint32_t bar0 = smem_u32(bars);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m0 = blockIdx.x * TG_M, n0 = blockIdx.y * TG_N;
    const uint32_t nkb = (k + TG_K - 1) / TG_K;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TG_STAGES; ++s) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 8 * (3 + s), 128);
            mbar_init(bar0 + 8 * (6 + s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar0 + 8 * (9 + a), 1);
            mbar_init(bar0 + 8 * (11 + a), TG_DRAIN_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // one warp allocates the tensor memory (and frees it at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sbase = smem_u32(smem);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const uint32_t s = kb % TG_STAGES, ph = (kb / TG_STAGES) & 1u;
                mbar_wait(bar0 + 8 * (6 + s), ph ^ 1u);
                const uint32_t full = bar0 + 8 * s;
                mbar_expect_tx(full, 2 * TG_TILE_BYTES);
                const uint32_t a_dst = sbase + s * TG_STAGE_BYTES, b_dst = a_dst + 2 * TG_TILE_BYTES;
                const int32_t k0 = (int32_t)(kb * TG_K);
                if (TA) {
                    tma_load_2d(a_dst, &mapA, full, k0, (int32_t)m0);                       // box {32 k, 128 m}: K-major rows of 128 bytes
                } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a)                                             // box {32 m, 32 k}: four M-atoms of 32 rows x 128 bytes
                        tma_load_2d(a_dst + a * 4096, &mapA, full, (int32_t)(m0 + 32 * a), k0);
                }
                tma_load_2d(b_dst, &mapB, full, k0, (int32_t)n0);                           // box {32 k, 128 n}
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread, accumulators in tensor memory =====
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, A major (1 = MN-major, the untransposed column-major A), B K-major, N, M
            const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((TA ? 0u : 1u) << 15) | (0u << 16) | ((uint32_t)(TG_M >> 4) << 24);
            const uint32_t idesc = idesc_base | ((uint32_t)(TG_N >> 3) << 17);            // N = 128
            const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * TG_N) >> 3) << 17);     // N = 256: B = [b_hi | b_lo], adjacent tiles of the stage
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const uint32_t s = kb % TG_STAGES, ph = (kb / TG_STAGES) & 1u;
                const uint32_t chunk = kb / kc_blocks, acc = chunk & 1u, use = chunk >> 1;
                const bool chunk_start = (kb % kc_blocks) == 0;
                if (chunk_start) {
                    mbar_wait(bar0 + 8 * (11 + acc), (use & 1u) ^ 1u);                      // the drain warps have emptied this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(bar0 + 8 * (3 + s), ph);                                          // converted tiles are in place
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = sbase + s * TG_STAGE_BYTES, a_lo = a_hi + TG_TILE_BYTES, b_hi = a_hi + 2 * TG_TILE_BYTES;   // b_lo follows b_hi
                const uint32_t d = tmem_base + acc * (2 * TG_N), ds = d + TG_N;            // main (hi*hi) and correction accumulators of this set
#pragma unroll
                for (int j = 0; j < TG_K / 8; ++j) {
                    // A untransposed = "MN-major" (m contiguous).  For 32-bit operands the tensor core only takes MN-major tiles in the 128-byte
                    // swizzle with 32-BYTE atoms (UMMA layout type 1; TMA SWIZZLE_128B_ATOM_32B): rows are k (128 B = 32 m each), K-atoms of 4 rows
                    // are 512 B apart (SBO), the four 32-m M-atoms 4096 B (LBO); k-step j (8 k) starts 1024 B further.
                    // A transposed and B = K-major, plain 128-byte swizzle (layout type 2): rows are m / n (128 B = 32 k each), 8-row atoms
                    // 1024 B apart (SBO); k-step j starts 32 B further inside the swizzled row.
                    const uint32_t aoff = TA ? (uint32_t)j * 32u : (uint32_t)j * 1024u;
                    const uint32_t albo = TA ? 16u : 4096u, asbo = TA ? 1024u : 512u, alay = TA ? 2u : 1u;
                    const uint64_t dah = umma_desc(a_hi + aoff, albo, asbo, alay), dal = umma_desc(a_lo + aoff, albo, asbo, alay);
                    const uint64_t dbh = umma_desc(b_hi + j * 32u, 16u, 1024u);
                    const uint32_t first = (chunk_start && j == 0) ? 0u : 1u;
                    // a_hi x [b_hi | b_lo] as ONE N = 256 instruction: columns 0-127 of the accumulator set take the main product, columns
                    // 128-255 the correction a_hi*b_lo (b_lo sits right behind b_hi in the stage, same 1024-byte atom stride) -- a_hi is
                    // read from shared memory once instead of twice; then the other correction a_lo x b_hi onto columns 128-255
                    umma_tf32(d, dah, dbh, idesc2, first);
                    umma_tf32(ds, dal, dbh, idesc, 1u);
                }
                umma_commit(bar0 + 8 * (6 + s));                                            // frees the smem stage when these MMAs have read it
                if ((kb + 1) % kc_blocks == 0 || kb + 1 == nkb) umma_commit(bar0 + 8 * (9 + acc));   // partial tile complete
            }
        }
    } else if (warp < 6) {
        // ===== converters: fp32 -> (tf32 hi, tf32 lo), elementwise, so the swizzled layout is irrelevant =====
        const int c = threadIdx.x - 64;
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            const uint32_t s = kb % TG_STAGES, ph = (kb / TG_STAGES) & 1u;
            mbar_wait(bar0 + 8 * s, ph);
            float4* a_hi = reinterpret_cast<float4*>(smem + s * TG_STAGE_BYTES);
            float4* a_lo = a_hi + TG_TILE_BYTES / 16;
            float4* b_hi = a_hi + 2 * (TG_TILE_BYTES / 16);
            float4* b_lo = a_hi + 3 * (TG_TILE_BYTES / 16);
#pragma unroll
            for (int q = 0; q < TG_TILE_BYTES / 16 / 128; ++q) {
                const int i = c + q * 128;
                float4 va = a_hi[i], vb = b_hi[i];
                if (RAWHI) {
                    a_lo[i] = lo_of_trunc(va);
                    b_lo[i] = lo_of_trunc(vb);
                } else {
                    const float4 la = split_tf32(va), lb = split_tf32(vb);
                    a_hi[i] = va;
                    a_lo[i] = la;
                    b_hi[i] = vb;
                    b_lo[i] = lb;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                   // generic-proxy stores -> visible to the tensor core's async proxy
            mbar_arrive(bar0 + 8 * (3 + s));
        }
    } else {
        // ===== drain + epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; lane == output row; two warps share a lane quarter and take
        // 64 of the 128 accumulator columns each (keeps the running fp32 sums in registers) =====
        const uint32_t quarter = (uint32_t)(warp & 3);
        const uint32_t half = (uint32_t)(warp - 6) >> 2;
        const uint32_t row = m0 + quarter * 32 + lane;
        constexpr int NC = TG_N / 2;
        float acc[NC];
        const uint32_t nchunks = (nkb + kc_blocks - 1) / kc_blocks;
        for (uint32_t ch = 0; ch < nchunks; ++ch) {
            const uint32_t a = ch & 1u, use = ch >> 1;
            mbar_wait(bar0 + 8 * (9 + a), use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c4 = 0; c4 < NC / 16; ++c4) {
                uint32_t v[16], w[16];
                const uint32_t taddr = tmem_base + a * (2 * TG_N) + half * NC + c4 * 16 + ((quarter * 32u) << 16);
                tmem_ld16(taddr, v);                  // main partial (hi*hi)
                tmem_ld16(taddr + TG_N, w);           // correction partial (lo*hi + hi*lo)
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p = __fadd_rn(__uint_as_float(v[i]), __uint_as_float(w[i]));
                    acc[c4 * 16 + i] = ch == 0 ? p : __fadd_rn(acc[c4 * 16 + i], p);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar0 + 8 * (11 + a));
        }
        if (row < m) {
            float* crow = C + row;
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const uint32_t col = n0 + half * NC + j;
                if (col < n) crow[(size_t)col * ldc] = nkb ? acc[j] : 0.0f;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TG_TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}

// 2-D fp32 tensor map over a column-major matrix: dim 0 = the contiguous direction (extent d0), dim 1 = columns (extent d1, ld elements apart)
int32_t make_map(dab_ctx* ctx, CUtensorMap* map, const float* base, size_t d0, size_t d1, size_t ld, uint32_t box0, uint32_t box1,
                 CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return dab_fail(ctx, DAB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return dab_fail(ctx, DAB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (dims %zu x %zu, ld %zu)", (int)r, d0, d1, ld);
    return DAB_OK;
}

bool tma_ok(const void* p, size_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (ld % 4) == 0 && ld < ((size_t)1 << 38); }

int32_t launch_tf32x3(dab_ctx* ctx, int transA, size_t m, size_t n, size_t k, const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc) {
    CUtensorMap mapA, mapB;
    // untransposed column-major A is "MN-major" for the tensor core; for 32-bit (tf32) operands that needs the 128-byte swizzle with
    // 32-BYTE atoms (TMA SWIZZLE_128B_ATOM_32B, UMMA layout type 1) -- the 16-byte-atom swizzle is only defined for K-major tf32
    int32_t st = transA ? make_map(ctx, &mapA, A, k, m, lda, 32, 128) : make_map(ctx, &mapA, A, m, k, lda, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (st != DAB_OK) return st;
    st = make_map(ctx, &mapB, B, k, n, ldb, 32, 128);
    if (st != DAB_OK) return st;
    const size_t gx = (m + TG_M - 1) / TG_M, gy = (n + TG_N - 1) / TG_N;
    if (gx > 0x7fffffffull || gy > 65535ull) return dab_fail(ctx, DAB_ERR_UNSUPPORTED, "dab_gemm: tile grid %zu x %zu too large", gx, gy);
    long long kc = ctx->opt_gemm_kc > 0 ? ctx->opt_gemm_kc : 64;
    uint32_t kc_blocks = (uint32_t)((kc + TG_K - 1) / TG_K);
    if (kc_blocks < 1) kc_blocks = 1;
    const bool raw = ctx->opt_gemm_rawhi != 0;
    auto launch = [&](auto kern) -> int32_t {
        static std::mutex mu;
        static std::map<std::pair<const void*, int>, bool> done;   // the >48 KiB opt-in is per (function, device)
        {
            std::lock_guard<std::mutex> lk(mu);
            auto key = std::make_pair((const void*)kern, ctx->device);
            if (!done.count(key)) {
                DAB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES));
                done[key] = true;
            }
        }
        dim3 grid((unsigned)gx, (unsigned)gy);
        kern<<<grid, TG_THREADS, TG_SMEM_BYTES, ctx->stream>>>(mapA, mapB, C, ldc, (uint32_t)m, (uint32_t)n, (uint32_t)k, kc_blocks);
        return DAB_OK;
    };
    int32_t rc;
    if (transA) rc = raw ? launch(gemm_tf32x3_kernel<true, true>) : launch(gemm_tf32x3_kernel<true, false>);
    else rc = raw ? launch(gemm_tf32x3_kernel<false, true>) : launch(gemm_tf32x3_kernel<false, false>);
    if (rc != DAB_OK) return rc;
    DAB_LAUNCHED(ctx);
    return DAB_OK;
}

}  // namespace

extern "C" {

int32_t dab_gemm(dab_ctx* ctx, int32_t dtype, int32_t transA, size_t m, size_t n, size_t k, const void* A, size_t lda, const void* B, size_t ldb, void* C,
                 size_t ldc) {
    DAB_ENTER(ctx);
    if (m == 0 || n == 0) return DAB_OK;
    DAB_REQUIRE(ctx, C && (k == 0 || (A && B)), DAB_ERR_ARG, "dab_gemm: null pointer");
    DAB_REQUIRE(ctx, ldc >= m && (k == 0 || (lda >= (transA ? k : m) && ldb >= k)), DAB_ERR_ARG, "dab_gemm: leading dimension smaller than the rows");
    // B with ONE column (A * b for a DMatrix b of width 1, the narrowest block of a column-split B): the product is the matrix-vector
    // product K9 already serves at the HBM roofline (one read of A, fp64 / wrap-around carriers); a 128-wide tensor-core tile would spend
    // 127/128 of its MMAs on padding and stream A at the shared-memory rate.  Needs a dense A (lda == rows), which is what a chunk is.
    if (n == 1 && k > 0 && lda == (transA ? k : m) && (dtype == DAB_F32 || dtype == DAB_F64 || dtype == DAB_I32 || dtype == DAB_I64))
        return dab_gemv(ctx, dtype, transA ? 1 : 0, A, lda, transA ? m : k, B, C);
    switch (dtype) {
        case DAB_F32: {
            const bool big = m < ((size_t)1 << 31) && n < ((size_t)1 << 31) && k < ((size_t)1 << 31);
            if (ctx->opt_gemm_simt == 0 && k > 0 && big && tma_ok(A, lda) && tma_ok(B, ldb))
                return launch_tf32x3(ctx, transA, m, n, k, (const float*)A, lda, (const float*)B, ldb, (float*)C, ldc);
            return launch_simt<float>(ctx, transA, m, n, k, (const float*)A, lda, (const float*)B, ldb, (float*)C, ldc);
        }
        case DAB_F64: return launch_simt<double>(ctx, transA, m, n, k, (const double*)A, lda, (const double*)B, ldb, (double*)C, ldc);
        case DAB_I32: return launch_simt<int32_t>(ctx, transA, m, n, k, (const int32_t*)A, lda, (const int32_t*)B, ldb, (int32_t*)C, ldc);
        case DAB_I64: return launch_simt<long long>(ctx, transA, m, n, k, (const long long*)A, lda, (const long long*)B, ldb, (long long*)C, ldc);
        default: return dab_fail(ctx, DAB_ERR_UNSUPPORTED, "dab_gemm: dtype %d (served: F32 F64 I32 I64)", dtype);
    }
}

}  // extern "C"
