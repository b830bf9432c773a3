// dab_common.cuh -- shared internals of libdab200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/dab200.h"

#define DAB_MAX_REDUCE_BLOCKS 262144
#define DAB_SLOT_BYTES 64
#define DAB_MAX_RANKS 64

struct dab_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    // reduction scratch (device)
    void* block_partials;   // DAB_MAX_REDUCE_BLOCKS * 16 bytes
    unsigned int* counter;  // single ticket counter, self-resetting
    void* result_slot;      // DAB_SLOT_BYTES: result of dab_reduce_host / dab_mapreduce_all
    void* gather_slots;     // DAB_MAX_RANKS * 8 bytes
    void* host_slot;        // pinned, DAB_MAX_RANKS * 8 bytes
    void* dim_scratch;      // scratch for split reducedim partials
    size_t dim_scratch_bytes;
    uint64_t launches;
    // NCCL
    void* comm;
    int rank, nranks;
    // fused reduce + all-gather + ordered fold over peer memory (dab_mailbox_*; dab_reduce.cu)
    void* mailbox;          // this rank's mailbox (device, cudaMalloc'ed, IPC-exported)
    void** peer_mbox_dev;   // device array [nranks] of mailbox addresses as mapped in THIS process
    void* peer_mbox_host[DAB_MAX_RANKS];
    int mbox_ranks;         // 0 = not attached
    unsigned long long mbox_seq;
    unsigned long long barrier_seq;   // dab_peer_barrier: how many device-side barriers this rank has entered
    int fuse_op;            // >= 0: the next launch_reduce appends the cross-rank combine for this DAB_* op
    struct dab_alloc_cache* cache;  // size-bucketed reuse of small cudaMalloc blocks (dab_core.cu)
    void* sort_dev;         // radix-sort scratch: digit histograms + per-tile counts (dab_sort.cu)
    size_t sort_dev_bytes;
    void* stage[2];         // pinned staging buffers of the pipelined pageable H2D path (dab_h2d)
    cudaEvent_t stage_ev[2];
    void* sort_host;        // pinned: split-point staging of dab_sorted_split
    unsigned long long sort_epoch;  // one per digit pass ever launched: tags the look-back words so the scratch is never re-cleared
    long long opt_combine_timeout_ms;  // dab_set_option("combine_timeout_ms"): how long the fused combine waits for a peer (default 120 s)
    long long opt_gemm_kc;  // dab_set_option("gemm_kc"): k extent summed inside tensor memory before a partial tile is drained (default 64)
    int opt_gemm_rawhi;     // dab_set_option("gemm_rawhi"): 1 = raw fp32 tile as the tf32 "hi" operand (hardware truncation), 0 = RN split
    int opt_gemm_simt;      // dab_set_option("gemm_simt"): 1 = force the SIMT tile kernel for Float32 (A/B measurements)
    int opt_gemv_phase;     // dab_set_option("gemv_phase"): 1 (default) = phase-class kernel for A*x, 0 = the single-wave aligned / unit-wise pair
    int opt_gemv_t_cols;    // dab_set_option("gemv_t_cols"): columns one thread of the A'*x kernel carries (4 or 8)
    int opt_gemv_t_waves;   // dab_set_option("gemv_t_waves"): waves of CTAs the A'*x kernel is split into
    int opt_ew_tma;         // dab_set_option("ew_tma"): route aligned unary elementwise launches through the TMA-staged kernel
    char err[512];
};

extern thread_local char dab_tls_err[512];

int32_t dab_fail(dab_ctx* ctx, int32_t status, const char* fmt, ...);
int32_t dab_fail_cuda(dab_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);

#define DAB_CUDA(ctx, call)                                                             \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) return dab_fail_cuda((ctx), e__, #call, __FILE__, __LINE__); \
    } while (0)

#define DAB_REQUIRE(ctx, cond, status, ...)                          \
    do {                                                             \
        if (!(cond)) return dab_fail((ctx), (status), __VA_ARGS__);  \
    } while (0)

#define DAB_ENTER(ctx)                                                      \
    do {                                                                    \
        if ((ctx) == nullptr) return dab_fail(nullptr, DAB_ERR_ARG, "null ctx"); \
        DAB_CUDA((ctx), cudaSetDevice((ctx)->device));                      \
    } while (0)

// after a kernel launch
#define DAB_LAUNCHED(ctx)                          \
    do {                                           \
        (ctx)->launches++;                         \
        DAB_CUDA((ctx), cudaGetLastError());       \
    } while (0)

static inline size_t dab_dtype_size(int32_t dt) {
    switch (dt) {
        case DAB_F32: return 4;
        case DAB_F64: return 8;
        case DAB_I32: return 4;
        case DAB_I64: return 8;
        case DAB_U8: return 1;
        default: return 0;
    }
}

// Cross-rank combine fused into the reduce kernel's last CTA (see reduce_kernel): nranks == 0 disables it.
#define DAB_MBOX_SLOT 32                                   /* [0,8) result  [8,16) wide carrier  [16,24) sequence flag */
#define DAB_MBOX_BARRIER_OFFSET (2 * DAB_MAX_RANKS * DAB_MBOX_SLOT) /* after the two parity banks: one 8-byte arrival counter per rank (dab_peer_barrier) */
#define DAB_MBOX_BYTES (DAB_MBOX_BARRIER_OFFSET + DAB_MAX_RANKS * 8)
struct FusedComm {
    void* const* peers;        // device array of the nranks mailboxes
    void* host_out;            // pinned host slot: [0,8) folded result, [8,16) status (0 ok, 1 timed out)
    unsigned long long seq;
    unsigned long long timeout_ns;   // wall-clock bound of the mailbox poll (%globaltimer), dab_set_option("combine_timeout_ms")
    int rank, nranks, op;
};

// ---- device helpers --------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dab_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// 16-byte streaming load/store (evict-first: every element of the hot path is touched once).
__device__ __forceinline__ int4 ld_stream(const int4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(int4* p, int4 v) { __stcs(p, v); }

template <typename T>
struct alignas(16) Pack {
    static constexpr int N = 16 / sizeof(T);
    T v[N];
};

template <typename T>
__device__ __forceinline__ Pack<T> as_pack(int4 r) {
    Pack<T> p;
    memcpy(&p, &r, 16);
    return p;
}
template <typename T>
__device__ __forceinline__ int4 as_int4(const Pack<T>& p) {
    int4 r;
    memcpy(&r, &p, 16);
    return r;
}

// counter-based RNG shared bit-for-bit with oracle/oracle_core.c (hash_u32)
__host__ __device__ __forceinline__ uint32_t dab_hash_u32(uint64_t seed, uint64_t idx) {
    uint64_t z = idx + (seed + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

static inline int dab_grid_for(const dab_ctx* ctx, size_t work_items, int per_sm) {
    size_t cap = (size_t)ctx->sm_count * (size_t)per_sm;
    size_t g = work_items < cap ? work_items : cap;
    return (int)(g < 1 ? 1 : g);
}

// resident CTAs per SM of `kernel` at `threads` threads (cached per kernel): persistent grids are sized to exactly one
// wave (sm_count x resident CTAs) so that the grid-stride loops have no partial second wave.
int dab_resident_ctas(const void* kernel, int threads);

template <typename K>
static inline int dab_persistent_grid(const dab_ctx* ctx, K kernel, int threads, size_t work_items) {
    return dab_grid_for(ctx, work_items, dab_resident_ctas((const void*)kernel, threads));
}
