/*
 * oracle_core.c -- C restatement of the per-localpart loops of the DArray hot path.
 *
 * *** TEST INFRASTRUCTURE / CPU BASELINE ONLY ***  Not part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs (cpu_baseline, --impl reference) may load it.
 *
 * The reference (DistributedArrays.jl v0.6.9) is pure Julia and cannot run in this image; the
 * loops it executes per worker live in Julia Base.  This file restates them (same structure as
 * oracle/darray_oracle.py, which it must match bit-for-bit -- tests/test_oracle_core.py):
 *
 *   orc_affine_*      Base.Broadcast.copyto! fused loop for  y .= a .* x .+ b
 *                     (reference call site src/broadcast.jl:80; map! at src/mapreduce.jl:8).
 *                     Two roundings, never an FMA: build with -ffp-contract=off.
 *   orc_sum_*         Base.mapreduce_impl(identity, add_sum, A, 1, n, 1024): pairwise, block 1024,
 *                     @simd base block modelled with lanes*interleave accumulators
 *                     (call site src/mapreduce.jl:31).
 *   orc_max_/orc_min_ NaN-propagating, +0.0 > -0.0 (Julia max/min; call site src/mapreduce.jl:31).
 *   orc_sumdim_*      Base._mapreducedim! on the collapsed (inner, reduce, outer) shape
 *                     (call site src/mapreduce.jl:64).
 *   orc_fold_*        caller-side  reduce(op, results)  left fold (src/mapreduce.jl:34).
 *   orc_rand_u01_*    the synthetic-input generator shared with the CUDA side.
 *   orc_workers_*     P single-threaded "workers" (one pthread each, one chunk each), the
 *                     way the reference runs one Julia process per worker; returns seconds.
 *
 * Parity status: float reduction ORDER is a model of Julia Base ("parity unpinned" at the bit
 * level; see oracle/darray_oracle.py header).  Everything else is exact.
 */
#include <sched.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_BLOCK 1024
#define ORC_MAXW 64

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------------------------------------------------------- RNG */
static inline uint32_t hash_u32(uint64_t seed, uint64_t idx) {
    uint64_t z = idx + (seed + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

void orc_rand_u01_f32(float* x, uint64_t seed, uint64_t start, size_t n) {
    for (size_t i = 0; i < n; ++i) x[i] = (float)(hash_u32(seed, start + i) >> 8) * 0x1p-24f;
}
void orc_rand_u01_f64(double* x, uint64_t seed, uint64_t start, size_t n) {
    for (size_t i = 0; i < n; ++i) x[i] = (double)(hash_u32(seed, start + i) >> 8) * 0x1p-24;
}
/* exact integer sum of k_i = hash>>8 ; the exact array sum is ksum * 2^-24 */
uint64_t orc_rand_ksum(uint64_t seed, uint64_t start, size_t n) {
    uint64_t s = 0;
    for (size_t i = 0; i < n; ++i) s += (uint64_t)(hash_u32(seed, start + i) >> 8);
    return s;
}
/* exact integer sum of an array whose values are multiples of 2^-24 in [0,1) */
uint64_t orc_ksum_f32(const float* x, size_t n) {
    uint64_t s = 0;
    for (size_t i = 0; i < n; ++i) s += (uint64_t)(x[i] * 0x1p24f);
    return s;
}

/* ---------------------------------------------------------------- elementwise */
void orc_affine_f32(float* y, const float* x, float a, float b, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        float t = a * x[i];
        y[i] = t + b;
    }
}
void orc_affine_f64(double* y, const double* x, double a, double b, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        double t = a * x[i];
        y[i] = t + b;
    }
}

/* ---------------------------------------------------------------- pairwise sum */
#define DEF_SUM(T, SUF)                                                                          \
    static T base_block_##SUF(const T* a, size_t n, int lanes, int inter) {                      \
        int W = lanes * inter;                                                                   \
        T acc0 = a[0] + a[1];                                                                    \
        size_t m = n - 2;                                                                        \
        size_t nvec = (W > 1) ? m / (size_t)W : 0;                                               \
        if (nvec == 0) {                                                                         \
            T r = acc0;                                                                          \
            for (size_t i = 2; i < n; ++i) r = r + a[i];                                         \
            return r;                                                                            \
        }                                                                                        \
        T acc[ORC_MAXW];                                                                         \
        const T* p = a + 2;                                                                      \
        for (int k = 0; k < W; ++k) acc[k] = p[k];                                               \
        acc[0] = acc0 + acc[0];                                                                  \
        for (size_t j = 1; j < nvec; ++j) {                                                      \
            const T* q = p + j * (size_t)W;                                                      \
            for (int k = 0; k < W; ++k) acc[k] = acc[k] + q[k];                                  \
        }                                                                                        \
        T vec[ORC_MAXW];                                                                         \
        for (int l = 0; l < lanes; ++l) vec[l] = acc[l];                                         \
        for (int u = 1; u < inter; ++u)                                                          \
            for (int l = 0; l < lanes; ++l) vec[l] = acc[u * lanes + l] + vec[l];                \
        for (int w = lanes; w > 1; w >>= 1) {                                                    \
            int h = w >> 1;                                                                      \
            for (int l = 0; l < h; ++l) vec[l] = vec[l] + vec[l + h];                            \
        }                                                                                        \
        T r = vec[0];                                                                            \
        for (size_t i = 2 + nvec * (size_t)W; i < n; ++i) r = r + a[i];                          \
        return r;                                                                                \
    }                                                                                            \
    static T pairwise_##SUF(const T* a, size_t n, int lanes, int inter) {                        \
        if (n == 1) return a[0];                                                                 \
        if (n - 1 < ORC_BLOCK) return base_block_##SUF(a, n, lanes, inter);                      \
        size_t half = ((n - 1) >> 1) + 1;                                                        \
        T l = pairwise_##SUF(a, half, lanes, inter);                                             \
        T r = pairwise_##SUF(a + half, n - half, lanes, inter);                                  \
        return l + r;                                                                            \
    }                                                                                            \
    /* Base._mapreduce(identity, add_sum, IndexLinear(), A) */                                   \
    T orc_sum_##SUF(const T* a, size_t n, int lanes, int inter) {                                \
        if (n == 0) return (T)0;                                                                 \
        if (n == 1) return a[0];                                                                 \
        if (n < 16) {                                                                            \
            T r = a[0];                                                                          \
            for (size_t i = 1; i < n; ++i) r = r + a[i];                                         \
            return r;                                                                            \
        }                                                                                        \
        return pairwise_##SUF(a, n, lanes, inter);                                               \
    }                                                                                            \
    /* reduce(+, results): plain left fold (P < 16), src/mapreduce.jl:34 */                      \
    T orc_fold_sum_##SUF(const T* r, size_t p) {                                                 \
        T s = r[0];                                                                              \
        for (size_t i = 1; i < p; ++i) s = s + r[i];                                             \
        return s;                                                                                \
    }                                                                                            \
    /* Base._mapreducedim!(identity, +, R, A) on the collapsed shape (inner, reduce, outer),  */ \
    /* column-major; out has inner*outer elements and is ACCUMULATED ONTO (caller zero-fills). */\
    void orc_sumdim_##SUF(const T* x, size_t inner, size_t red, size_t outer, T* out, int lanes, \
                          int inter) {                                                           \
        if (inner == 1 && red > 16) {                                                            \
            for (size_t o = 0; o < outer; ++o)                                                   \
                out[o] = out[o] + pairwise_##SUF(x + o * red, red, lanes, inter);                \
            return;                                                                              \
        }                                                                                        \
        if (inner == 1) {                                                                        \
            for (size_t o = 0; o < outer; ++o) {                                                 \
                T r = out[o];                                                                    \
                for (size_t k = 0; k < red; ++k) r = r + x[o * red + k];                         \
                out[o] = r;                                                                      \
            }                                                                                    \
            return;                                                                              \
        }                                                                                        \
        for (size_t o = 0; o < outer; ++o)                                                       \
            for (size_t k = 0; k < red; ++k) {                                                   \
                const T* col = x + (o * red + k) * inner;                                        \
                T* dst = out + o * inner;                                                        \
                for (size_t i = 0; i < inner; ++i) dst[i] = dst[i] + col[i];                     \
            }                                                                                    \
    }

DEF_SUM(float, f32)
DEF_SUM(double, f64)

/* ---------------------------------------------------------------- max / min */
static inline float jl_max_f32(float x, float y) {
    if (isnan(x) || isnan(y)) return NAN;
    if (y > x) return y;
    if (x == y && signbit(x) && !signbit(y)) return y;
    return x;
}
static inline float jl_min_f32(float x, float y) {
    if (isnan(x) || isnan(y)) return NAN;
    if (y < x) return y;
    if (x == y && !signbit(x) && signbit(y)) return y;
    return x;
}
/* returns 0 and writes *out, or -1 for an empty collection (Julia throws) */
int orc_max_f32(const float* a, size_t n, float* out) {
    if (n == 0) return -1;
    float r = a[0];
    for (size_t i = 1; i < n; ++i) r = jl_max_f32(r, a[i]);
    *out = r;
    return 0;
}
int orc_min_f32(const float* a, size_t n, float* out) {
    if (n == 0) return -1;
    float r = a[0];
    for (size_t i = 1; i < n; ++i) r = jl_min_f32(r, a[i]);
    *out = r;
    return 0;
}
/* fast finite-only max used by the CPU baseline timing loop (8 running values like Base's
 * chunked mapreduce_impl for max/min; NaN handled by a flag) */
static float fast_max_f32(const float* a, size_t n) {
    float m[8];
    int nan = 0;
    size_t i = 0;
    if (n < 8) {
        float r;
        orc_max_f32(a, n, &r);
        return r;
    }
    for (int k = 0; k < 8; ++k) m[k] = a[k];
    for (i = 8; i + 8 <= n; i += 8)
        for (int k = 0; k < 8; ++k) {
            float v = a[i + k];
            nan |= (v != v);
            m[k] = v > m[k] ? v : m[k];
        }
    float r = m[0];
    for (int k = 0; k < 8; ++k) {
        nan |= (m[k] != m[k]);
        r = jl_max_f32(r, m[k]);
    }
    for (; i < n; ++i) r = jl_max_f32(r, a[i]);
    if (nan) return NAN;
    if (r == 0.0f) { /* signed-zero fix-up scan, as Base does */
        for (size_t j = 0; j < n; ++j)
            if (a[j] == 0.0f && !signbit(a[j])) return 0.0f;
    }
    return r;
}

/* ---------------------------------------------------------------- P workers, one thread each */
/* Each worker w owns chunk w of n_per elements (allocated + first-touched by its own thread), the
 * way the reference runs one single-threaded Julia process per worker.
 * op: 0 = affine in place (map!(x->a*x+b, d, d)), 1 = sum, 2 = maximum, 3 = affine then sum, 4 = sum(dims=1) of the chunk as a
 * 4096-row matrix, 5 = memcpy of the chunk.
 * Runs `iters` timed passes after `warm` warm-ups; returns best-of seconds per pass (wall time of
 * the slowest worker + the caller-side fold), writes the folded result of the last pass. */
typedef struct {
    int op, w, nworkers, passes;
    size_t n_per;
    uint64_t seed;
    float a, b;
    float* partial;
    pthread_barrier_t* bar;
} worker_arg;

static void* worker_main(void* p) {
    worker_arg* g = (worker_arg*)p;
    float* x = (float*)aligned_alloc(64, ((g->n_per * sizeof(float) + 63) / 64) * 64);
    orc_rand_u01_f32(x, g->seed, (uint64_t)g->w * g->n_per, g->n_per);
    /* op 4: sum(A, dims=1) of the chunk seen as a (4096 x n_per/4096) column-major matrix (per-column pairwise, A.3);
     * op 5: memcpy of the chunk (generous upper bound for the reference's serialise + TCP slab fetch) */
    const size_t rows = 4096, cols = g->n_per / rows;
    float* aux = NULL;
    if (g->op == 4) aux = (float*)calloc(cols ? cols : 1, sizeof(float));
    if (g->op == 5) aux = (float*)aligned_alloc(64, ((g->n_per * sizeof(float) + 63) / 64) * 64);
    for (int it = 0; it < g->passes; ++it) {
        pthread_barrier_wait(g->bar); /* pass start */
        if (g->op == 0 || g->op == 3) orc_affine_f32(x, x, g->a, g->b, g->n_per);
        if (g->op == 1 || g->op == 3) g->partial[g->w] = orc_sum_f32(x, g->n_per, 8, 4);
        if (g->op == 2) g->partial[g->w] = fast_max_f32(x, g->n_per);
        if (g->op == 4) {
            memset(aux, 0, cols * sizeof(float));
            orc_sumdim_f32(x, 1, rows, cols, aux, 8, 4);
            g->partial[g->w] = aux[cols / 2];
        }
        if (g->op == 5) {
            memcpy(aux, x, g->n_per * sizeof(float));
            g->partial[g->w] = aux[g->n_per / 2];
        }
        pthread_barrier_wait(g->bar); /* pass end */
    }
    free(x);
    free(aux);
    return NULL;
}

double orc_workers_run(int op, int nworkers, size_t n_per, uint64_t seed, float a, float b, int warm,
                       int iters, float* result, double* mean_s) {
    pthread_t* th = (pthread_t*)calloc((size_t)nworkers, sizeof(pthread_t));
    worker_arg* args = (worker_arg*)calloc((size_t)nworkers, sizeof(worker_arg));
    float* partial = (float*)calloc((size_t)nworkers, sizeof(float));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)nworkers + 1);
    double best = 1e30, tot = 0.0;
    for (int w = 0; w < nworkers; ++w) {
        worker_arg g = {op, w, nworkers, warm + iters, n_per, seed, a, b, partial, &bar};
        args[w] = g;
        pthread_create(&th[w], NULL, worker_main, &args[w]);
    }
    for (int it = 0; it < warm + iters; ++it) {
        pthread_barrier_wait(&bar);
        double t0 = now_s();
        pthread_barrier_wait(&bar);
        float r = partial[0];
        if (op == 2)
            for (int w = 1; w < nworkers; ++w) r = jl_max_f32(r, partial[w]);
        else
            r = orc_fold_sum_f32(partial, (size_t)nworkers);
        double dt = now_s() - t0;
        if (it >= warm) {
            if (dt < best) best = dt;
            tot += dt;
        }
        if (result) *result = r;
    }
    for (int w = 0; w < nworkers; ++w) pthread_join(th[w], NULL);
    pthread_barrier_destroy(&bar);
    free(th);
    free(args);
    free(partial);
    if (mean_s) *mean_s = tot / (double)(iters > 0 ? iters : 1);
    return best;
}

int orc_num_procs(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }

/* ---------------------------------------------------------------- exact statistics of the synthetic bench inputs (multi-threaded)
 * The checker behind bench.py's in-run `parity` block: for x_i = (hash>>8) * 2^-24 (the generator above) and
 * y_i = fl(fl(a*x_i) + b) (the broadcast y .= a .* x .+ b, two roundings) it returns, over [start, start+n):
 *   out[0] = sum of k_i            (the exact sum of x is out[0] * 2^-24)
 *   out[1] = sum of y_i * 2^yscale (exact integer when every y_i is a multiple of 2^-yscale; out[3] counts the ones that were not)
 *   out[2] = bits of max(y) (Julia max), out[4] = bits of max(x)
 * Pure integer / per-element IEEE arithmetic, so it is an exact ground truth, independent of any summation order. */
typedef struct {
    uint64_t seed, start;
    size_t n;
    float a, b;
    int yscale;
    uint64_t ksum, ysum, inexact;
    float ymax, xmax;
} stats_arg;

static void* stats_main(void* p) {
    stats_arg* g = (stats_arg*)p;
    uint64_t ks = 0, ys = 0, bad = 0;
    float ym = -INFINITY, xm = -INFINITY;
    for (size_t i = 0; i < g->n; ++i) {
        const uint32_t k = hash_u32(g->seed, g->start + i) >> 8;
        const float x = (float)k * 0x1p-24f;
        const float t = g->a * x;
        const float y = t + g->b;
        ks += k;
        const double sc = ldexp((double)y, g->yscale);
        const uint64_t q = (uint64_t)sc;
        bad += ((double)q != sc);
        ys += q;
        ym = jl_max_f32(ym, y);
        xm = jl_max_f32(xm, x);
    }
    g->ksum = ks;
    g->ysum = ys;
    g->inexact = bad;
    g->ymax = ym;
    g->xmax = xm;
    return NULL;
}

void orc_rand_stats_mt(uint64_t seed, uint64_t start, size_t n, float a, float b, int yscale, int nthreads, uint64_t* out) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    stats_arg* args = (stats_arg*)calloc((size_t)nthreads, sizeof(stats_arg));
    const size_t per = n / (size_t)nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const size_t lo = (size_t)t * per, hi = (t == nthreads - 1) ? n : lo + per;
        stats_arg g = {seed, start + lo, hi - lo, a, b, yscale, 0, 0, 0, 0.0f, 0.0f};
        args[t] = g;
        pthread_create(&th[t], NULL, stats_main, &args[t]);
    }
    uint64_t ks = 0, ys = 0, bad = 0;
    float ym = -INFINITY, xm = -INFINITY;
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        ks += args[t].ksum;
        ys += args[t].ysum;
        bad += args[t].inexact;
        ym = jl_max_f32(ym, args[t].ymax);
        xm = jl_max_f32(xm, args[t].xmax);
    }
    uint32_t yb, xb;
    memcpy(&yb, &ym, 4);
    memcpy(&xb, &xm, 4);
    out[0] = ks;
    out[1] = ys;
    out[2] = yb;
    out[3] = bad;
    out[4] = xb;
    free(th);
    free(args);
}

/* usable CPUs of THIS process (cgroup/affinity aware), for sizing the checker's thread pool */
int orc_num_usable_procs(void) {
#ifdef __linux__
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        int c = CPU_COUNT(&set);
        if (c > 0) return c;
    }
#endif
    return orc_num_procs();
}
