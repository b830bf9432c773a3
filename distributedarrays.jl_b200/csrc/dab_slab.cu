// dab_slab.cu -- K8: the "halo" read.  Copies an up-to-4-D box out of a (possibly PEER) chunk into a local dense array.
//
// Replaces, per intersecting chunk, the owner-side  localpart(d)[idxs...]  + Julia serialisation + TCP + a[idxs...] = ...
// of setindex!(a::Array, s::SubDArray, I...) (reference src/darray.jl:798-820), chunk(d, pid) (:458) and the non-local branch
// of makelocal (:361-366).  The reference's read is pull-style and one-sided from the reader's point of view; so is this:
// the READER launches the kernel and loads straight from the owner's HBM (peer mapping / CUDA IPC) over NVLink 5 with the
// widest vector the box geometry allows (16-byte LDG when base, pitches and row length are 16-byte multiples), 4 loads in
// flight per thread, and stores into its own HBM.  Roofline: NVLink (peer) or HBM (local): elem_bytes moved once per element.
#include "dab_common.cuh"

namespace {

struct BoxGeom {
    // all in units of `vec` bytes along dim 0, elements of pitch along dims 1..3 (bytes)
    unsigned long long upr;        // units per row
    unsigned long long e1, e2, e3; // rows along dims 1..3
    long long sp1, sp2, sp3;       // src pitches in bytes
    long long dp1, dp2, dp3;       // dst pitches in bytes
};

template <typename U, typename I>
__global__ void __launch_bounds__(256) copy_box_kernel(char* __restrict__ dst, const char* __restrict__ src, BoxGeom g, I total) {
    constexpr int UNROLL = 4;
    const I stride = (I)gridDim.x * blockDim.x;
    I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
    const I upr = (I)g.upr, e1 = (I)g.e1, e2 = (I)g.e2;
    for (; idx + (UNROLL - 1) * stride < total; idx += UNROLL * stride) {
        U v[UNROLL];
        size_t doff[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            I id = idx + u * stride;
            I row = id / upr, col = id - row * upr;
            I j = row % e1, t = row / e1;
            I k = t % e2, l = t / e2;
            size_t soff = (size_t)col * sizeof(U) + (size_t)j * g.sp1 + (size_t)k * g.sp2 + (size_t)l * g.sp3;
            doff[u] = (size_t)col * sizeof(U) + (size_t)j * g.dp1 + (size_t)k * g.dp2 + (size_t)l * g.dp3;
            v[u] = *reinterpret_cast<const U*>(src + soff);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) *reinterpret_cast<U*>(dst + doff[u]) = v[u];
    }
    for (; idx < total; idx += stride) {
        I row = idx / upr, col = idx - row * upr;
        I j = row % e1, t = row / e1;
        I k = t % e2, l = t / e2;
        size_t soff = (size_t)col * sizeof(U) + (size_t)j * g.sp1 + (size_t)k * g.sp2 + (size_t)l * g.sp3;
        size_t doff = (size_t)col * sizeof(U) + (size_t)j * g.dp1 + (size_t)k * g.dp2 + (size_t)l * g.dp3;
        *reinterpret_cast<U*>(dst + doff) = *reinterpret_cast<const U*>(src + soff);
    }
}

template <typename U>
int32_t launch_copy(dab_ctx* ctx, char* dst, const char* src, const BoxGeom& g) {
    unsigned long long total = g.upr * g.e1 * g.e2 * g.e3;
    if (total == 0) return DAB_OK;
    int grid = total < (1ull << 31) ? dab_persistent_grid(ctx, copy_box_kernel<U, unsigned int>, 256, (size_t)((total + 1023) / 1024))
                                    : dab_persistent_grid(ctx, copy_box_kernel<U, unsigned long long>, 256, (size_t)((total + 1023) / 1024));
    if (total < (1ull << 31)) {
        copy_box_kernel<U, unsigned int><<<grid, 256, 0, ctx->stream>>>(dst, src, g, (unsigned int)total);
    } else {
        copy_box_kernel<U, unsigned long long><<<grid, 256, 0, ctx->stream>>>(dst, src, g, total);
    }
    DAB_LAUNCHED(ctx);
    return DAB_OK;
}

}  // namespace

extern "C" {

int32_t dab_copy_box(dab_ctx* ctx, int32_t elem_bytes, void* dst, const size_t dst_shape[4], const size_t dst_off[4], const void* src,
                     const size_t src_shape[4], const size_t src_off[4], const size_t extent[4]) {
    DAB_ENTER(ctx);
    DAB_REQUIRE(ctx, dst && src && dst_shape && dst_off && src_shape && src_off && extent, DAB_ERR_ARG, "dab_copy_box: null pointer");
    DAB_REQUIRE(ctx, elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8 || elem_bytes == 16, DAB_ERR_ARG,
                "dab_copy_box: elem_bytes %d", elem_bytes);
    for (int d = 0; d < 4; ++d) {
        if (extent[d] == 0) return DAB_OK;
        DAB_REQUIRE(ctx, src_off[d] + extent[d] <= src_shape[d] && dst_off[d] + extent[d] <= dst_shape[d], DAB_ERR_DIM_MISMATCH,
                    "dab_copy_box: box exceeds array in dim %d (BoundsError)", d);
    }
    // byte geometry
    size_t es = (size_t)elem_bytes;
    size_t sp[4], dp[4];  // pitch (bytes) of one step along dim d
    sp[0] = dp[0] = es;
    for (int d = 1; d < 4; ++d) {
        sp[d] = sp[d - 1] * src_shape[d - 1];
        dp[d] = dp[d - 1] * dst_shape[d - 1];
    }
    const char* s = (const char*)src;
    char* t = (char*)dst;
    for (int d = 0; d < 4; ++d) {
        s += src_off[d] * sp[d];
        t += dst_off[d] * dp[d];
    }
    // collapse: a dim that spans both arrays entirely merges into the run below it
    size_t ext[4] = {extent[0], extent[1], extent[2], extent[3]};
    size_t row_bytes = ext[0] * es;
    size_t e[3] = {ext[1], ext[2], ext[3]};
    long long spp[3] = {(long long)sp[1], (long long)sp[2], (long long)sp[3]};
    long long dpp[3] = {(long long)dp[1], (long long)dp[2], (long long)dp[3]};
    int nd = 3;
    while (nd > 0 && (size_t)spp[0] == row_bytes && (size_t)dpp[0] == row_bytes) {
        row_bytes *= e[0];
        for (int d = 0; d + 1 < nd; ++d) {
            e[d] = e[d + 1];
            spp[d] = spp[d + 1];
            dpp[d] = dpp[d + 1];
        }
        e[nd - 1] = 1;
        spp[nd - 1] = 0;
        dpp[nd - 1] = 0;
        --nd;
    }
    // widest unit dividing every address component
    size_t align = (size_t)((uintptr_t)s | (uintptr_t)t | row_bytes);
    for (int d = 0; d < 3; ++d)
        if (e[d] > 1) align |= (size_t)spp[d] | (size_t)dpp[d];
    size_t vec = 16;
    while (vec > 1 && (align & (vec - 1))) vec >>= 1;
    if (vec > 16) vec = 16;
    BoxGeom g;
    g.upr = row_bytes / vec;
    g.e1 = e[0];
    g.e2 = e[1];
    g.e3 = e[2];
    g.sp1 = spp[0]; g.sp2 = spp[1]; g.sp3 = spp[2];
    g.dp1 = dpp[0]; g.dp2 = dpp[1]; g.dp3 = dpp[2];
    switch (vec) {
        case 16: return launch_copy<int4>(ctx, t, s, g);
        case 8: return launch_copy<long long>(ctx, t, s, g);
        case 4: return launch_copy<int>(ctx, t, s, g);
        case 2: return launch_copy<short>(ctx, t, s, g);
        default: return launch_copy<char>(ctx, t, s, g);
    }
}

}  // extern "C"
