// dab_jit.cu -- general fused broadcast: ONE kernel per localpart for an arbitrary expression tree, compiled at run time.
//
// Replaces what Julia's JIT does for  copyto!(localpart(dest), lbc::Broadcasted)  (reference src/broadcast.jl:80) and
// copy(lbc) (:96): the whole tree  f.(args...)  is fused into a single pass over the chunk -- N-ary, nested, with size-1
// ("extruded", src/broadcast.jl:112-113) dims and mixed element types.  The host runtime (distributedarrays.jl_b200/
// _broadcast.py) traces the user's function into C source for ONE element (Julia promotion already applied); this file
// wraps it into two sm_100a kernels with NVRTC (-fmad=false: no FMA contraction, Julia semantics):
//   dab_bc_linear  : every array argument is dense and has the destination's shape -> 4 consecutive elements per thread,
//                    16/32-byte vector loads and stores, grid-stride (HBM roofline: sum of element sizes per element);
//   dab_bc_general : per-argument strides (0 = extruded dim), coalesced along dim 0.
// Kernels are cached per (device, expression, element types, argument kinds).  NVRTC and the driver entry points are
// resolved at run time (dlopen / cudaGetDriverEntryPoint), so libdab200.so loads on a machine without a GPU.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "dab_common.cuh"

namespace {

const char* kPrelude = R"PRELUDE(
typedef unsigned long long u64;
typedef long long i64;
#define DEV __device__ __forceinline__
// ---- Julia scalar semantics: one IEEE rounding per operation, never contracted ----
DEV float jl_add(float a, float b) { return __fadd_rn(a, b); }
DEV float jl_sub(float a, float b) { return __fsub_rn(a, b); }
DEV float jl_mul(float a, float b) { return __fmul_rn(a, b); }
DEV float jl_div(float a, float b) { return __fdiv_rn(a, b); }
DEV double jl_add(double a, double b) { return __dadd_rn(a, b); }
DEV double jl_sub(double a, double b) { return __dsub_rn(a, b); }
DEV double jl_mul(double a, double b) { return __dmul_rn(a, b); }
DEV double jl_div(double a, double b) { return __ddiv_rn(a, b); }
DEV int jl_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
DEV int jl_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
DEV int jl_mul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
DEV i64 jl_add(i64 a, i64 b) { return (i64)((u64)a + (u64)b); }
DEV i64 jl_sub(i64 a, i64 b) { return (i64)((u64)a - (u64)b); }
DEV i64 jl_mul(i64 a, i64 b) { return (i64)((u64)a * (u64)b); }
DEV float jl_max(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
DEV float jl_min(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
DEV double jl_max(double a, double b) {
    if (a != a || b != b) return __longlong_as_double(0x7ff8000000000000ll);
    if (a == b) return (__double_as_longlong(a) < 0) ? b : a;
    return a > b ? a : b;
}
DEV double jl_min(double a, double b) {
    if (a != a || b != b) return __longlong_as_double(0x7ff8000000000000ll);
    if (a == b) return (__double_as_longlong(a) < 0) ? a : b;
    return a < b ? a : b;
}
DEV int jl_max(int a, int b) { return a > b ? a : b; }
DEV int jl_min(int a, int b) { return a < b ? a : b; }
DEV i64 jl_max(i64 a, i64 b) { return a > b ? a : b; }
DEV i64 jl_min(i64 a, i64 b) { return a < b ? a : b; }
DEV bool jl_max(bool a, bool b) { return a || b; }
DEV bool jl_min(bool a, bool b) { return a && b; }
DEV float jl_rem(float a, float b) { return fmodf(a, b); }
DEV double jl_rem(double a, double b) { return fmod(a, b); }
DEV int jl_rem(int a, int b) { return (b == 0 || b == -1) ? 0 : a % b; }
DEV i64 jl_rem(i64 a, i64 b) { return (b == 0 || b == -1) ? 0 : a % b; }
DEV float jl_mod(float x, float y) { float r = fmodf(x, y); if (r == 0.f) return copysignf(r, y); return ((r > 0.f) != (y > 0.f)) ? __fadd_rn(r, y) : r; }
DEV double jl_mod(double x, double y) { double r = fmod(x, y); if (r == 0.0) return copysign(r, y); return ((r > 0.0) != (y > 0.0)) ? __dadd_rn(r, y) : r; }
DEV int jl_mod(int a, int b) { if (b == 0 || b == -1) return 0; int r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? r + b : r; }
DEV i64 jl_mod(i64 a, i64 b) { if (b == 0 || b == -1) return 0; i64 r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? r + b : r; }
DEV int jl_idiv(int a, int b) { if (b == 0) return 0; if (b == -1) return (int)(0u - (unsigned)a); return a / b; }
DEV i64 jl_idiv(i64 a, i64 b) { if (b == 0) return 0; if (b == -1) return (i64)(0ull - (u64)a); return a / b; }
DEV float jl_pow(float a, float b) { return powf(a, b); }
DEV double jl_pow(double a, double b) { return pow(a, b); }
DEV int jl_and(int a, int b) { return a & b; }
DEV int jl_or(int a, int b) { return a | b; }
DEV int jl_xor(int a, int b) { return a ^ b; }
DEV i64 jl_and(i64 a, i64 b) { return a & b; }
DEV i64 jl_or(i64 a, i64 b) { return a | b; }
DEV i64 jl_xor(i64 a, i64 b) { return a ^ b; }
DEV bool jl_and(bool a, bool b) { return a && b; }
DEV bool jl_or(bool a, bool b) { return a || b; }
DEV bool jl_xor(bool a, bool b) { return a != b; }
template <typename T> DEV bool jl_lt(T a, T b) { return a < b; }
template <typename T> DEV bool jl_le(T a, T b) { return a <= b; }
template <typename T> DEV bool jl_gt(T a, T b) { return a > b; }
template <typename T> DEV bool jl_ge(T a, T b) { return a >= b; }
template <typename T> DEV bool jl_eq(T a, T b) { return a == b; }
template <typename T> DEV bool jl_ne(T a, T b) { return a != b; }
DEV float jl_neg(float a) { return -a; }
DEV double jl_neg(double a) { return -a; }
DEV int jl_neg(int a) { return (int)(0u - (unsigned)a); }
DEV i64 jl_neg(i64 a) { return (i64)(0ull - (u64)a); }
DEV float jl_abs(float a) { return fabsf(a); }
DEV double jl_abs(double a) { return fabs(a); }
DEV int jl_abs(int a) { return a < 0 ? (int)(0u - (unsigned)a) : a; }
DEV i64 jl_abs(i64 a) { return a < 0 ? (i64)(0ull - (u64)a) : a; }
template <typename T> DEV T jl_abs2(T a) { return jl_mul(a, a); }
DEV float jl_sqrt(float a) { return __fsqrt_rn(a); }
DEV double jl_sqrt(double a) { return __dsqrt_rn(a); }
DEV float jl_inv(float a) { return __fdiv_rn(1.0f, a); }
DEV double jl_inv(double a) { return __ddiv_rn(1.0, a); }
DEV float jl_floor(float a) { return floorf(a); }
DEV double jl_floor(double a) { return floor(a); }
DEV float jl_ceil(float a) { return ceilf(a); }
DEV double jl_ceil(double a) { return ceil(a); }
template <typename T> DEV T jl_floor(T a) { return a; }
template <typename T> DEV T jl_ceil(T a) { return a; }
template <typename T> DEV T jl_sign(T a) { return a > (T)0 ? (T)1 : (a < (T)0 ? (T)(-1) : a); }
template <typename T> DEV bool jl_isnan(T a) { return a != a; }
template <typename T> DEV bool jl_isinf(T a) { return (a == a) && ((a - a) != (a - a)); }
template <typename T> DEV bool jl_isfinite(T a) { return (a - a) == (a - a); }
// transcendental functions: CUDA's libdevice (<= 1-2 ulp); NOT bit-identical to Julia's openlibm-derived kernels
#define JL_F1(name, ff, fd) DEV float jl_##name(float a) { return ff(a); } DEV double jl_##name(double a) { return fd(a); }
JL_F1(sin, sinf, sin) JL_F1(cos, cosf, cos) JL_F1(tan, tanf, tan) JL_F1(exp, expf, exp) JL_F1(exp2, exp2f, exp2)
JL_F1(log, logf, log) JL_F1(log2, log2f, log2) JL_F1(log10, log10f, log10) JL_F1(tanh, tanhf, tanh) JL_F1(sinh, sinhf, sinh)
JL_F1(cosh, coshf, cosh) JL_F1(atan, atanf, atan) JL_F1(asin, asinf, asin) JL_F1(acos, acosf, acos) JL_F1(expm1, expm1f, expm1)
JL_F1(log1p, log1pf, log1p) JL_F1(cbrt, cbrtf, cbrt)

struct BcParams {
    void* out;
    u64 shape[4];
    i64 ostr[4];
    const void* ptr[8];
    i64 str[8][4];
    u64 scalar[8];
};
template <typename T, int N> struct __align__(sizeof(T) * N) VecN { T v[N]; };
template <typename T> DEV T bits_as(u64 b) { T r; memcpy(&r, &b, sizeof(T)); return r; }
)PRELUDE";

// Int128 as the VALUE type of a fused map + reduce (f widens its argument: x -> Int128(x)^2 + 2 Int128(x) - 1, test/darray.jl:286-294).
// Appended to the prelude -- and NVRTC's --device-int128 switched on -- only for sources that mention the type, so every other generated
// kernel is byte-for-byte what it was.
const char* kPreludeI128 = R"PRELUDE(
typedef __int128 i128;
typedef unsigned __int128 u128;
DEV i128 jl_add(i128 a, i128 b) { return (i128)((u128)a + (u128)b); }
DEV i128 jl_sub(i128 a, i128 b) { return (i128)((u128)a - (u128)b); }
DEV i128 jl_mul(i128 a, i128 b) { return (i128)((u128)a * (u128)b); }
DEV i128 jl_neg(i128 a) { return (i128)((u128)0 - (u128)a); }
DEV i128 jl_abs(i128 a) { return a < 0 ? (i128)((u128)0 - (u128)a) : a; }
DEV i128 jl_max(i128 a, i128 b) { return a > b ? a : b; }
DEV i128 jl_min(i128 a, i128 b) { return a < b ? a : b; }
DEV i128 jl_and(i128 a, i128 b) { return a & b; }
DEV i128 jl_or(i128 a, i128 b) { return a | b; }
DEV i128 jl_xor(i128 a, i128 b) { return a ^ b; }
)PRELUDE";

bool mentions_i128(const char* expr) { return expr && strstr(expr, "i128") != nullptr; }

// Further unary functions of the reference's "scalar math" test (test/darray.jl:775-797) that CUDA's libdevice provides; same accuracy
// note as the transcendental block of the prelude (<= 1-2 ulp, not bit-identical to Julia's openlibm / SpecialFunctions kernels).  The
// tracer spells them jl_x_*; the block is appended only to sources that use one, so every other generated kernel stays byte-identical.
const char* kPreludeExt = R"PRELUDE(
#define JL_X1(name, ff, fd) DEV float jl_x_##name(float a) { return ff(a); } DEV double jl_x_##name(double a) { return fd(a); }
JL_X1(asinh, asinhf, asinh) JL_X1(acosh, acoshf, acosh) JL_X1(atanh, atanhf, atanh) JL_X1(exp10, exp10f, exp10)
JL_X1(sinpi, sinpif, sinpi) JL_X1(cospi, cospif, cospi) JL_X1(trunc, truncf, trunc) JL_X1(round, rintf, rint)
JL_X1(erf, erff, erf) JL_X1(erfc, erfcf, erfc) JL_X1(erfinv, erfinvf, erfinv) JL_X1(erfcinv, erfcinvf, erfcinv) JL_X1(erfcx, erfcxf, erfcx)
JL_X1(gamma, tgammaf, tgamma) JL_X1(loggamma, lgammaf, lgamma)
template <typename T> DEV T jl_x_trunc(T a) { return a; }
template <typename T> DEV T jl_x_round(T a) { return a; }
// x << n, x >> n (test/darray.jl:863-867) with Julia's semantics: the result has the type of x; n counts bits as an Int64; a negative n
// shifts the other way; shifting out every bit gives 0 (<<) or the sign fill (>>, arithmetic).
DEV i64 jl_x_shr(i64 a, i64 n);
DEV i64 jl_x_shl(i64 a, i64 n) {
    if (n < 0) return n <= -64 ? (a < 0 ? -1ll : 0ll) : (a >> (int)(-n));
    return n >= 64 ? 0ll : (i64)((u64)a << (int)n);
}
DEV i64 jl_x_shr(i64 a, i64 n) {
    if (n < 0) return n <= -64 ? 0ll : (i64)((u64)a << (int)(-n));
    return n >= 64 ? (a < 0 ? -1ll : 0ll) : (a >> (int)n);
}
DEV int jl_x_shl(int a, i64 n) {
    if (n < 0) return n <= -32 ? (a < 0 ? -1 : 0) : (a >> (int)(-n));
    return n >= 32 ? 0 : (int)((unsigned)a << (int)n);
}
DEV int jl_x_shr(int a, i64 n) {
    if (n < 0) return n <= -32 ? 0 : (int)((unsigned)a << (int)(-n));
    return n >= 32 ? (a < 0 ? -1 : 0) : (a >> (int)n);
}
)PRELUDE";

bool mentions_ext(const char* expr) { return expr && strstr(expr, "jl_x_") != nullptr; }

const char* ctype_of(int32_t dt) {
    switch (dt) {
        case DAB_F32: return "float";
        case DAB_F64: return "double";
        case DAB_I32: return "int";
        case DAB_I64: return "long long";
        case DAB_U8: return "bool";
        default: return nullptr;
    }
}

// value type of dab_mapreduce_expr: the array element types plus Int128
const char* vtype_of(int32_t dt) { return dt == DAB_I128 ? "i128" : ctype_of(dt); }

struct BcParamsHost {
    void* out;
    unsigned long long shape[4];
    long long ostr[4];
    const void* ptr[8];
    long long str[8][4];
    unsigned long long scalar[8];
};

struct Nvrtc {
    void* h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    bool ok = false;
    char why[256] = "";
};

Nvrtc& nvrtc() {
    static Nvrtc api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", nullptr};
    for (int i = 0; names[i] && !api.h; ++i) api.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!api.h) {
        snprintf(api.why, sizeof(api.why), "dlopen(libnvrtc.so.12) failed: %s", dlerror());
        return api;
    }
#define SYM(f, n)                                                                  \
    do {                                                                           \
        *(void**)(&api.f) = dlsym(api.h, n);                                       \
        if (!api.f) {                                                              \
            snprintf(api.why, sizeof(api.why), "libnvrtc lacks symbol %s", n);     \
            return api;                                                            \
        }                                                                          \
    } while (0)
    SYM(CreateProgram, "nvrtcCreateProgram");
    SYM(CompileProgram, "nvrtcCompileProgram");
    SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    SYM(GetCUBIN, "nvrtcGetCUBIN");
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    SYM(GetProgramLog, "nvrtcGetProgramLog");
    SYM(DestroyProgram, "nvrtcDestroyProgram");
    SYM(GetErrorString, "nvrtcGetErrorString");
#undef SYM
    api.ok = true;
    return api;
}

struct Driver {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    bool ok = false;
    char why[256] = "";
};

Driver& driver() {
    static Driver d;
    static bool tried = false;
    if (tried) return d;
    tried = true;
#define ENTRY(f, n)                                                                                       \
    do {                                                                                                  \
        cudaDriverEntryPointQueryResult qr;                                                               \
        if (cudaGetDriverEntryPoint(n, (void**)&d.f, cudaEnableDefault, &qr) != cudaSuccess || !d.f) {    \
            cudaGetLastError();                                                                           \
            snprintf(d.why, sizeof(d.why), "driver entry point %s unavailable", n);                       \
            return d;                                                                                     \
        }                                                                                                 \
    } while (0)
    ENTRY(ModuleLoadData, "cuModuleLoadData");
    ENTRY(ModuleGetFunction, "cuModuleGetFunction");
    ENTRY(LaunchKernel, "cuLaunchKernel");
    ENTRY(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
    ENTRY(GetErrorString, "cuGetErrorString");
#undef ENTRY
    d.ok = true;
    return d;
}

struct Compiled {
    CUfunction linear = nullptr, general = nullptr, rows = nullptr;
    int occ_linear = 1, occ_general = 1, occ_rows = 1;
};

std::mutex g_mu;
std::unordered_map<std::string, Compiled> g_cache;

std::string build_source(const char* expr, int32_t out_dt, int nargs, const int32_t* dts, const bool* is_arr) {
    std::string s = kPrelude;
    if (mentions_ext(expr)) s += kPreludeExt;
    s += "typedef ";
    s += ctype_of(out_dt);
    s += " OUT_T;\n";
    for (int k = 0; k < nargs; ++k) s += std::string("typedef ") + ctype_of(dts[k]) + " T" + std::to_string(k) + ";\n";
    s += "#define DAB_EXPR (";
    s += expr;
    s += ")\n";
    // ---- linear kernel: flat grid, one CTA per 2 x 256 vectors of 4 elements (same shape as ew1_kernel: measured 6.9 vs 6.6 TB/s
    // for the persistent grid-stride form); the extra last CTA takes the remainder vectors and the scalar tail
    s += "extern \"C\" __global__ void __launch_bounds__(256) dab_bc_linear(BcParams p) {\n"
         "  const u64 n = p.shape[0] * p.shape[1] * p.shape[2] * p.shape[3];\n"
         "  const u64 nv = n / 4;\n"
         "  const u64 ntiles = nv / 512;\n"
         "  OUT_T* o = (OUT_T*)p.out;\n";
    for (int k = 0; k < nargs; ++k) {
        std::string K = std::to_string(k);
        if (is_arr[k]) s += "  const T" + K + "* q" + K + " = (const T" + K + "*)p.ptr[" + K + "];\n";
        else s += "  const T" + K + " a" + K + " = bits_as<T" + K + ">(p.scalar[" + K + "]);\n";
    }
    s += "  if ((u64)blockIdx.x < ntiles) {\n"
         "    const u64 i0 = (u64)blockIdx.x * 512 + threadIdx.x;\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "    const VecN<T" + K + ", 4> v" + K + "_0 = *(const VecN<T" + K + ", 4>*)(q" + K + " + 4 * i0);\n";
            s += "    const VecN<T" + K + ", 4> v" + K + "_1 = *(const VecN<T" + K + ", 4>*)(q" + K + " + 4 * (i0 + 256));\n";
        }
    for (int u = 0; u < 2; ++u) {
        std::string U = std::to_string(u);
        s += "    { VecN<OUT_T, 4> r;\n"
             "#pragma unroll\n"
             "      for (int j = 0; j < 4; ++j) {\n";
        for (int k = 0; k < nargs; ++k)
            if (is_arr[k]) {
                std::string K = std::to_string(k);
                s += "        const T" + K + " a" + K + " = v" + K + "_" + U + ".v[j];\n";
            }
        s += "        r.v[j] = (OUT_T)DAB_EXPR;\n"
             "      }\n"
             "      *(VecN<OUT_T, 4>*)(o + 4 * (i0 + " + std::to_string(256 * u) + ")) = r; }\n";
    }
    s += "    return;\n"
         "  }\n"
         "  for (u64 i = ntiles * 512 + threadIdx.x; i < nv; i += blockDim.x) {\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "    const VecN<T" + K + ", 4> v" + K + " = *(const VecN<T" + K + ", 4>*)(q" + K + " + 4 * i);\n";
        }
    s += "    VecN<OUT_T, 4> r;\n"
         "#pragma unroll\n"
         "    for (int j = 0; j < 4; ++j) {\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "      const T" + K + " a" + K + " = v" + K + ".v[j];\n";
        }
    s += "      r.v[j] = (OUT_T)DAB_EXPR;\n"
         "    }\n"
         "    *(VecN<OUT_T, 4>*)(o + 4 * i) = r;\n"
         "  }\n"
         "  for (u64 i = nv * 4 + threadIdx.x; i < n; i += blockDim.x) {\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "    const T" + K + " a" + K + " = q" + K + "[i];\n";
        }
    s += "    o[i] = (OUT_T)DAB_EXPR;\n"
         "  }\n"
         "}\n";
    // ---- general (strided / extruded) kernel
    s += "extern \"C\" __global__ void __launch_bounds__(256) dab_bc_general(BcParams p) {\n"
         "  const u64 n0 = p.shape[0], n1 = p.shape[1], n2 = p.shape[2];\n"
         "  const u64 n = n0 * n1 * n2 * p.shape[3];\n"
         "  OUT_T* o = (OUT_T*)p.out;\n";
    for (int k = 0; k < nargs; ++k) {
        std::string K = std::to_string(k);
        if (is_arr[k]) s += "  const T" + K + "* q" + K + " = (const T" + K + "*)p.ptr[" + K + "];\n";
        else s += "  const T" + K + " a" + K + " = bits_as<T" + K + ">(p.scalar[" + K + "]);\n";
    }
    s += "  const u64 stride = (u64)gridDim.x * blockDim.x;\n"
         "  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {\n"
         "    const u64 i0 = i % n0, t0 = i / n0, i1 = t0 % n1, t1 = t0 / n1, i2 = t1 % n2, i3 = t1 / n2;\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "    const T" + K + " a" + K + " = q" + K + "[(i64)i0 * p.str[" + K + "][0] + (i64)i1 * p.str[" + K + "][1] + (i64)i2 * p.str[" + K +
                 "][2] + (i64)i3 * p.str[" + K + "][3]];\n";
        }
    s += "    o[(i64)i0 * p.ostr[0] + (i64)i1 * p.ostr[1] + (i64)i2 * p.ostr[2] + (i64)i3 * p.ostr[3]] = (OUT_T)DAB_EXPR;\n"
         "  }\n"
         "}\n";
    // ---- "rows" kernel: the destination is dense, every array argument is either dense along dim 0 (stride 1, 16-byte aligned
    // rows) or extruded along dim 0 (stride 0): each thread produces 4 consecutive elements of one row with vector accesses and
    // decomposes the index once per 4 elements.  Serves  a .- m  with a 1 x n  m,  a .* v  with a column vector v, ... (the
    // extrusion cases of reference src/broadcast.jl:103-120) at streaming speed.
    s += "extern \"C\" __global__ void __launch_bounds__(256) dab_bc_rows(BcParams p) {\n"
         "  const u64 n0v = p.shape[0] / 4, n1 = p.shape[1], n2 = p.shape[2];\n"
         "  const u64 total = n0v * n1 * n2 * p.shape[3];\n"
         "  OUT_T* o = (OUT_T*)p.out;\n";
    for (int k = 0; k < nargs; ++k) {
        std::string K = std::to_string(k);
        if (is_arr[k]) s += "  const T" + K + "* q" + K + " = (const T" + K + "*)p.ptr[" + K + "];\n  const bool d" + K + " = p.str[" + K + "][0] != 0;\n";
        else s += "  const T" + K + " a" + K + " = bits_as<T" + K + ">(p.scalar[" + K + "]);\n";
    }
    s += "  const u64 stride = (u64)gridDim.x * blockDim.x;\n"
         "  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {\n"
         "    const u64 i0 = (i % n0v) * 4, t0 = i / n0v, i1 = t0 % n1, t1 = t0 / n1, i2 = t1 % n2, i3 = t1 / n2;\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "    const i64 f" + K + " = (i64)i1 * p.str[" + K + "][1] + (i64)i2 * p.str[" + K + "][2] + (i64)i3 * p.str[" + K + "][3];\n";
            s += "    VecN<T" + K + ", 4> v" + K + ";\n";
            s += "    if (d" + K + ") v" + K + " = *(const VecN<T" + K + ", 4>*)(q" + K + " + f" + K + " + (i64)i0);\n";
            s += "    else { const T" + K + " b = q" + K + "[f" + K + "]; v" + K + ".v[0] = b; v" + K + ".v[1] = b; v" + K + ".v[2] = b; v" + K + ".v[3] = b; }\n";
        }
    s += "    VecN<OUT_T, 4> r;\n"
         "#pragma unroll\n"
         "    for (int j = 0; j < 4; ++j) {\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "      const T" + K + " a" + K + " = v" + K + ".v[j];\n";
        }
    s += "      r.v[j] = (OUT_T)DAB_EXPR;\n"
         "    }\n"
         "    *(VecN<OUT_T, 4>*)(o + (i64)i0 + (i64)i1 * p.ostr[1] + (i64)i2 * p.ostr[2] + (i64)i3 * p.ostr[3]) = r;\n"
         "  }\n"
         "}\n";
    return s;
}

int32_t compile_cubin(dab_ctx* ctx, const std::string& src, std::vector<char>* cubin_out, bool int128 = false) {
    Nvrtc& rt = nvrtc();
    if (!rt.ok) return dab_fail(ctx, DAB_ERR_NVRTC, "NVRTC unavailable: %s", rt.why);
    nvrtcProgram prog;
    nvrtcResult r = rt.CreateProgram(&prog, src.c_str(), "dab_broadcast.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return dab_fail(ctx, DAB_ERR_NVRTC, "nvrtcCreateProgram: %s", rt.GetErrorString(r));
    const char* opts[] = {"--gpu-architecture=sm_100a", "-fmad=false", "--std=c++17", "-lineinfo", "-default-device", "--device-int128"};
    r = rt.CompileProgram(prog, int128 ? 6 : 5, opts);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        rt.GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        if (ls) rt.GetProgramLog(prog, &log[0]);
        rt.DestroyProgram(&prog);
        if (log.size() > 400) log.resize(400);
        return dab_fail(ctx, DAB_ERR_NVRTC, "NVRTC compile failed (%s): %s", rt.GetErrorString(r), log.c_str());
    }
    size_t cs = 0;
    rt.GetCUBINSize(prog, &cs);
    cubin_out->resize(cs);
    rt.GetCUBIN(prog, cubin_out->data());
    rt.DestroyProgram(&prog);
    return DAB_OK;
}

int32_t compile(dab_ctx* ctx, const std::string& src, Compiled* out) {
    Driver& drv = driver();
    if (!drv.ok) return dab_fail(ctx, DAB_ERR_NVRTC, "CUDA driver API unavailable: %s", drv.why);
    std::vector<char> cubin;
    int32_t st = compile_cubin(ctx, src, &cubin);
    if (st != DAB_OK) return st;
    CUmodule mod;
    CUresult cr = drv.ModuleLoadData(&mod, cubin.data());
    if (cr != CUDA_SUCCESS) {
        const char* es = "?";
        drv.GetErrorString(cr, &es);
        return dab_fail(ctx, DAB_ERR_NVRTC, "cuModuleLoadData failed: %s", es);
    }
    if (drv.ModuleGetFunction(&out->linear, mod, "dab_bc_linear") != CUDA_SUCCESS ||
        drv.ModuleGetFunction(&out->general, mod, "dab_bc_general") != CUDA_SUCCESS ||
        drv.ModuleGetFunction(&out->rows, mod, "dab_bc_rows") != CUDA_SUCCESS)
        return dab_fail(ctx, DAB_ERR_NVRTC, "cuModuleGetFunction failed");
    if (drv.OccupancyMaxActiveBlocksPerMultiprocessor(&out->occ_rows, out->rows, 256, 0) != CUDA_SUCCESS || out->occ_rows < 1) out->occ_rows = 1;
    if (drv.OccupancyMaxActiveBlocksPerMultiprocessor(&out->occ_linear, out->linear, 256, 0) != CUDA_SUCCESS || out->occ_linear < 1)
        out->occ_linear = 1;
    if (drv.OccupancyMaxActiveBlocksPerMultiprocessor(&out->occ_general, out->general, 256, 0) != CUDA_SUCCESS || out->occ_general < 1)
        out->occ_general = 1;
    return DAB_OK;
}


// ------------------------------------------------------------------------------------------------------------------------------
// Fused map + reduce for an arbitrary traced expression:  mapreduce(f, op, args...)  on one localpart in ONE pass over HBM
// (reference src/mapreduce.jl:31 with a general closure f; also dot = mapreduce(*, +, x, y), isequal = all(x .== y), ...).
// Same structure as the hand-written reduce_kernel: flat grid, 2 x 16/32-byte vector loads per argument in flight, 8-value tree in
// the value type, wide (fp64 / int64) carrier, one partial per CTA; a second tiny launch folds the <= 16384 partials in a fixed
// order (deterministic) and writes the 16-byte result slot.
struct MrSpec {
    const char *tile_t, *acc_t, *out_t, *tile_comb, *acc_comb, *acc_id;
};

bool mr_spec(int32_t val_dt, int32_t op, MrSpec* sp) {
    const bool flt = val_dt == DAB_F32 || val_dt == DAB_F64, boolean = val_dt == DAB_U8;
    const char* vt = vtype_of(val_dt);
    if (val_dt == DAB_I128) {  // tile, carrier and result are all Int128; + and * wrap, so any grouping gives the same bits
        switch (op) {
            case DAB_SUM: *sp = {vt, vt, vt, "jl_add(a, b)", "jl_add(a, b)", "0"}; return true;
            case DAB_PROD: *sp = {vt, vt, vt, "jl_mul(a, b)", "jl_mul(a, b)", "1"}; return true;
            case DAB_MAX: *sp = {vt, vt, vt, "jl_max(a, b)", "jl_max(a, b)", "((u128)1 << 127)"}; return true;
            case DAB_MIN: *sp = {vt, vt, vt, "jl_min(a, b)", "jl_min(a, b)", "(~((u128)1 << 127))"}; return true;
            default: return false;
        }
    }
    switch (op) {
        case DAB_SUM:
        case DAB_COUNT:
            if (op == DAB_COUNT && !boolean) return false;
            if (flt) *sp = {vt, "double", vt, "jl_add(a, b)", "jl_add(a, b)", "0.0"};
            else if (boolean) *sp = {"int", "long long", "long long", "(a + b)", "jl_add(a, b)", "0ll"};
            else *sp = {"long long", "long long", "long long", "jl_add(a, b)", "jl_add(a, b)", "0ll"};
            return true;
        case DAB_PROD:
            if (boolean) return false;
            if (flt) *sp = {vt, "double", vt, "jl_mul(a, b)", "jl_mul(a, b)", "1.0"};
            else *sp = {"long long", "long long", "long long", "jl_mul(a, b)", "jl_mul(a, b)", "1ll"};
            return true;
        case DAB_MAX:
        case DAB_MIN: {
            if (boolean) return false;
            const bool mx = op == DAB_MAX;
            const char* id = val_dt == DAB_F32   ? (mx ? "(-__int_as_float(0x7f800000))" : "__int_as_float(0x7f800000)")
                             : val_dt == DAB_F64 ? (mx ? "(-__longlong_as_double(0x7ff0000000000000ll))" : "__longlong_as_double(0x7ff0000000000000ll)")
                             : val_dt == DAB_I32 ? (mx ? "((int)0x80000000)" : "0x7fffffff")
                                                 : (mx ? "((long long)0x8000000000000000ll)" : "0x7fffffffffffffffll");
            *sp = {vt, vt, vt, mx ? "jl_max(a, b)" : "jl_min(a, b)", mx ? "jl_max(a, b)" : "jl_min(a, b)", id};
            return true;
        }
        case DAB_ALL:
        case DAB_ANY:
            if (!boolean) return false;
            *sp = {"int", "long long", "long long", "(a + b)", "jl_add(a, b)", "0ll"};
            return true;
        default: return false;
    }
}

struct MrParamsHost {
    const void* ptr[8];
    unsigned long long scalar[8];
    unsigned long long n;
    void* partials;
    int tiles_per_cta;
};
struct MrFinalHost {
    const void* partials;
    void* out;
    long long n;
    unsigned int nparts;
    int mode;
};

std::string build_mr_source(const char* expr, int32_t val_dt, int32_t op, int nargs, const int32_t* dts, const bool* is_arr, const MrSpec& sp) {
    std::string s = kPrelude;
    if (val_dt == DAB_I128 || mentions_i128(expr)) s += kPreludeI128;
    if (mentions_ext(expr)) s += kPreludeExt;
    if (val_dt == DAB_I128) s += "#define DAB_ACC16 1\n";   // 16-byte carrier: shuffles and the result slot move four words
    s += std::string("typedef ") + vtype_of(val_dt) + " VAL_T;\n";
    for (int k = 0; k < nargs; ++k) s += std::string("typedef ") + ctype_of(dts[k]) + " T" + std::to_string(k) + ";\n";
    s += std::string("typedef ") + sp.tile_t + " TILE_T;\ntypedef " + sp.acc_t + " ACC_T;\ntypedef " + sp.out_t + " OUT_T;\n";
    s += "#define DAB_EXPR (";
    s += expr;
    s += ")\n";
    s += std::string("DEV TILE_T tile_comb(TILE_T a, TILE_T b) { return ") + sp.tile_comb + "; }\n";
    s += std::string("DEV ACC_T acc_comb(ACC_T a, ACC_T b) { return ") + sp.acc_comb + "; }\n";
    s += std::string("#define ACC_ID ((ACC_T)") + sp.acc_id + ")\n";
    s += R"MR(
struct MrParams { const void* ptr[8]; u64 scalar[8]; u64 n; void* partials; int tiles_per_cta; };
struct MrFinal { const void* partials; void* out; i64 n; unsigned int nparts; int mode; };
DEV ACC_T acc_shfl(ACC_T v, int d) {
#ifdef DAB_ACC16
    int w[4]; memcpy(w, &v, 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = __shfl_down_sync(0xffffffffu, w[k], d);
    ACC_T r; memcpy(&r, w, 16); return r;
#else
    if (sizeof(ACC_T) == 8) {
        i64 x; memcpy(&x, &v, 8);
        int lo = __shfl_down_sync(0xffffffffu, (int)(x & 0xffffffffll), d), hi = __shfl_down_sync(0xffffffffu, (int)(x >> 32), d);
        x = ((i64)hi << 32) | (unsigned int)lo;
        ACC_T r; memcpy(&r, &x, 8); return r;
    } else {
        int x = 0; memcpy(&x, &v, sizeof(ACC_T));
        x = __shfl_down_sync(0xffffffffu, x, d);
        ACC_T r; memcpy(&r, &x, sizeof(ACC_T)); return r;
    }
#endif
}
DEV ACC_T block_reduce(ACC_T acc, ACC_T* smem) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = acc_comb(acc, acc_shfl(acc, d));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = lane < 8 ? smem[lane] : ACC_ID;
#pragma unroll
        for (int d = 4; d > 0; d >>= 1) acc = acc_comb(acc, acc_shfl(acc, d));
    }
    return acc;
}
extern "C" __global__ void __launch_bounds__(256) dab_mr_final(MrFinal p) {
    __shared__ ACC_T smem[8];
    const ACC_T* parts = (const ACC_T*)p.partials;
    ACC_T acc = ACC_ID;
    unsigned int i = threadIdx.x;
    for (; i + 768 < p.nparts; i += 1024) {   // 4 independent loads in flight
        ACC_T a0 = parts[i], a1 = parts[i + 256], a2 = parts[i + 512], a3 = parts[i + 768];
        acc = acc_comb(acc, acc_comb(acc_comb(a0, a1), acc_comb(a2, a3)));
    }
    for (; i < p.nparts; i += 256) acc = acc_comb(acc, parts[i]);
    acc = block_reduce(acc, smem);
    if (threadIdx.x == 0) {
        OUT_T res;
        if (p.mode == 1) res = (OUT_T)(acc == (ACC_T)p.n);
        else if (p.mode == 2) res = (OUT_T)(acc != (ACC_T)0);
        else res = (OUT_T)acc;
#ifdef DAB_ACC16
        memcpy(p.out, &res, 16);
#else
        u64 w0 = 0, w1 = 0;
        memcpy(&w0, &res, sizeof(OUT_T));
        memcpy(&w1, &acc, sizeof(ACC_T));
        ((u64*)p.out)[0] = w0;
        ((u64*)p.out)[1] = w1;
#endif
    }
}
)MR";
    // ---- partial kernel
    s += "extern \"C\" __global__ void __launch_bounds__(256) dab_mr_partial(MrParams p) {\n"
         "  __shared__ ACC_T smem[8];\n"
         "  const u64 n = p.n, nv = n / 4, ntiles = nv / 512;\n";
    for (int k = 0; k < nargs; ++k) {
        std::string K = std::to_string(k);
        if (is_arr[k]) s += "  const T" + K + "* q" + K + " = (const T" + K + "*)p.ptr[" + K + "];\n";
        else s += "  const T" + K + " a" + K + " = bits_as<T" + K + ">(p.scalar[" + K + "]);\n";
    }
    s += "  ACC_T acc = ACC_ID;\n"
         "  u64 t_end = ((u64)blockIdx.x + 1) * (u64)p.tiles_per_cta;\n"
         "  if (t_end > ntiles) t_end = ntiles;\n"
         "#pragma unroll 1\n"
         "  for (u64 t = (u64)blockIdx.x * (u64)p.tiles_per_cta; t < t_end; ++t) {\n"
         "    const u64 i0 = t * 512 + threadIdx.x;\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "    const VecN<T" + K + ", 4> v" + K + "_0 = *(const VecN<T" + K + ", 4>*)(q" + K + " + 4 * i0);\n";
            s += "    const VecN<T" + K + ", 4> v" + K + "_1 = *(const VecN<T" + K + ", 4>*)(q" + K + " + 4 * (i0 + 256));\n";
        }
    s += "    TILE_T m[8];\n";
    for (int u = 0; u < 2; ++u) {
        std::string U = std::to_string(u);
        s += "#pragma unroll\n    for (int j = 0; j < 4; ++j) {\n";
        for (int k = 0; k < nargs; ++k)
            if (is_arr[k]) {
                std::string K = std::to_string(k);
                s += "      const T" + K + " a" + K + " = v" + K + "_" + U + ".v[j];\n";
            }
        s += "      m[" + std::to_string(4 * u) + " + j] = (TILE_T)((VAL_T)DAB_EXPR);\n    }\n";
    }
    s += "#pragma unroll\n"
         "    for (int w = 8; w > 1; w >>= 1)\n"
         "#pragma unroll\n"
         "      for (int k = 0; k < w / 2; ++k) m[k] = tile_comb(m[k], m[k + w / 2]);\n"
         "    acc = acc_comb(acc, (ACC_T)m[0]);\n"
         "  }\n"
         "  if (blockIdx.x == gridDim.x - 1) {\n"
         "    for (u64 i = ntiles * 2048 + threadIdx.x; i < n; i += blockDim.x) {\n";
    for (int k = 0; k < nargs; ++k)
        if (is_arr[k]) {
            std::string K = std::to_string(k);
            s += "      const T" + K + " a" + K + " = q" + K + "[i];\n";
        }
    s += "      acc = acc_comb(acc, (ACC_T)(TILE_T)((VAL_T)DAB_EXPR));\n"
         "    }\n"
         "  }\n"
         "  acc = block_reduce(acc, smem);\n"
         "  if (threadIdx.x == 0) ((ACC_T*)p.partials)[blockIdx.x] = acc;\n"
         "}\n";
    return s;
}

struct CompiledMr {
    CUfunction partial = nullptr, final_ = nullptr;
};
std::unordered_map<std::string, CompiledMr> g_mr_cache;

}  // namespace

extern "C" {

// Diagnostic (no GPU needed): run the same source generation + NVRTC compilation as dab_broadcast_expr and report the size
// of the sm_100a cubin.  Lets the host-side tests validate the tracer's code generation on a CPU-only machine.
int32_t dab_jit_compile_check(const char* expr, int32_t out_dtype, int32_t nargs, const int32_t* arg_dtypes,
                              const int32_t* arg_is_array, size_t* cubin_bytes) {
    if (!expr || !ctype_of(out_dtype) || nargs < 0 || nargs > 8 || (nargs && (!arg_dtypes || !arg_is_array)))
        return dab_fail(nullptr, DAB_ERR_ARG, "dab_jit_compile_check: bad argument");
    bool is_arr[8] = {false};
    for (int k = 0; k < nargs; ++k) {
        if (!ctype_of(arg_dtypes[k])) return dab_fail(nullptr, DAB_ERR_ARG, "dab_jit_compile_check: bad dtype of arg %d", k);
        is_arr[k] = arg_is_array[k] != 0;
    }
    std::vector<char> cubin;
    int32_t st = compile_cubin(nullptr, build_source(expr, out_dtype, nargs, arg_dtypes, is_arr), &cubin);
    if (st != DAB_OK) return st;
    if (cubin_bytes) *cubin_bytes = cubin.size();
    return DAB_OK;
}

int32_t dab_broadcast_expr(dab_ctx* ctx, const char* expr, int32_t out_dtype, void* out, const size_t shape[4],
                           const size_t out_strides[4], int32_t nargs, const int32_t* arg_dtypes, const void* const* arg_ptrs,
                           const size_t* arg_strides, const uint64_t* arg_scalars) {
    DAB_ENTER(ctx);
    DAB_REQUIRE(ctx, expr && out && shape && out_strides, DAB_ERR_ARG, "dab_broadcast_expr: null pointer");
    DAB_REQUIRE(ctx, nargs >= 0 && nargs <= 8, DAB_ERR_ARG, "dab_broadcast_expr: nargs %d (max 8)", nargs);
    DAB_REQUIRE(ctx, ctype_of(out_dtype), DAB_ERR_ARG, "dab_broadcast_expr: bad out dtype %d", out_dtype);
    DAB_REQUIRE(ctx, nargs == 0 || (arg_dtypes && arg_ptrs && arg_strides && arg_scalars), DAB_ERR_ARG, "dab_broadcast_expr: null arg table");
    size_t n = shape[0] * shape[1] * shape[2] * shape[3];
    if (n == 0) return DAB_OK;
    bool is_arr[8] = {false};
    std::string key = std::to_string(ctx->device) + "|" + std::to_string(out_dtype) + "|";
    for (int k = 0; k < nargs; ++k) {
        DAB_REQUIRE(ctx, ctype_of(arg_dtypes[k]), DAB_ERR_ARG, "dab_broadcast_expr: bad dtype of arg %d", k);
        is_arr[k] = arg_ptrs[k] != nullptr;
        key += std::to_string(arg_dtypes[k]) + (is_arr[k] ? "a" : "s");
    }
    key += "|";
    key += expr;
    Compiled comp;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_cache.find(key);
        if (it == g_cache.end()) {
            DAB_CUDA(ctx, cudaFree(0));  // make sure the primary context is current for the driver calls
            int32_t st = compile(ctx, build_source(expr, out_dtype, nargs, arg_dtypes, is_arr), &comp);
            if (st != DAB_OK) return st;
            g_cache[key] = comp;
        } else {
            comp = it->second;
        }
    }
    BcParamsHost p;
    memset(&p, 0, sizeof(p));
    p.out = out;
    size_t dense[4], acc = 1;
    for (int d = 0; d < 4; ++d) {
        p.shape[d] = shape[d];
        p.ostr[d] = (long long)out_strides[d];
        dense[d] = acc;
        acc *= shape[d];
    }
    bool linear = true;
    for (int d = 0; d < 4; ++d)
        if (shape[d] > 1 && out_strides[d] != dense[d]) linear = false;
    if ((uintptr_t)out % (4 * dab_dtype_size(out_dtype))) linear = false;
    for (int k = 0; k < nargs; ++k) {
        p.ptr[k] = arg_ptrs[k];
        p.scalar[k] = arg_scalars[k];
        for (int d = 0; d < 4; ++d) {
            p.str[k][d] = (long long)arg_strides[4 * k + d];
            if (is_arr[k] && shape[d] > 1 && arg_strides[4 * k + d] != dense[d]) linear = false;
        }
        if (is_arr[k] && ((uintptr_t)arg_ptrs[k] % (4 * dab_dtype_size(arg_dtypes[k])))) linear = false;
    }
    // rows kernel: dense destination, 4 | shape[0], every array argument dense-along-dim-0 with 4-element-aligned rows, or extruded
    bool rows = !linear && shape[0] % 4 == 0 && out_strides[0] == 1 && ((uintptr_t)out % (4 * dab_dtype_size(out_dtype))) == 0;
    for (int d = 1; d < 4 && rows; ++d)
        if (shape[d] > 1 && out_strides[d] % 4) rows = false;
    for (int k = 0; k < nargs && rows; ++k) {
        if (!is_arr[k]) continue;
        if (arg_strides[4 * k] == 0) continue;  // extruded along dim 0: scalar load per row
        if (arg_strides[4 * k] != 1 || ((uintptr_t)arg_ptrs[k] % (4 * dab_dtype_size(arg_dtypes[k])))) rows = false;
        for (int d = 1; d < 4 && rows; ++d)
            if (shape[d] > 1 && arg_strides[4 * k + d] % 4) rows = false;
    }
    Driver& drv = driver();
    void* args[] = {&p};
    CUfunction fn = linear ? comp.linear : (rows ? comp.rows : comp.general);
    size_t work = rows ? (n / 4 + 255) / 256 : (n + 255) / 256;
    size_t grid = linear ? (n / 4) / 512 + 1 : (size_t)dab_grid_for(ctx, work, rows ? comp.occ_rows : comp.occ_general);
    if (grid > 0x7fffffffull) return dab_fail(ctx, DAB_ERR_ARG, "array too large for one launch");
    CUresult cr = drv.LaunchKernel(fn, (unsigned)grid, 1, 1, 256, 1, 1, 0, (CUstream)ctx->stream, args, nullptr);
    if (cr != CUDA_SUCCESS) {
        const char* es = "?";
        drv.GetErrorString(cr, &es);
        return dab_fail(ctx, DAB_ERR_CUDA, "cuLaunchKernel failed: %s", es);
    }
    ctx->launches++;
    return DAB_OK;
}


int32_t dab_mapreduce_expr(dab_ctx* ctx, const char* expr, int32_t val_dtype, int32_t op, size_t n, int32_t nargs, const int32_t* arg_dtypes,
                           const void* const* arg_ptrs, const uint64_t* arg_scalars, void* out_dev) {
    DAB_ENTER(ctx);
    DAB_REQUIRE(ctx, expr && out_dev, DAB_ERR_ARG, "dab_mapreduce_expr: null pointer");
    DAB_REQUIRE(ctx, nargs >= 1 && nargs <= 8 && arg_dtypes && arg_ptrs && arg_scalars, DAB_ERR_ARG, "dab_mapreduce_expr: bad argument table");
    DAB_REQUIRE(ctx, vtype_of(val_dtype), DAB_ERR_ARG, "dab_mapreduce_expr: bad value dtype %d", val_dtype);
    DAB_REQUIRE(ctx, n > 0, DAB_ERR_EMPTY, "dab_mapreduce_expr: empty input (the host runtime handles n == 0)");
    MrSpec sp;
    if (!mr_spec(val_dtype, op, &sp))
        return dab_fail(ctx, DAB_ERR_UNSUPPORTED, "mapreduce op %d on value dtype %d is not served (no host fallback)", op, val_dtype);
    bool is_arr[8] = {false};
    bool vec_ok = true;
    std::string key = "mr|" + std::to_string(ctx->device) + "|" + std::to_string(val_dtype) + "|" + std::to_string(op) + "|";
    for (int k = 0; k < nargs; ++k) {
        DAB_REQUIRE(ctx, ctype_of(arg_dtypes[k]), DAB_ERR_ARG, "dab_mapreduce_expr: bad dtype of arg %d", k);
        is_arr[k] = arg_ptrs[k] != nullptr;
        key += std::to_string(arg_dtypes[k]) + (is_arr[k] ? "a" : "s");
        if (is_arr[k] && ((uintptr_t)arg_ptrs[k] % (4 * dab_dtype_size(arg_dtypes[k])))) vec_ok = false;
    }
    if (!vec_ok) return dab_fail(ctx, DAB_ERR_UNSUPPORTED, "dab_mapreduce_expr: arguments must be aligned to 4 elements");
    key += "|";
    key += expr;
    CompiledMr comp;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mr_cache.find(key);
        if (it == g_mr_cache.end()) {
            DAB_CUDA(ctx, cudaFree(0));
            Driver& drv = driver();
            if (!drv.ok) return dab_fail(ctx, DAB_ERR_NVRTC, "CUDA driver API unavailable: %s", drv.why);
            std::vector<char> cubin;
            int32_t st = compile_cubin(ctx, build_mr_source(expr, val_dtype, op, nargs, arg_dtypes, is_arr, sp), &cubin,
                                       val_dtype == DAB_I128 || mentions_i128(expr));
            if (st != DAB_OK) return st;
            CUmodule mod;
            if (drv.ModuleLoadData(&mod, cubin.data()) != CUDA_SUCCESS) return dab_fail(ctx, DAB_ERR_NVRTC, "cuModuleLoadData failed");
            if (drv.ModuleGetFunction(&comp.partial, mod, "dab_mr_partial") != CUDA_SUCCESS ||
                drv.ModuleGetFunction(&comp.final_, mod, "dab_mr_final") != CUDA_SUCCESS)
                return dab_fail(ctx, DAB_ERR_NVRTC, "cuModuleGetFunction failed");
            g_mr_cache[key] = comp;
        } else {
            comp = it->second;
        }
    }
    const size_t ntiles = (n / 4) / 512;
    size_t k = 2;
    const size_t max_parts = 16384;
    if ((ntiles + k - 1) / k > max_parts) k = (ntiles + max_parts - 1) / max_parts;
    size_t grid = (ntiles + k - 1) / k;
    if (grid < 1) grid = 1;
    MrParamsHost p;
    memset(&p, 0, sizeof(p));
    for (int a = 0; a < nargs; ++a) {
        p.ptr[a] = arg_ptrs[a];
        p.scalar[a] = arg_scalars[a];
    }
    p.n = n;
    p.partials = ctx->block_partials;
    p.tiles_per_cta = (int)k;
    MrFinalHost f;
    f.partials = ctx->block_partials;
    f.out = out_dev;
    f.n = (long long)n;
    f.nparts = (unsigned int)grid;
    f.mode = op == DAB_ALL ? 1 : (op == DAB_ANY ? 2 : 0);
    Driver& drv = driver();
    void* a1[] = {&p};
    void* a2[] = {&f};
    if (drv.LaunchKernel(comp.partial, (unsigned)grid, 1, 1, 256, 1, 1, 0, (CUstream)ctx->stream, a1, nullptr) != CUDA_SUCCESS ||
        drv.LaunchKernel(comp.final_, 1, 1, 1, 256, 1, 1, 0, (CUstream)ctx->stream, a2, nullptr) != CUDA_SUCCESS)
        return dab_fail(ctx, DAB_ERR_CUDA, "cuLaunchKernel failed (dab_mapreduce_expr)");
    ctx->launches += 2;
    return DAB_OK;
}

// Diagnostic twin of dab_jit_compile_check for the fused map+reduce kernels.
int32_t dab_jit_compile_check_reduce(const char* expr, int32_t val_dtype, int32_t op, int32_t nargs, const int32_t* arg_dtypes,
                                     const int32_t* arg_is_array, size_t* cubin_bytes) {
    if (!expr || !vtype_of(val_dtype) || nargs < 1 || nargs > 8 || !arg_dtypes || !arg_is_array)
        return dab_fail(nullptr, DAB_ERR_ARG, "dab_jit_compile_check_reduce: bad argument");
    MrSpec sp;
    if (!mr_spec(val_dtype, op, &sp)) return dab_fail(nullptr, DAB_ERR_UNSUPPORTED, "op %d on value dtype %d not served", op, val_dtype);
    bool is_arr[8] = {false};
    for (int k = 0; k < nargs; ++k) {
        if (!ctype_of(arg_dtypes[k])) return dab_fail(nullptr, DAB_ERR_ARG, "bad dtype of arg %d", k);
        is_arr[k] = arg_is_array[k] != 0;
    }
    std::vector<char> cubin;
    int32_t st = compile_cubin(nullptr, build_mr_source(expr, val_dtype, op, nargs, arg_dtypes, is_arr, sp), &cubin,
                               val_dtype == DAB_I128 || mentions_i128(expr));
    if (st != DAB_OK) return st;
    if (cubin_bytes) *cubin_bytes = cubin.size();
    return DAB_OK;
}

}  // extern "C"
