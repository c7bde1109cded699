// vpt_math.cuh -- small float3 toolkit + counter-based Philox used by every kernel of the path.
//
// The render path is a chaotic estimator: one flipped accept/reject comparison changes a pixel's
// whole remaining path.  Per-seed parity with the reference kernel therefore needs the same IEEE
// operations in the same order wherever a value feeds a decision.  The expressions below are written
// in the operand order nvcc's contraction turns into the same mul/fma chains as the reference build
// (checked against its PTX: sum-of-products -> mul(second) , fma(first) , fma(third) ...), and the
// translation unit is compiled with the reference's own numeric flags (--use_fast_math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpt {

#define VPT_DEV __device__ __forceinline__

#define VPT_EPS      0.001f               // EPS, render_kernel.cu:83
#define VPT_M_INF    3.402823466e+38F     // M_INF, common/helper_math.h:41
#define VPT_PI_F     3.14159265358979323846f
#define VPT_PI_4_F   0.785398163397448309616f   // M_PI_4 (quirk Q1: HG is scaled by pi/4)
#define VPT_BLACK_EPS 1.192092896e-07F    // isBlack threshold, helper_math.h:1473

VPT_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
VPT_DEV float3 f3(float s) { return make_float3(s, s, s); }
VPT_DEV float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
VPT_DEV float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
VPT_DEV float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
VPT_DEV float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
VPT_DEV float3 operator*(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
VPT_DEV float3 operator/(float3 a, float3 b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
VPT_DEV float3 operator/(float3 a, float s) { return make_float3(a.x / s, a.y / s, a.z / s); }
VPT_DEV void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
VPT_DEV void operator*=(float3& a, float3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
VPT_DEV void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
// Pinned arithmetic.  nvcc's contraction of `a*b + c` into fma depends on the surrounding control flow (the same
// source line compiled to mul+sub in one version of the trace kernel and to fma in another, which moved a handful
// of samples by one texture-filter quantum).  Everything that feeds a tracking decision is therefore spelled with
// explicit single-rounding intrinsics, in the operation order of the reference build's PTX:
//   x*x + y*y + z*z  ->  fma(z, z, fma(x, x, y*y))      a*b - c*d  ->  fma(a, b, -(c*d))
// The PTX alone is not enough: ptxas contracts single-use `mul.ftz` results into the following `add/sub.ftz` (the PTX of
// a cross product reads mul, mul, sub; its SASS is FMUL, FFMA).  The orders here are read off the SASS (nvdisasm -gi).
VPT_DEV float  pmul(float a, float b) { return __fmul_rn(a, b); }
VPT_DEV float  padd(float a, float b) { return __fadd_rn(a, b); }
VPT_DEV float  psub(float a, float b) { return __fsub_rn(a, b); }
VPT_DEV float  pfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
VPT_DEV float  dot(float3 a, float3 b) { return pfma(a.z, b.z, pfma(a.x, b.x, pmul(a.y, b.y))); }
VPT_DEV float  length(float3 v) { return sqrtf(dot(v, v)); }
VPT_DEV float3 normalize(float3 v) { float inv = rsqrtf(dot(v, v)); return make_float3(pmul(v.x, inv), pmul(v.y, inv), pmul(v.z, inv)); }
// a*b - c*d: the reference's PTX shows mul, mul, sub, but ptxas contracts the first product (SASS: FMUL c*d; FFMA a*b - t)
VPT_DEV float3 cross(float3 a, float3 b) {
    return make_float3(pfma(a.y, b.z, -pmul(a.z, b.y)), pfma(a.z, b.x, -pmul(a.x, b.z)), pfma(a.x, b.y, -pmul(a.y, b.x)));
}
// p + d * t with one rounding per component (fma), the form every position update of the reference compiles to
VPT_DEV float3 madd3(float3 p, float3 d, float t) { return make_float3(pfma(d.x, t, p.x), pfma(d.y, t, p.y), pfma(d.z, t, p.z)); }
VPT_DEV float3 lerp3(float3 a, float3 b, float t) { return a + t * (b - a); }
VPT_DEV float3 reflect3(float3 i, float3 n) { return i - 2.0f * n * dot(n, i); }
VPT_DEV float  clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
VPT_DEV float3 fmax3(float3 a, float3 b) { return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
VPT_DEV bool   is_black(float3 v) { return length(v) < VPT_BLACK_EPS; }
VPT_DEV bool   any_nan(float3 v) { return isnan(v.x) || isnan(v.y) || isnan(v.z); }
VPT_DEV bool   any_inf(float3 v) { return isinf(v.x) || isinf(v.y) || isinf(v.z); }

// ---- Philox4x32-10, addressed by (key, block) instead of carrying the 64-byte cuRAND state ------
// Reference stream (curand_kernel.h:1022-1037, curand_philox4x32_x.h:93-197):
// curand_init(seed = pixel idx, subsequence 0, offset = iteration*4096) => key = (idx, 0),
// counter = ((iteration*4096 mod 2^32) / 4, 0, 0, 0); draw k is lane (k & 3) of block (k >> 2).
struct PhiloxBlock { uint32_t x, y, z, w; };

// one out-of-line copy: ~60 integer instructions that would otherwise be inlined at every draw site
__device__ __noinline__ PhiloxBlock philox4x32_10(uint32_t c0, uint32_t c1, uint32_t key0) {
#ifdef VPT_EXP_CHEAP_RNG   // development experiment only (breaks parity): what would a free generator buy?
    { PhiloxBlock b; b.x = (c0 ^ key0) * 2654435761u; b.y = b.x * 40503u + c1; b.z = b.y ^ (b.x >> 7); b.w = b.z * 2246822519u; return b; }
#endif
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t x0 = c0, x1 = c1, x2 = 0u, x3 = 0u, k0 = key0, k1 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, x0), lo0 = M0 * x0;
        const uint32_t hi1 = __umulhi(M1, x2), lo1 = M1 * x2;
        const uint32_t n0 = hi1 ^ x1 ^ k0, n1 = lo1, n2 = hi0 ^ x3 ^ k1, n3 = lo0;
        x0 = n0; x1 = n1; x2 = n2; x3 = n3;
        k0 += W0; k1 += W1;
    }
    PhiloxBlock b; b.x = x0; b.y = x1; b.z = x2; b.w = x3; return b;
}

// curand_uniform: x * 2^-32 + 2^-33, in (0, 1]  (curand_uniform.h:69-72)
VPT_DEV float u32_to_unit(uint32_t x) { return pfma((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f); }

struct Rng {
    uint32_t key;      // global pixel index
    uint32_t base;     // counter word 0 at draw 0: (iteration*4096 mod 2^32) >> 2
    uint32_t k;        // draws consumed so far
    uint32_t cached;   // block index held in `blk` (0xffffffff = none)
    PhiloxBlock blk;
    VPT_DEV void init(uint32_t key_, uint32_t iteration, uint32_t k0) {
        key = key_; base = (iteration * 4096u) >> 2; k = k0; cached = 0xffffffffu;
    }
    VPT_DEV float next() {
        const uint32_t b = k >> 2;
        if (b != cached) {
            // 64-bit counter increment as curand's Philox_State_Incr: carry from word 0 into word 1
            const uint32_t c0 = base + b;
            const uint32_t c1 = (c0 < base) ? 1u : 0u;
            blk = philox4x32_10(c0, c1, key);
            cached = b;
        }
        const uint32_t lane = k & 3u;
        const uint32_t v = lane == 0 ? blk.x : lane == 1 ? blk.y : lane == 2 ? blk.z : blk.w;
        ++k;
        return u32_to_unit(v);
    }
};

} // namespace vpt
