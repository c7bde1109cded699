// vpt_frame.cuh -- device helpers shared by the wavefront kernels (vpt_kernels.cu) and the level (A) entry (vpt_level_a.cu):
// the per-volume record arithmetic, the per-CTA staged scene, small frame utilities, and -- at the end -- the trace machinery
// (vpt_trace.cuh: PathState, walk_step, advance, ...).
#pragma once
#include <type_traits>
#include <cstring>
#include "vpt_walk.cuh"
#include "vpt_atmosphere.cuh"
#include "vpt_kernels.h"

namespace vpt {

constexpr uint32_t kMissSentinel = 0xffffffffu;   // planeA.w bit pattern marking a miss sample (a NaN no arithmetic produces)

__device__ __forceinline__ const vpt_octnode* oct_ptr(vpt_devptr_t p) { return reinterpret_cast<const vpt_octnode*>(p); }

// per-volume world->index affine, evaluated with the reference's own adjugate formula order
__device__ VolumeRec make_volume_rec(const vpt_gpu_vdb& g)
{
    // n_rc of the transposed matrix == xform[r-1][c-1] in memory order
    const float n11 = g.xform[0][0], n12 = g.xform[0][1], n13 = g.xform[0][2], n14 = g.xform[0][3];
    const float n21 = g.xform[1][0], n22 = g.xform[1][1], n23 = g.xform[1][2], n24 = g.xform[1][3];
    const float n31 = g.xform[2][0], n32 = g.xform[2][1], n33 = g.xform[2][2], n34 = g.xform[2][3];
    const float n41 = g.xform[3][0], n42 = g.xform[3][1], n43 = g.xform[3][2], n44 = g.xform[3][3];

    const float t11 = n23 * n34 * n42 - n24 * n33 * n42 + n24 * n32 * n43 - n22 * n34 * n43 - n23 * n32 * n44 + n22 * n33 * n44;
    const float t12 = n14 * n33 * n42 - n13 * n34 * n42 - n14 * n32 * n43 + n12 * n34 * n43 + n13 * n32 * n44 - n12 * n33 * n44;
    const float t13 = n13 * n24 * n42 - n14 * n23 * n42 + n14 * n22 * n43 - n12 * n24 * n43 - n13 * n22 * n44 + n12 * n23 * n44;
    const float t14 = n14 * n23 * n32 - n13 * n24 * n32 - n14 * n22 * n33 + n12 * n24 * n33 + n13 * n22 * n34 - n12 * n23 * n34;

    const float det = n11 * t11 + n21 * t12 + n31 * t13 + n41 * t14;
    const float idet = 1.0f / det;

    // second and third output rows of the inverse (unscaled adjugate entries)
    const float a01 = n24 * n33 * n41 - n23 * n34 * n41 - n24 * n31 * n43 + n21 * n34 * n43 + n23 * n31 * n44 - n21 * n33 * n44;
    const float a11 = n13 * n34 * n41 - n14 * n33 * n41 + n14 * n31 * n43 - n11 * n34 * n43 - n13 * n31 * n44 + n11 * n33 * n44;
    const float a21 = n14 * n23 * n41 - n13 * n24 * n41 - n14 * n21 * n43 + n11 * n24 * n43 + n13 * n21 * n44 - n11 * n23 * n44;
    const float a31 = n13 * n24 * n31 - n14 * n23 * n31 + n14 * n21 * n33 - n11 * n24 * n33 - n13 * n21 * n34 + n11 * n23 * n34;

    const float a02 = n22 * n34 * n41 - n24 * n32 * n41 + n24 * n31 * n42 - n21 * n34 * n42 - n22 * n31 * n44 + n21 * n32 * n44;
    const float a12 = n14 * n32 * n41 - n12 * n34 * n41 - n14 * n31 * n42 + n11 * n34 * n42 + n12 * n31 * n44 - n11 * n32 * n44;
    const float a22 = n12 * n24 * n41 - n14 * n22 * n41 + n14 * n21 * n42 - n11 * n24 * n42 - n12 * n21 * n44 + n11 * n22 * n44;
    const float a32 = n14 * n22 * n31 - n12 * n24 * n31 - n14 * n21 * n32 + n11 * n24 * n32 + n12 * n21 * n34 - n11 * n22 * n34;

    VolumeRec r;
    r.m[0][0] = t11 * idet; r.m[0][1] = t12 * idet; r.m[0][2] = t13 * idet; r.adj3[0] = t14;
    r.m[1][0] = a01 * idet; r.m[1][1] = a11 * idet; r.m[1][2] = a21 * idet; r.adj3[1] = a31;
    r.m[2][0] = a02 * idet; r.m[2][1] = a12 * idet; r.m[2][2] = a22 * idet; r.adj3[2] = a32;
    r.idet = idet;
    r.bmin[0] = g.vdb_info.bmin.x; r.bmin[1] = g.vdb_info.bmin.y; r.bmin[2] = g.vdb_info.bmin.z;
    r.rdim[0] = 1.0f / float(g.vdb_info.dim.x); r.rdim[1] = 1.0f / float(g.vdb_info.dim.y); r.rdim[2] = 1.0f / float(g.vdb_info.dim.z);
    r.flags = (g.vdb_info.has_color ? 1u : 0u) | (g.vdb_info.has_emission ? 2u : 0u);
    r.density_tex = g.vdb_info.density_texture; r.emission_tex = g.vdb_info.emission_texture; r.color_tex = g.vdb_info.color_texture;
    return r;
}

// =====================================================================================================
// shared helpers of the per-frame kernels
// =====================================================================================================
struct FrameShared {
    SceneTables sc;
    OctShared   oct;
    VolumeRec   vol0;      // volume 0 staged for single-volume scenes (the headline case)
};

VPT_DEV void load_frame_shared(FrameShared& fs, const SceneTables* sc_dev) {
    if (threadIdx.x == 0 && threadIdx.y == 0) fs.sc = *sc_dev;
    __syncthreads();
    // stage_octree assumes a 1-D thread index
    const int t = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    const uint4* src = reinterpret_cast<const uint4*>(fs.sc.internal);
    uint4* d = reinterpret_cast<uint4*>(fs.oct.node);
    for (int i = t; i < kOctInternalNodes * 3; i += nt) d[i] = __ldg(src + i);
    {   // 96-byte record of volume 0
        const uint4* vs = reinterpret_cast<const uint4*>(fs.sc.volumes);
        uint4* vd = reinterpret_cast<uint4*>(&fs.vol0);
        for (int i = t; i < (int)(sizeof(VolumeRec) / 16); i += nt) vd[i] = __ldg(vs + i);
    }
    __syncthreads();
}

VPT_DEV SphereRec load_sphere(const vpt_sphere* s) {
    SphereRec r;
    r.center = f3(s->center.x, s->center.y, s->center.z); r.radius = s->radius;
    r.color = f3(s->color.x, s->color.y, s->color.z); r.roughness = s->roughness;
    return r;
}

VPT_DEV float3 ld3(const vpt_f3& v) { return f3(v.x, v.y, v.z); }

// global row of a local row under the interleaved-stripe partition (identity for one rank)
VPT_DEV int global_row(const FrameGeom& g, int lr) {
    const int s = (g.stripe_h & (g.stripe_h - 1)) == 0 ? (lr >> (__ffs(g.stripe_h) - 1)) : lr / g.stripe_h;   // power-of-two stripes: no division
    return (s * g.n_ranks + g.rank) * g.stripe_h + (lr - s * g.stripe_h);
}

// The pixel word a ray carries (queue record, parked record).  One rank: the pixel index itself.  Several ranks: (local row << 16) | x --
// the Philox stream is keyed by the GLOBAL pixel every time a ray is un-parked and the sample is written at the LOCAL pixel, and from
// this form both are a shift / multiply-add away (a runtime division on the un-park path cost 5 % of the trace kernel).
VPT_DEV uint32_t ray_pixel_word(const FrameGeom& g, int lr, int x) {
    return g.n_ranks > 1 ? ((uint32_t)lr << 16) | (uint32_t)x : (uint32_t)lr * (uint32_t)g.width + (uint32_t)x;
}
VPT_DEV uint32_t ray_global_pixel(const FrameGeom& g, uint32_t w) {
    if (g.n_ranks <= 1) return w;
    return (uint32_t)global_row(g, (int)(w >> 16)) * (uint32_t)g.width + (w & 0xffffu);
}
VPT_DEV uint32_t ray_local_pixel(const FrameGeom& g, uint32_t w) {
    if (g.n_ranks <= 1) return w;
    return (w >> 16) * (uint32_t)g.width + (w & 0xffffu);
}

// radical inverse of int(xi*100) in `BASE` (reference vanDerCorput, gpu_vdb/camera.h:49-62)
template <int BASE>
VPT_DEV float van_der_corput(Rng& rng) {
    int n = int(rng.next() * 100);
    float rand_int = 0, denom = 1, invBase = 1.f / BASE;
    while (n) {
        denom *= BASE;
        rand_int = padd(rand_int, __fdividef((float)(n % BASE), denom));
        n *= invBase;
    }
    return rand_int;
}

VPT_DEV float3 aces_fit(float3 v) {                           // reference rtt_and_odt_fit, :2208-2213
    float3 a = v * (v + f3(0.0245786f)) - f3(0.000090537f);
    float3 b = v * (0.983729f * v + f3(0.4329510f)) + f3(0.238081f);
    return a / b;
}

VPT_DEV float3 mat3_mul(const float m[9], float3 v) {
    return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}

#include "vpt_texfilter.cuh"
#include "vpt_trace.cuh"
#ifndef VPT_LEVEL_A_MODULE          // the level (A) cubin exports exactly one entry: no wavefront kernels in it
#include "vpt_trace_brick.cuh"
#endif

} // namespace vpt
