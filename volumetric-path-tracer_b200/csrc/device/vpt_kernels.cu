// vpt_kernels.cu -- sm_100a wavefront kernels that replace the reference megakernel
// `volume_rt_kernel` (source/render_kernel.cu:2216-2326) for the direct integrator.
//
//   k_prepare_scene : GPU_VDB[] + OCTNode tree (the reference launch parameters) -> flat scene tables
//   k_generate      : one thread per (pixel, pass): Philox stream, blue-noise jitter, thin-lens ray,
//                     root/sphere test; misses write their sample record, hits are pushed into a
//                     warp-compacted ray queue (one atomic per warp, 32-byte records)
//   k_trace         : persistent threads; every lane owns one ray and runs ONE unified tracking
//                     step body (delta / residual-ratio / emission walks); estimator transitions are
//                     batched so the step loop stays converged; finished lanes refill from the queue
//   k_resolve       : per pixel, passes in order: environment term, NaN guard, running mean, ACES
//                     tonemap, display/raw/depth/cost writes (the tail of the reference kernel)
//   k_bn_advance    : golden-ratio advance of the 256x256 blue-noise buffer (race-free, quirk Q6)
//
// Numerics: compiled with the reference's flags (--use_fast_math); decision-relevant expressions keep
// the reference's operand order so a pixel's path is reproduced sample for sample (see vpt_math.cuh).
#include "vpt_walk.cuh"
#include "vpt_kernels.h"

namespace vpt {

// =====================================================================================================
// k_prepare_scene
// =====================================================================================================
__device__ __forceinline__ const vpt_octnode* oct_ptr(vpt_devptr_t p) { return reinterpret_cast<const vpt_octnode*>(p); }

__global__ void k_prepare_scene(const vpt_gpu_vdb* __restrict__ vols, const vpt_octnode* __restrict__ root,
                                SceneTables* out, OctInternal* internal, uint2* leaf_list, int* leaf_indices,
                                VolumeRec* vrec, int max_volumes)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int N = root->num_volumes;

    if (tid == 0) {
        out->root_pmin[0] = root->bbox.pmin.x; out->root_pmin[1] = root->bbox.pmin.y; out->root_pmin[2] = root->bbox.pmin.z;
        out->root_pmax[0] = root->bbox.pmax.x; out->root_pmax[1] = root->bbox.pmax.y; out->root_pmax[2] = root->bbox.pmax.z;
        out->max_extinction = root->max_extinction;
        out->min_extinction = root->min_extinction;
        out->num_volumes = N;
        out->single_volume = (N == 1) ? 1 : 0;
        out->internal = internal; out->leaf_list = leaf_list; out->leaf_indices = leaf_indices; out->volumes = vrec;
    }

    // internal nodes: 0 = root, 1..8 = level 1, 9..72 = level 2
    for (int j = tid; j < kOctInternalNodes; j += nthreads) {
        const vpt_octnode* n = root;
        bool exists = true;
        if (j >= 1 && j < 9) { n = oct_ptr(root->children[j - 1]); }
        else if (j >= 9) {
            const int c1 = (j - 9) >> 3, c2 = (j - 9) & 7;
            const vpt_octnode* p = oct_ptr(root->children[c1]);
            if (p->num_volumes > 0) n = oct_ptr(p->children[c2]); else exists = false;
        }
        OctInternal o;
        for (int a = 0; a < 3; ++a) { o.pmin[a] = 0.f; o.half[a] = 0.f; o.pmax[a] = 0.f; }
        o.child_empty = 0xffu; o.pad[0] = o.pad[1] = 0u;
        if (exists && n->num_volumes > 0) {
            o.pmin[0] = n->bbox.pmin.x; o.pmin[1] = n->bbox.pmin.y; o.pmin[2] = n->bbox.pmin.z;
            o.pmax[0] = n->bbox.pmax.x; o.pmax[1] = n->bbox.pmax.y; o.pmax[2] = n->bbox.pmax.z;
            const vpt_octnode* c0 = oct_ptr(n->children[0]);       // child 0 = (x-, y+, z-)
            o.half[0] = c0->bbox.pmax.x; o.half[1] = c0->bbox.pmin.y; o.half[2] = c0->bbox.pmax.z;
            uint32_t mask = 0;
            for (int c = 0; c < 8; ++c) if (oct_ptr(n->children[c])->num_volumes == 0) mask |= 1u << c;
            o.child_empty = mask;
        }
        internal[j] = o;
    }

    // leaves: volume lists (instanced scenes); stride VPT_OCT_MAX_VOLUMES per leaf
    for (int l = tid; l < kOctLeaves; l += nthreads) {
        const int c1 = l >> 6, c2 = (l >> 3) & 7, c3 = l & 7;
        uint2 lst = make_uint2((uint32_t)l * VPT_OCT_MAX_VOLUMES, 0u);
        const vpt_octnode* p1 = oct_ptr(root->children[c1]);
        if (p1->num_volumes > 0) {
            const vpt_octnode* p2 = oct_ptr(p1->children[c2]);
            if (p2->num_volumes > 0) {
                const vpt_octnode* p3 = oct_ptr(p2->children[c3]);
                const int cnt = p3->num_volumes;
                lst.y = (uint32_t)cnt;
                if (N > 1) for (int i = 0; i < cnt; ++i) leaf_indices[lst.x + i] = p3->vol_indices[i];
            }
        }
        leaf_list[l] = lst;
    }

    // per-volume world->index affine, evaluated with the reference's own adjugate formula order
    for (int v = tid; v < N && v < max_volumes; v += nthreads) {
        const vpt_gpu_vdb& g = vols[v];
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
        vrec[v] = r;
    }
}

// =====================================================================================================
// shared helpers of the per-frame kernels
// =====================================================================================================
struct FrameShared {
    SceneTables sc;
    OctShared   oct;
};

VPT_DEV void load_frame_shared(FrameShared& fs, const SceneTables* sc_dev) {
    if (threadIdx.x == 0 && threadIdx.y == 0) fs.sc = *sc_dev;
    __syncthreads();
    // stage_octree assumes a 1-D thread index
    const int t = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    const uint4* src = reinterpret_cast<const uint4*>(fs.sc.internal);
    uint4* d = reinterpret_cast<uint4*>(fs.oct.node);
    for (int i = t; i < kOctInternalNodes * 3; i += nt) d[i] = __ldg(src + i);
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
    const int s = lr / g.stripe_h;
    return (s * g.n_ranks + g.rank) * g.stripe_h + (lr - s * g.stripe_h);
}

// radical inverse of int(xi*100) in `BASE` (reference vanDerCorput, gpu_vdb/camera.h:49-62)
template <int BASE>
VPT_DEV float van_der_corput(Rng& rng) {
    int n = int(rng.next() * 100);
    float rand_int = 0, denom = 1, invBase = 1.f / BASE;
    while (n) {
        denom *= BASE;
        rand_int += (n % BASE) / denom;
        n *= invBase;
    }
    return rand_int;
}

// =====================================================================================================
// k_generate
// =====================================================================================================
__global__ void __launch_bounds__(128)
k_generate(const FrameArgs fa)
{
    __shared__ FrameShared fs;
    load_frame_shared(fs, fa.scene);
    const SceneTables& sc = fs.sc;
    const FrameGeom& g = fa.geom;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + warp * 8 + (lane & 7);
    const int lr = blockIdx.y * 4 + (lane >> 3);
    const int pass = blockIdx.z;                                 // pass within this chunk
    const int y = global_row(g, lr);
    const bool valid = (x < g.width) && (lr < g.local_rows) && (y < g.height);

    bool hit = false;
    float3 org = f3(0.f), dir = f3(0.f, 0.f, 1.f);
    uint32_t kdraws = 0;
    const uint32_t lp = (uint32_t)lr * (uint32_t)g.width + (uint32_t)x;

    if (valid) {
        const vpt_kernel_params& kp = fa.kp;
        const vpt_camera& cam = fa.cam;
        const uint32_t idx = (uint32_t)y * (uint32_t)g.width + (uint32_t)x;
        Rng rng; rng.init(idx, kp.iteration + (uint32_t)pass, 0u);

        // blue-noise jitter; the buffer holds the state at the first pass of the chunk
        const int bn_index = (y % 256) * 256 + (x % 256);
        const float3* bnbuf = reinterpret_cast<const float3*>(kp.blue_noise_buffer);
        float bnx = bnbuf[bn_index].x, bny = bnbuf[bn_index].y;
        for (int i = 0; i < pass; ++i) {
            bnx += (1.0f + sqrtf(5.0f)) / 2.0f; bnx = fmodf(bnx, 1.0f);
            bny += (1.0f + sqrtf(5.0f)) / 2.0f; bny = fmodf(bny, 1.0f);
        }
        const float u = float(x + bnx) / float(kp.resolution.x);
        const float v = float(y + bny) / float(kp.resolution.y);

        // thin-lens ray (reference camera::get_ray, camera.h:131-136)
        float3 p;
        do {
            const float a = van_der_corput<2>(rng);
            const float b = van_der_corput<3>(rng);
            p = 2.0f * f3(a, b, 0) - f3(1.0f, 1.0f, 0.0f);
        } while (dot(p, p) >= 1.0);
        const float3 rd = cam.lens_radius * p;
        const float3 offset = ld3(cam.u) * rd.x + ld3(cam.v) * rd.y;
        (void)rng.next();                                        // shutter time draw (value unused by the path)
        org = ld3(cam.origin) + offset;
        const float3 b = ld3(cam.lower_left_corner) + u * ld3(cam.horizontal) + v * ld3(cam.vertical) - ld3(cam.origin) - offset;
        dir = normalize(b);
        kdraws = rng.k;

        const SphereRec sph = load_sphere(fa.sphere);
        float t_min;
        const int obj = closest_object(sc, sph, org, dir, t_min);
        hit = (obj != 0);

        if (!hit) {
            const size_t s = (size_t)pass * g.n_local + lp;
            fa.planeA[s] = make_float4(dir.x, dir.y, dir.z, 0.0f);     // final direction, tr
            fa.planeB[s] = make_float4(0.f, 0.f, 0.f, 0.f);           // L, depth
            fa.planeC[s] = make_float4(1.f, 1.f, 1.f, 0.f);           // beta
            if (fa.planeD) fa.planeD[s] = make_float4(org.x, org.y, org.z, 0.f);
        }
    }

    // warp-compacted push: one atomic per warp, records contiguous per 8x4 pixel tile
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
        unsigned base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(fa.queue_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (hit) {
            const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
            float4* q = fa.queue + 2 * (size_t)slot;
            q[0] = make_float4(org.x, org.y, org.z, dir.x);
            q[1] = make_float4(dir.y, dir.z, __uint_as_float(lp), __uint_as_float((uint32_t)pass | (kdraws << 8)));
        }
    }
}

// =====================================================================================================
// k_trace -- persistent wavefront over the hit queue
// =====================================================================================================
enum WalkMode : int { W_NONE = 0, W_DELTA = 1, W_RATIO = 2, W_EMIT = 3 };
enum Phase : int {
    PH_IDLE = 0,          // lane has no ray
    PH_BOUNCE,            // top of a ray_depth iteration
    PH_VOL_ITER,          // start a delta-tracking walk
    PH_AFTER_DELTA,       // delta walk ended
    PH_VOL_DONE,          // volume_depth loop finished
    PH_AFTER_SUN,         // ratio walk toward the sun ended
    PH_POINT_NEXT,        // next point-light iteration
    PH_AFTER_POINT,       // ratio walk toward a point light ended
    PH_EMISSION,          // maybe start the emission walk
    PH_AFTER_EMIT,
    PH_AFTER_VOLUME,      // second closest-object test of the bounce
    PH_AFTER_SPHERE_TR,   // ratio walk from the sphere toward the sun ended
    PH_FINISH
};
enum ExitReason : int { EX_NONE = 0, EX_OUTSIDE, EX_DISTANCE, EX_SCATTER, EX_TR_DONE };

struct PathState {
    // ray
    float3 pos, dir;
    float3 org;           // camera-ray origin (depth reference + default env_pos)
    float3 env_pos;
    float3 beta, L;
    float  alpha;         // the reference's `tr` out-parameter (accumulated density, capped at 1)
    float  depth;
    // walk
    float3 wpos, wdir;    // position / direction of the running walk (Tr and emission walk on copies)
    float  t, distance;
    float  trv;           // running residual-ratio transmittance (all three channels equal)
    float  T_c;
    float3 emis;          // emission walk accumulator
    float3 thr;           // throughput returned by the last delta walk
    int    mode, phase, exit_reason;
    // loop counters
    int    rd, vd, light_budget, light_index;
    float3 Ld;            // point-light accumulator
    float3 sph_normal;
    bool   mi, first_walk, geo;
    int    obj;
    uint32_t lp, pass;
    uint32_t nlook;       // density / emission lookups issued by this lane (statistics)
    Rng    rng;
};

struct TraceConsts {
    float inv_max, inv_mult, sigma_c, sigma_r_inv;
    float3 sun_dir;
};

// One unified tracking step.  Returns with st.mode == W_NONE when the walk has ended (st.exit_reason set).
VPT_DEV void walk_step(PathState& st, const FrameShared& fs, const FrameArgs& fa, const TraceConsts& tc, const SphereRec& sph)
{
    const SceneTables& sc = fs.sc;
    const vpt_kernel_params& kp = fa.kp;
    const int leaf = oct_locate_or_skip(fs.oct, sc, st.wpos, st.wdir);
    if (leaf == -2) return;                                   // skipped an empty node, no draw consumed
    if (leaf == -1) { st.mode = W_NONE; st.exit_reason = EX_OUTSIDE; return; }

    if (st.mode == W_DELTA) {
        // distance to the box exit (or to the sphere) from the CURRENT position, every step (:1647-1651)
        float t_min, t_max, geo_dist = .0f;
        aabb_intersect(sc.root_pmin, sc.root_pmax, st.wpos, st.wdir, t_min, st.distance);
        if (sphere_intersect(sph, st.wpos, st.wdir, geo_dist, t_max)) { st.distance = geo_dist; st.geo = true; }
        st.t -= logf(1 - st.rng.next()) * tc.inv_max * tc.inv_mult;
        if (st.t >= st.distance) { st.mode = W_NONE; st.exit_reason = EX_DISTANCE; return; }
    } else if (st.mode == W_RATIO) {
        st.t -= logf(1 - st.rng.next()) * tc.sigma_r_inv * kp.tr_depth;
        if (st.t >= st.distance) { st.mode = W_NONE; st.exit_reason = EX_DISTANCE; return; }
    } else {
        float inv_max_density = 1 / sc.max_extinction;
        st.t -= logf(1 - st.rng.next()) * inv_max_density * kp.tr_depth / kp.extinction.x;
    }

    st.wpos += st.wdir * st.t;                                // cumulative t, never reset (quirk Q2)
    if (!aabb_contains(sc.root_pmin, sc.root_pmax, st.wpos)) { st.mode = W_NONE; st.exit_reason = EX_OUTSIDE; return; }

    st.nlook++;
    if (st.mode == W_EMIT) {
        st.emis += leaf_emission(sc, leaf, st.wpos, reinterpret_cast<const float3*>(kp.emission_texture), kp.emission_pivot, kp.emission_scale);
        return;
    }

    const float density = leaf_density(sc, leaf, st.wpos);
    if (st.mode == W_DELTA) {
        const float3 Cd = leaf_color(sc, leaf, st.wpos);
        const int index = int(floorf(fminf(fmaxf((density * tc.inv_max * 255.0f / kp.emission_pivot), 0.0f), 255.0f)));
        const float3 density_color = reinterpret_cast<const float3*>(kp.density_color_texture)[index];
        if (st.alpha < 1.0f) st.alpha += density;
        if (density * tc.inv_max > st.rng.next()) {
            st.thr = (ld3(kp.albedo) * Cd * density_color / ld3(kp.extinction)) * float(kp.energy_inject);
            st.mode = W_NONE; st.exit_reason = EX_SCATTER;
        }
    } else {
        st.trv *= 1 - ((density - tc.sigma_c) * tc.sigma_r_inv);
        if (length(f3(st.trv)) < VPT_EPS) { st.mode = W_NONE; st.exit_reason = EX_TR_DONE; }
    }
}

// Set up a residual-ratio walk from (p, d) (reference Tr prologue, :1150-1167).
// Returns false when the transmittance is already known (then `result` holds it).
VPT_DEV bool begin_ratio_walk(PathState& st, const FrameShared& fs, const TraceConsts& tc, const SphereRec& sph,
                              float3 p, float3 d, float& result)
{
    const SceneTables& sc = fs.sc;
    float t_min, t_max, geo_dist = .0f, distance = .0f;
    if (!aabb_contains(sc.root_pmin, sc.root_pmax, p)) {
        if (aabb_intersect(sc.root_pmin, sc.root_pmax, p, d, t_min, t_max)) p += d * (t_min + VPT_EPS);
        else { result = 1.0f; return false; }
    }
    aabb_intersect(sc.root_pmin, sc.root_pmax, p, d, t_min, distance);
    if (sphere_intersect(sph, p, d, geo_dist, t_max)) { result = 0.0f; return false; }
    st.T_c = expf(-tc.sigma_c * distance);
    st.wpos = p; st.wdir = d; st.t = 0.0f; st.distance = distance; st.trv = 1.0f;
    st.mode = W_RATIO;
    return true;
}

VPT_DEV float finish_ratio_walk(const PathState& st) { return clampf(st.trv * st.T_c, .0f, 1.0f); }

// Estimator transitions.  Runs until the lane either starts a walk (st.mode != W_NONE) or retires.
VPT_DEV void transition(PathState& st, const FrameShared& fs, const FrameArgs& fa, const TraceConsts& tc, const SphereRec& sph)
{
    const SceneTables& sc = fs.sc;
    const vpt_kernel_params& kp = fa.kp;
    const FrameGeom& g = fa.geom;

    for (;;) {
        switch (st.phase) {
        case PH_BOUNCE: {
            if (st.rd > kp.ray_depth) { st.phase = PH_FINISH; break; }
            float t_min;
            st.obj = closest_object(sc, sph, st.pos, st.dir, t_min);
            if (st.first_walk && st.obj != 1) {                    // depth pass without a volume walk (:1883-1888)
                if (st.obj == 2) st.depth = length(st.org - (st.pos + st.dir * t_min));
                st.first_walk = false;
            }
            if (st.obj == 0) { st.phase = PH_FINISH; break; }      // nothing ahead: every later bounce is a no-op
            if (st.obj == 1) {
                st.pos += st.dir * (t_min + VPT_EPS);
                st.vd = 1;
                st.phase = PH_VOL_ITER;
            } else st.phase = PH_AFTER_VOLUME;
            break;
        }
        case PH_VOL_ITER: {
            if (st.vd > kp.volume_depth) { st.phase = PH_VOL_DONE; break; }
            st.mi = false;
            st.wpos = st.pos; st.wdir = st.dir; st.t = 0.0f; st.distance = .0f; st.geo = false;
            st.mode = W_DELTA; st.phase = PH_AFTER_DELTA;
            return;
        }
        case PH_AFTER_DELTA: {
            st.pos = st.wpos;                                      // `sample` advances the caller's ray_pos
            if (st.exit_reason == EX_SCATTER) { st.beta *= st.thr; st.mi = true; }
            else st.beta *= f3(1.0f);
            if (st.exit_reason == EX_DISTANCE) st.obj = 2;          // compiled reference sets obj = 2 on every distance exit (Q4)
            if (st.first_walk) {
                st.depth = st.mi ? length(st.org - st.pos) : .0f;
                // the reference runs this identical walk twice (depth pass + integrator) and accumulates
                // `tr` in both; the second replay adds the same densities again while tr < 1
                if (st.alpha < 1.0f) st.alpha += st.alpha;
                st.first_walk = false;
            }
            if (is_black(st.beta) || st.obj == 2) { st.phase = PH_VOL_DONE; break; }
            if (st.mi) hg_sample(st.dir, st.rng, kp.phase_g1);
            st.vd++;
            st.phase = PH_VOL_ITER;
            break;
        }
        case PH_VOL_DONE: {
            if (st.mi) {
                float res;
                if (begin_ratio_walk(st, fs, tc, sph, st.pos, tc.sun_dir, res)) { st.phase = PH_AFTER_SUN; return; }
                st.trv = res; st.T_c = 1.0f; st.exit_reason = EX_NONE;
                st.phase = PH_AFTER_SUN;
                // fall through with a "finished" walk whose result is res
                st.mode = W_NONE;
                // encode the known result so that finish_ratio_walk returns it unchanged
                // (res is 0 or 1, clamp(res * 1) == res)
                break;
            }
            st.phase = PH_EMISSION;
            break;
        }
        case PH_AFTER_SUN: {
            const float tr = finish_ratio_walk(st);
            const float cos_theta = dot(st.dir, tc.sun_dir);
            const float phase_pdf = hg_phase(cos_theta, kp.phase_g1);
            const float3 Ld = f3(tr) * phase_pdf;
            st.L += Ld * ld3(kp.sun_color) * kp.sun_mult * st.beta;
            if (fa.lights.num_lights > 0) { st.Ld = f3(.0f); st.light_budget = 10; st.phase = PH_POINT_NEXT; }
            else st.phase = PH_EMISSION;
            break;
        }
        case PH_POINT_NEXT: {
            if (st.light_budget < 0) { st.L += st.Ld * st.beta; st.phase = PH_EMISSION; break; }
            const vpt_point_light* lp = reinterpret_cast<const vpt_point_light*>(fa.lights.light_ptr);
            st.light_index = int(floorf(st.rng.next() * fa.lights.num_lights));
            const float3 d = normalize(ld3(lp[st.light_index].pos) - st.pos);
            float res;
            if (begin_ratio_walk(st, fs, tc, sph, st.pos, d, res)) { st.phase = PH_AFTER_POINT; return; }
            st.trv = res; st.T_c = 1.0f; st.mode = W_NONE;
            st.phase = PH_AFTER_POINT;
            break;
        }
        case PH_AFTER_POINT: {
            const float tr = finish_ratio_walk(st);
            if (st.light_budget < (int)fa.lights.num_lights) {       // reference point_light::Le, light.h:104-121
                const vpt_point_light& pl = reinterpret_cast<const vpt_point_light*>(fa.lights.light_ptr)[st.light_index];
                const float3 lpos = ld3(pl.pos);
                const float3 wi = normalize(lpos - st.pos);
                const float cos_theta = dot(st.dir, wi);
                const float phase_pdf = hg_phase(cos_theta, kp.phase_g1);
                const float sqr_dist = length(lpos * lpos - st.pos * st.pos);
                const float falloff = 1 / sqr_dist;
                st.Ld += ld3(pl.color) * pl.power * f3(tr) * phase_pdf * falloff;
            }
            st.light_budget--;
            st.phase = PH_POINT_NEXT;
            break;
        }
        case PH_EMISSION: {
            if (kp.emission_scale > 0 && st.mi) {
                st.wpos = st.pos; st.wdir = st.dir; st.t = 0.0f; st.emis = f3(.0f);
                st.mode = W_EMIT; st.phase = PH_AFTER_EMIT;
                return;
            }
            st.phase = PH_AFTER_VOLUME;
            break;
        }
        case PH_AFTER_EMIT: {
            st.L += st.emis;
            st.phase = PH_AFTER_VOLUME;
            break;
        }
        case PH_AFTER_VOLUME: {
            float t_min;
            st.obj = closest_object(sc, sph, st.pos, st.dir, t_min);
            if (st.obj == 2) {                                     // diffuse/mirror bounce off the reference sphere (:1807-1834)
                st.pos += st.dir * t_min;
                const float3 normal = normalize((st.pos - sph.center) / sph.radius);
                const float3 nl = dot(normal, st.dir) < 0 ? normal : normal * -1;
                const float phi = 2 * VPT_PI_F * st.rng.next();
                const float r2 = st.rng.next();
                const float r2s = sqrtf(r2);
                const float3 w = normalize(nl);
                const float3 u = normalize(cross((fabs(w.x) > .1 ? f3(0, 1, 0) : f3(1, 0, 0)), w));
                const float3 v = cross(w, u);
                const float3 hemisphere_dir = normalize(u * cosf(phi) * r2s + v * sinf(phi) * r2s + w * sqrtf(1 - r2));
                const float3 ref = reflect3(st.dir, nl);
                st.dir = lerp3(ref, hemisphere_dir, sph.roughness);
                st.pos += normal * VPT_EPS;
                st.beta *= sph.color;
                st.sph_normal = normal;
                float res;
                if (begin_ratio_walk(st, fs, tc, sph, st.pos, tc.sun_dir, res)) { st.phase = PH_AFTER_SPHERE_TR; return; }
                st.trv = res; st.T_c = 1.0f; st.mode = W_NONE;
                st.phase = PH_AFTER_SPHERE_TR;
                break;
            }
            st.rd++;
            st.phase = PH_BOUNCE;
            break;
        }
        case PH_AFTER_SPHERE_TR: {
            const float v_tr = finish_ratio_walk(st);
            st.L += ld3(kp.sun_color) * kp.sun_mult * f3(v_tr) * fmaxf(dot(tc.sun_dir, st.sph_normal), .0f) * st.beta;
            st.env_pos = st.pos;
            st.rd++;
            st.phase = PH_BOUNCE;
            break;
        }
        case PH_FINISH: {
            const size_t s = (size_t)st.pass * g.n_local + st.lp;
            fa.planeA[s] = make_float4(st.dir.x, st.dir.y, st.dir.z, st.alpha);
            fa.planeB[s] = make_float4(st.L.x, st.L.y, st.L.z, st.depth);
            fa.planeC[s] = make_float4(st.beta.x, st.beta.y, st.beta.z, 1.f);
            if (fa.planeD) fa.planeD[s] = make_float4(st.env_pos.x, st.env_pos.y, st.env_pos.z, 0.f);
            st.phase = PH_IDLE;
            return;
        }
        default:
            return;
        }
    }
}

template <int kServiceThreshold>
__global__ void __launch_bounds__(kTraceThreads, 2)
k_trace(const FrameArgs fa)
{
    __shared__ FrameShared fs;
    load_frame_shared(fs, fa.scene);
    const SceneTables& sc = fs.sc;
    const vpt_kernel_params& kp = fa.kp;
    const FrameGeom& g = fa.geom;
    const int lane = threadIdx.x & 31;

    TraceConsts tc;
    tc.inv_max = 1.0f / sc.max_extinction;
    tc.inv_mult = 1.0f / kp.density_mult;
    tc.sigma_c = sc.min_extinction;
    tc.sigma_r_inv = 1.0f / (sc.max_extinction - tc.sigma_c);
    tc.sun_dir = sun_direction(kp.azimuth, kp.elevation);
    const SphereRec sph = load_sphere(fa.sphere);

    const unsigned q_count = *fa.queue_count;
    PathState st;
    st.phase = PH_IDLE; st.mode = W_NONE; st.nlook = 0;
    bool queue_dry = false;
    uint32_t lane_steps = 0, warp_iters = 0, lane_trans = 0, warp_trans = 0;

    for (;;) {
        // ---- refill idle lanes from the ray queue (one atomic per warp) ----
        const unsigned idle = __ballot_sync(0xffffffffu, st.phase == PH_IDLE);
        if (idle && !queue_dry) {
            unsigned base = 0;
            const int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(fa.queue_head, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + __popc(idle) >= q_count) queue_dry = true;
            if (st.phase == PH_IDLE) {
                const unsigned slot = base + __popc(idle & ((1u << lane) - 1u));
                if (slot < q_count) {
                    const float4 r0 = __ldg(fa.queue + 2 * (size_t)slot), r1 = __ldg(fa.queue + 2 * (size_t)slot + 1);
                    st.org = f3(r0.x, r0.y, r0.z); st.dir = f3(r0.w, r1.x, r1.y);
                    st.lp = __float_as_uint(r1.z);
                    const uint32_t pk = __float_as_uint(r1.w);
                    st.pass = pk & 0xffu;
                    const uint32_t lr = st.lp / (uint32_t)g.width, x = st.lp - lr * (uint32_t)g.width;
                    const uint32_t idx = (uint32_t)global_row(g, (int)lr) * (uint32_t)g.width + x;
                    st.rng.init(idx, kp.iteration + st.pass, pk >> 8);
                    st.pos = st.org; st.env_pos = st.org;
                    st.beta = f3(1.0f); st.L = f3(.0f); st.alpha = .0f; st.depth = .0f;
                    st.mi = false; st.first_walk = true; st.rd = 1; st.obj = 0;
                    st.phase = PH_BOUNCE; st.mode = W_NONE;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, st.phase != PH_IDLE) == 0u) break;

        // ---- estimator transitions for every lane that is between walks ----
        if (st.phase != PH_IDLE && st.mode == W_NONE) { transition(st, fs, fa, tc, sph); lane_trans++; }
        warp_trans++;

        // ---- converged step loop: keep stepping while enough lanes are inside a walk ----
        for (;;) {
            const unsigned walking = __ballot_sync(0xffffffffu, st.mode != W_NONE);
            if (walking == 0u) break;
            const unsigned waiting = __ballot_sync(0xffffffffu, st.mode == W_NONE && (st.phase != PH_IDLE || !queue_dry));
            if (waiting != 0u && __popc(walking) < kServiceThreshold) break;
            if (st.mode != W_NONE) { walk_step(st, fs, fa, tc, sph); lane_steps++; }
            warp_iters++;
        }
    }

    if (fa.counters) {                                        // optional statistics (one atomic set per warp)
        unsigned long long a = st.nlook, b = lane_steps, c = lane_trans;
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            atomicAdd(fa.counters + 0, a); atomicAdd(fa.counters + 1, b); atomicAdd(fa.counters + 2, (unsigned long long)warp_iters);
            atomicAdd(fa.counters + 3, c); atomicAdd(fa.counters + 4, (unsigned long long)warp_trans);
        }
    }
}

// =====================================================================================================
// k_resolve
// =====================================================================================================
VPT_DEV float3 aces_fit(float3 v) {                           // reference rtt_and_odt_fit, :2208-2213
    float3 a = v * (v + f3(0.0245786f)) - f3(0.000090537f);
    float3 b = v * (0.983729f * v + f3(0.4329510f)) + f3(0.238081f);
    return a / b;
}

VPT_DEV float3 mat3_mul(const float m[9], float3 v) {
    return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}

__global__ void __launch_bounds__(256)
k_resolve(const FrameArgs fa, const int n_passes, const int sampled, const int write_display)
{
    const FrameGeom& g = fa.geom;
    const vpt_kernel_params& kp = fa.kp;
    const vpt_camera& cam = fa.cam;
    const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= (uint32_t)g.n_local) return;
    const int lr = lp / g.width, x = lp - lr * g.width;
    const int y = global_row(g, lr);
    if (y >= g.height) return;

    float3* accum_buf = reinterpret_cast<float3*>(kp.accum_buffer);
    float3* cost_buf = reinterpret_cast<float3*>(kp.cost_buffer);
    float*  depth_buf = reinterpret_cast<float*>(kp.depth_buffer);
    // output buffers are indexed by local pixel (== global pixel index for a single rank)
    float3 accum = f3(0.f), costv = f3(0.f); float depthv = 0.f;
    const bool have_prev = kp.iteration > 0;
    if (have_prev || !sampled) { accum = accum_buf[lp]; costv = cost_buf[lp]; depthv = depth_buf[lp]; }
    float tr = .0f;

    for (int p = 0; p < n_passes; ++p) {
        const uint32_t iteration = kp.iteration + (uint32_t)p;
        float3 value = f3(1.0f);                                // WHITE when nothing is sampled (:2248)
        float depth = .0f;
        tr = .0f;
        if (sampled) {
            const size_t s = (size_t)p * g.n_local + lp;
            const float4 A = fa.planeA[s], B = fa.planeB[s], C = fa.planeC[s];
            const float3 ray_dir = f3(A.x, A.y, A.z), beta = f3(C.x, C.y, C.z);
            float3 L = f3(B.x, B.y, B.z);
            depth = B.w;
            tr = A.w;
            if (kp.environment_type == 0) {
                // Bruneton sky (reference sample_atmosphere): not part of this build yet; the host API
                // refuses environment_type == 0, so this branch is never taken.
            } else {
                const float4 texval = tex2D<float4>((cudaTextureObject_t)kp.env_tex,
                    atan2f(ray_dir.z, ray_dir.x) * (float)(0.5 / 3.14159265358979323846) + 0.5f,
                    acosf(fmaxf(fminf(ray_dir.y, 1.0f), -1.0f)) * (float)(1.0 / 3.14159265358979323846));
                L += f3(texval.x, texval.y, texval.z) * ld3(kp.sky_color) * beta * (1.0f / (4.0f * VPT_PI_F));   // isotropic() returns the folded constant
            }
            tr = fminf(tr, 1.0f);
            value = L;
        }
        // NaN / Inf guard (:2263-2264)
        if (any_nan(value) || any_inf(value)) value = accum;
        if (isnan(tr) || isinf(tr)) tr = 1.0f;

        float aof = 1 / cam.lens_radius;
        aof = clampf(aof, .0f, 3.402823466e+38F);
        if (cam.viz_dof) {
            if (depth > (cam.focus_dist + aof)) value = lerp3(value, f3(1.f, 0.f, 0.f), 0.5f);
            if (depth < (cam.focus_dist - aof)) value = lerp3(value, f3(0.f, 0.f, 1.f), 0.5f);
            if (depth > (cam.focus_dist - aof) && depth < (cam.focus_dist + aof)) value = lerp3(value, f3(0.f, 1.f, 0.f), 0.5f);
        }

        if (iteration == 0) { accum = value; costv = f3(0.f); depthv = depth; }
        else if (iteration < kp.max_interactions) {
            accum = accum + (value - accum) / (float)(iteration + 1);
            costv = costv + (f3(0.f) - costv) / (float)(iteration + 1);
            depthv = depthv + (depth - depthv) / (float)(iteration + 1);
        }
    }
    accum_buf[lp] = accum; cost_buf[lp] = costv; depth_buf[lp] = depthv;

    if (write_display) {
        const float aces_in[9]  = { 0.59719f, 0.35458f, 0.04823f, 0.07600f, 0.90834f, 0.01566f, 0.02840f, 0.13383f, 0.83777f };
        const float aces_out[9] = { 1.60475f, -0.53108f, -0.07367f, -0.10208f, 1.10813f, -0.00605f, -0.00327f, -0.07276f, 1.07602f };
        float3 val = mat3_mul(aces_in, accum);
        val = aces_fit(val);
        val = mat3_mul(aces_out, val) * kp.exposure_scale;
        const unsigned int r = (unsigned int)(255.0f * fminf(powf(fmaxf(val.x, 0.0f), (float)(1.0 / 2.2)), 1.0f));
        const unsigned int gg = (unsigned int)(255.0f * fminf(powf(fmaxf(val.y, 0.0f), (float)(1.0 / 2.2)), 1.0f));
        const unsigned int b = (unsigned int)(255.0f * fminf(powf(fmaxf(val.z, 0.0f), (float)(1.0 / 2.2)), 1.0f));
        reinterpret_cast<unsigned int*>(kp.display_buffer)[lp] = 0xff000000 | (r << 16) | (gg << 8) | b;
        reinterpret_cast<float4*>(kp.raw_buffer)[lp] = make_float4(val.x, val.y, val.z, tr);
    }
}

// =====================================================================================================
// k_bn_advance: blue-noise buffer += golden ratio (mod 1), `n` passes worth (:2319-2325)
// =====================================================================================================
__global__ void k_bn_advance(float3* bn, int n)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 256 * 256) return;
    float3 val = bn[idx];
    for (int i = 0; i < n; ++i) {
        val.x += (1.0f + sqrtf(5.0f)) / 2.0f; val.y += (1.0f + sqrtf(5.0f)) / 2.0f; val.z += (1.0f + sqrtf(5.0f)) / 2.0f;
        val.x = fmodf(val.x, 1.0f); val.y = fmodf(val.y, 1.0f); val.z = fmodf(val.z, 1.0f);
    }
    bn[idx] = val;
}

// gather permutation for the multi-GPU path: rank-contiguous stripes -> full frame
__global__ void k_unpermute(const uint8_t* __restrict__ gathered, uint8_t* __restrict__ full, FrameGeom g, int elem_bytes)
{
    // one thread per (rank, local pixel) element of 4-byte words
    const size_t words_per_px = elem_bytes / 4;
    const size_t total = (size_t)g.n_ranks * g.n_local * words_per_px;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t px = i / words_per_px, w = i - px * words_per_px;
        const int rank = (int)(px / g.n_local);
        const size_t lp = px - (size_t)rank * g.n_local;
        const int lr = (int)(lp / g.width), x = (int)(lp - (size_t)lr * g.width);
        const int s = lr / g.stripe_h;
        const int y = (s * g.n_ranks + rank) * g.stripe_h + (lr - s * g.stripe_h);
        if (y < g.height)
            reinterpret_cast<uint32_t*>(full)[((size_t)y * g.width + x) * words_per_px + w] = reinterpret_cast<const uint32_t*>(gathered)[i];
    }
}

// =====================================================================================================
// host-callable launchers
// =====================================================================================================
cudaError_t launch_prepare_scene(const vpt_gpu_vdb* vols, const vpt_octnode* root, SceneTables* out, OctInternal* internal,
                                 uint2* leaf_list, int* leaf_indices, VolumeRec* vrec, int max_volumes, cudaStream_t s)
{
    k_prepare_scene<<<4, 256, 0, s>>>(vols, root, out, internal, leaf_list, leaf_indices, vrec, max_volumes);
    return cudaGetLastError();
}

cudaError_t launch_generate(const FrameArgs& fa, int n_passes, cudaStream_t s)
{
    dim3 grid((fa.geom.width + 31) / 32, (fa.geom.local_rows + 3) / 4, n_passes);
    k_generate<<<grid, 128, 0, s>>>(fa);
    return cudaGetLastError();
}

cudaError_t launch_trace(const FrameArgs& fa, int n_ctas, int service_threshold, cudaStream_t s)
{
    switch (service_threshold) {
    case 8:  k_trace<8><<<n_ctas, kTraceThreads, 0, s>>>(fa); break;
    case 16: k_trace<16><<<n_ctas, kTraceThreads, 0, s>>>(fa); break;
    case 24: k_trace<24><<<n_ctas, kTraceThreads, 0, s>>>(fa); break;
    case 32: k_trace<32><<<n_ctas, kTraceThreads, 0, s>>>(fa); break;
    default: k_trace<20><<<n_ctas, kTraceThreads, 0, s>>>(fa); break;
    }
    return cudaGetLastError();
}

int trace_max_ctas_per_sm(int service_threshold)
{
    int n = 0;
    switch (service_threshold) {
    case 8:  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<8>, kTraceThreads, 0); break;
    case 16: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<16>, kTraceThreads, 0); break;
    case 24: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<24>, kTraceThreads, 0); break;
    case 32: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<32>, kTraceThreads, 0); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<20>, kTraceThreads, 0); break;
    }
    return n;
}

cudaError_t launch_resolve(const FrameArgs& fa, int n_passes, int sampled, int write_display, cudaStream_t s)
{
    const int threads = 256;
    k_resolve<<<(fa.geom.n_local + threads - 1) / threads, threads, 0, s>>>(fa, n_passes, sampled, write_display);
    return cudaGetLastError();
}

cudaError_t launch_bn_advance(void* bn, int n, cudaStream_t s)
{
    k_bn_advance<<<256, 256, 0, s>>>(reinterpret_cast<float3*>(bn), n);
    return cudaGetLastError();
}

cudaError_t launch_unpermute(const void* gathered, void* full, const FrameGeom& g, int elem_bytes, cudaStream_t s)
{
    k_unpermute<<<1184, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(gathered), reinterpret_cast<uint8_t*>(full), g, elem_bytes);
    return cudaGetLastError();
}

} // namespace vpt
