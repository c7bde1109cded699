// vpt_trace.cuh -- k_trace: persistent wavefront over the hit queue, three rays per lane.
// (included from vpt_kernels.cu inside `namespace vpt`)
//
// Divergence is the enemy of this path: a ray needs a different heavy operation every few hundred
// instructions (closest-object test, tracking step, HG resampling, transmittance set-up, bookkeeping)
// and walks have geometric length.  One-ray-per-lane designs measured 10-12 active lanes per issued
// instruction (profiles/).  Here every LANE owns kSlots = 3 rays, parked in shared memory (34 words each,
// word-major so that lane l only ever touches bank l: conflict-free, no inter-lane hand-over) and tagged,
// in registers, with the operation they wait for:
//
//     OP_STEP     one unified tracking step (delta / residual-ratio / emission walk)
//     OP_CLOSEST  nearest of {octree root box, sphere}            (reference get_closest_object)
//     OP_TRBEGIN  set-up of a residual-ratio transmittance walk   (reference Tr prologue)
//     OP_GLUE     integrator bookkeeping after a walk ended (HG resampling, NEE results, retirement)      OP_IDLE  free slot
//
// Each round the warp votes between two code sites: the STEPPING LOOP (a lane steps one of its walkers; when that walk
// ends it parks the ray and swaps in its next walker, so the step body keeps most lanes busy until the warp runs out of
// walkers) and a SERVICE ROUND (one parked ray per lane runs [closest-object test] -> the integrator's control flow
// `advance` -> [transmittance set-up], each block once in the code).  Free slots are refilled from the global ray queue
// with one atomic per warp.  The walk path moves 13 + 10 words of the record per swap and carries the packed flag word
// as is; the kernel is instantiated per integrator and in a "lean" form without multi-volume / emission / point-light
// code for the common scene (the kernel is fetch / issue bound: 30 % less code measured 20 % faster).
//
// Control flow restated from the reference direct integrator (render_kernel.cu:1760-1857) with three
// bit-exact shortcuts: (1) the depth pass replays the integrator's first walk on a copy of the RNG
// (:1859-1889), so it is taken from that walk instead of being run again; (2) the closest-object test
// at the end of a bounce and the one at the top of the next have identical inputs unless the sphere
// branch ran, so the result is reused (and the very first one comes from k_generate); (3) a bounce that
// finds nothing ahead makes every later bounce a no-op, so the path retires there.
#pragma once

enum WalkMode : int { W_DELTA = 1, W_RATIO = 2, W_EMIT = 3 };
enum Op : int { OP_IDLE = 0, OP_GLUE, OP_STEP, OP_CLOSEST, OP_TRBEGIN, OP_FINISH };
enum Phase : int {
    PH_BOUNCE_TOP = 0, PH_TOP_HAVE, PH_VOL_ITER, PH_AFTER_DELTA, PH_AFTER_HG, PH_VOL_DONE, PH_AFTER_TR,
    PH_POINT_NEXT, PH_EMISSION, PH_AFTER_EMIT, PH_AFTER_VOLUME, PH_AFTERVOL_HAVE, PH_SPHERE
};
enum ExitReason : int { EX_NONE = 0, EX_OUTSIDE, EX_DISTANCE, EX_SCATTER, EX_TR_DONE };
enum TrKind : int { TR_SUN = 0, TR_POINT = 1, TR_SPHERE = 2 };

// rays per lane: 3 (4 CTAs per SM, <= 128 registers) or 2 (5 CTAs per SM, <= 96 registers: more warps in flight, the better
// trade when every look-up misses all caches -- 1024^3 grid: trace 157 -> 99 ms; slower when the look-ups hit L2)
constexpr int trace_min_ctas(int slots) { return slots == 2 ? 5 : kTraceMinCtas; }
constexpr int kRayWords = 34;             // 32-bit words per ray in shared memory
constexpr int kTraceWarps = kTraceThreads / 32;

struct PathState {
    float3 pos, dir;      // the integrator's ray
    float3 org;           // camera-ray origin (depth reference, default env_pos); not stored: pinhole -> cam.origin, else queue_org[qslot]
    float3 beta, L;
    float  alpha;         // the reference's `tr` out-parameter (accumulated density, capped at 1 at the end)
    float  depth;
    float3 wpos, wdir;    // running walk (Tr and the emission walk run on copies of the ray)
    float  t, distance;
    float  trv, T_c;      // residual-ratio running value and control-variate factor
    float3 aux;           // emission accumulator | point-light accumulator | sphere normal (never live together)
    float  tmin_c;        // cached closest-object result
    int    obj_c;
    int    op, phase, mode, exit_reason, tr_kind;
    int    rd, vd, light_budget, light_index;
    bool   mi, first_walk, have_closest, sphere_bounced;
    bool   sphere_free;   // the running walk's line provably misses the sphere: its per-step intersection tests are skipped
    uint32_t lp, pass, qslot;
    uint32_t packed;      // word 30 as parked (walk path only: the stepping loop edits mode-independent bits in place)
    Rng    rng;
};

struct TraceConsts {
    float inv_max, inv_mult, sigma_c, sigma_r_inv;
    float3 sun_dir;
};

// ---- ray record <-> registers ------------------------------------------------------------------------------
// pool layout: word w of slot j of lane l at pool[w * kPool + j * 32 + l]  (bank == lane: conflict-free)
template <int kPool>                                             // rays per warp = 32 * rays per lane
struct PoolView {
    float* w;                                                    // this warp's kRayWords * kPool words, already offset by the lane
    VPT_DEV float& f(int word, int slot) const { return w[word * kPool + slot * 32]; }
    VPT_DEV uint32_t& u(int word, int slot) const { return reinterpret_cast<uint32_t*>(w)[word * kPool + slot * 32]; }
};

template <class PV>
VPT_DEV void store_ray(const PV& pv, int s, const PathState& st)
{
    pv.f(0, s) = st.pos.x;  pv.f(1, s) = st.pos.y;  pv.f(2, s) = st.pos.z;
    pv.f(3, s) = st.dir.x;  pv.f(4, s) = st.dir.y;  pv.f(5, s) = st.dir.z;
    pv.f(6, s) = st.L.x;    pv.f(7, s) = st.L.y;    pv.f(8, s) = st.L.z;
    pv.f(9, s) = st.beta.x; pv.f(10, s) = st.beta.y; pv.f(11, s) = st.beta.z;
    pv.f(12, s) = st.wpos.x; pv.f(13, s) = st.wpos.y; pv.f(14, s) = st.wpos.z;
    pv.f(15, s) = st.wdir.x; pv.f(16, s) = st.wdir.y; pv.f(17, s) = st.wdir.z;
    pv.f(18, s) = st.aux.x; pv.f(19, s) = st.aux.y; pv.f(20, s) = st.aux.z;
    pv.f(21, s) = st.alpha; pv.f(22, s) = st.depth; pv.f(23, s) = st.t; pv.f(24, s) = st.distance;
    pv.f(25, s) = st.trv;   pv.f(26, s) = st.T_c;   pv.f(27, s) = st.tmin_c;
    pv.u(28, s) = st.rng.k; pv.u(29, s) = st.lp;
    pv.u(30, s) = (uint32_t)st.phase | ((uint32_t)st.mode << 4) | ((uint32_t)st.exit_reason << 6) | ((uint32_t)st.tr_kind << 9) |
                  ((uint32_t)st.obj_c << 11) | ((uint32_t)st.mi << 13) | ((uint32_t)st.first_walk << 14) | ((uint32_t)st.have_closest << 15) |
                  ((uint32_t)st.sphere_bounced << 16) | ((uint32_t)st.sphere_free << 17) | (st.pass << 18);
    pv.u(31, s) = (uint32_t)(st.rd & 0xffff) | ((uint32_t)(st.vd & 0xffff) << 16);
    pv.u(32, s) = (uint32_t)(st.light_budget & 0xff) | ((uint32_t)st.light_index << 8);
    pv.u(33, s) = st.qslot;
}

VPT_DEV void ray_rng_init(PathState& st, const FrameArgs& fa, uint32_t k)
{
    uint32_t idx = st.lp;                                    // the ray's pixel word (vpt_frame.cuh): the pixel index itself for one rank
    if (fa.geom.n_ranks > 1) idx = ray_global_pixel(fa.geom, st.lp);
    st.rng.init(idx, fa.kp.iteration + st.pass, k);
}

template <class PV>
VPT_DEV void load_ray(const PV& pv, int s, PathState& st, const FrameArgs& fa)
{
    st.pos = f3(pv.f(0, s), pv.f(1, s), pv.f(2, s));    st.dir = f3(pv.f(3, s), pv.f(4, s), pv.f(5, s));
    st.L = f3(pv.f(6, s), pv.f(7, s), pv.f(8, s));      st.beta = f3(pv.f(9, s), pv.f(10, s), pv.f(11, s));
    st.wpos = f3(pv.f(12, s), pv.f(13, s), pv.f(14, s)); st.wdir = f3(pv.f(15, s), pv.f(16, s), pv.f(17, s));
    st.aux = f3(pv.f(18, s), pv.f(19, s), pv.f(20, s));
    st.alpha = pv.f(21, s); st.depth = pv.f(22, s); st.t = pv.f(23, s); st.distance = pv.f(24, s);
    st.trv = pv.f(25, s);   st.T_c = pv.f(26, s);   st.tmin_c = pv.f(27, s);
    const uint32_t k = pv.u(28, s); st.lp = pv.u(29, s);
    const uint32_t a = pv.u(30, s), b = pv.u(31, s), c = pv.u(32, s);
    st.phase = a & 15; st.mode = (a >> 4) & 3; st.exit_reason = (a >> 6) & 7; st.tr_kind = (a >> 9) & 3; st.obj_c = (a >> 11) & 3;
    st.mi = (a >> 13) & 1; st.first_walk = (a >> 14) & 1; st.have_closest = (a >> 15) & 1; st.sphere_bounced = (a >> 16) & 1; st.sphere_free = (a >> 17) & 1; st.pass = a >> 18;
    st.rd = b & 0xffff; st.vd = b >> 16;
    st.light_budget = (int)(int8_t)(c & 0xff); st.light_index = c >> 8;
    st.qslot = pv.u(33, s);
    st.org = fa.thin_lens ? f3(__ldg(fa.queue_aux + st.qslot).x, __ldg(fa.queue_aux + st.qslot).y, __ldg(fa.queue_aux + st.qslot).z) : ld3(fa.cam.origin);
    ray_rng_init(st, fa, k);
}

// A walk only touches part of the record (13 words in, 10 out instead of 34 each way) and only two of the eleven fields packed
// in word 30: the packed word is carried as is and patched on the way out (unpacking/packing everything cost 5 % of the kernel's
// instructions).  beta is not carried either: the one event that changes it (a scatter) updates it in place in shared memory.
constexpr uint32_t kExitMask = 7u << 6;
template <class PV>
VPT_DEV void load_walk(const PV& pv, int s, PathState& st, const FrameArgs& fa)
{
    st.wpos = f3(pv.f(12, s), pv.f(13, s), pv.f(14, s)); st.wdir = f3(pv.f(15, s), pv.f(16, s), pv.f(17, s));
    st.aux = f3(pv.f(18, s), pv.f(19, s), pv.f(20, s));
    st.alpha = pv.f(21, s); st.t = pv.f(23, s); st.distance = pv.f(24, s); st.trv = pv.f(25, s);
    const uint32_t k = pv.u(28, s); st.lp = pv.u(29, s);
    const uint32_t a = pv.u(30, s);
    st.packed = a;
    st.mode = (a >> 4) & 3; st.exit_reason = (a >> 6) & 7; st.sphere_free = (a >> 17) & 1; st.pass = a >> 18;
    ray_rng_init(st, fa, k);
}

template <class PV>
VPT_DEV void store_walk(const PV& pv, int s, const PathState& st)
{
    pv.f(12, s) = st.wpos.x; pv.f(13, s) = st.wpos.y; pv.f(14, s) = st.wpos.z;
    pv.f(18, s) = st.aux.x; pv.f(19, s) = st.aux.y; pv.f(20, s) = st.aux.z;
    pv.f(21, s) = st.alpha; pv.f(23, s) = st.t; pv.f(25, s) = st.trv;
    pv.u(28, s) = st.rng.k;
    pv.u(30, s) = (st.packed & ~kExitMask) | ((uint32_t)st.exit_reason << 6);
}

// Density of volume 0 from its cell table (vpt_cells_create): the eight corner texels of the cell under the point are 32 contiguous,
// aligned bytes -- one DRAM sector per look-up instead of the four or five a 2x2x2 footprint costs in the tiled texture array --
// blended with the texture unit's own arithmetic (vpt_texfilter.cuh).  Measured on the 1024^3 grid: DRAM traffic falls to the algorithmic
// 32 B per look-up, the time does not change: an experiment kept for that evidence, not a faster path.
VPT_DEV float cell_density(const FrameArgs& fa, const VolumeRec& v, float3 p)
{
    float3 uvw;
    if (!volume_coord(v, p, uvw)) return .0f;
    const TexCell q = tex_cell<0>(uvw, fa.cell_nx, fa.cell_ny, fa.cell_nz);
    const float4* c = fa.cell_table + 2 * (((size_t)q.k * (size_t)fa.cell_ny + (size_t)q.j) * (size_t)fa.cell_nx + (size_t)q.i);
    const float4 lo = __ldg(c), hi = __ldg(c + 1);
    return tex_blend<0>(q, lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w);
}

// ---- OP_STEP -------------------------------------------------------------------------------------------------
// kLean: single volume, no emission walk, no point lights -- the headline configuration; those features' code is compiled out
// of that instantiation (smaller hot loop: the kernel is fetch-stall bound)
// the per-step sphere test of a delta walk is rare (only lines that may hit the sphere): out of line, it keeps the hot loop shorter (6.88 -> 6.82 ms)
__device__ __noinline__ bool sphere_intersect_cold(const SphereRec& s, float3 ray_pos, float3 ray_dir, float& t_min, float& t_max) { return sphere_intersect(s, ray_pos, ray_dir, t_min, t_max); }
template <bool kLean, bool kCells = false>
VPT_DEV void walk_step(PathState& st, const FrameShared& fs, const FrameArgs& fa, const TraceConsts& tc, const SphereRec& sph, uint32_t& nlook,
                       float* beta = nullptr, int beta_stride = 0)   // kLean: the path's throughput in the parked record (x, y, z one stride apart)
{
    const SceneTables& sc = fs.sc;
    const vpt_kernel_params& kp = fa.kp;
    // ONE point location per call: an empty node is hopped over (no draw consumed) and the call returns.  A second and a third in-line copy
    // were measured: 7.17 ms and 8.04 ms of trace time on the headline frame against 6.88 ms -- the kernel is bound by instruction fetch,
    // every copy of the octree descent in the hot loop costs more than the iterations it saves.
    int leaf = oct_locate_or_skip(fs.oct, sc, st.wpos, st.wdir);
#if defined(VPT_WALK_HOPS) && VPT_WALK_HOPS >= 2
    if (leaf == -2) leaf = oct_locate_or_skip(fs.oct, sc, st.wpos, st.wdir);
#endif
    if (leaf == -2) return;
    if (leaf == -1) { st.op = OP_GLUE; st.exit_reason = EX_OUTSIDE; return; }

    if (st.mode == W_DELTA) {
        // distance to the box exit (or to the sphere) from the CURRENT position, every step (:1647-1651)
        float t_min, t_max, geo_dist = .0f;
        aabb_intersect(sc.root_pmin, sc.root_pmax, st.wpos, st.wdir, t_min, st.distance);
        if (!st.sphere_free && sphere_intersect_cold(sph, st.wpos, st.wdir, geo_dist, t_max)) st.distance = geo_dist;
    }
    const float u = st.rng.next();
    // t -= log(1-u) * a * b, as the reference build evaluates it: fma(b, a * (lg2(1-u) * -ln2), t)
    const float l2 = __log2f(psub(1.0f, u));
    if (st.mode == W_DELTA)      st.t = pfma(tc.inv_mult, pmul(tc.inv_max, pmul(l2, -0.693147182f)), st.t);
    else if (kLean || st.mode == W_RATIO) st.t = pfma(kp.tr_depth, pmul(tc.sigma_r_inv, pmul(l2, -0.693147182f)), st.t);
    else st.t = pfma(-pmul(pmul(pmul(l2, 0.693147182f), tc.inv_max), kp.tr_depth), 1.0f / kp.extinction.x, st.t);   // fma(-(log*a*b), rcp(ext), t): reference SASS
    if ((kLean || st.mode != W_EMIT) && st.t >= st.distance) { st.op = OP_GLUE; st.exit_reason = EX_DISTANCE; return; }

    st.wpos = madd3(st.wpos, st.wdir, st.t);                  // cumulative t, never reset (quirk Q2)
    if (!aabb_contains(sc.root_pmin, sc.root_pmax, st.wpos)) { st.op = OP_GLUE; st.exit_reason = EX_OUTSIDE; return; }

    nlook++;
    if (!kLean && st.mode == W_EMIT) {
        st.aux += leaf_emission(sc, fs.vol0, leaf, st.wpos, reinterpret_cast<const float3*>(kp.emission_texture), kp.emission_pivot, kp.emission_scale);
        return;
    }
    const float density = kLean ? 0.0f + (kCells ? cell_density(fa, fs.vol0, st.wpos) : volume_density(fs.vol0, st.wpos)) : leaf_density(sc, fs.vol0, leaf, st.wpos);
    if (st.mode == W_DELTA) {
        if (st.alpha < 1.0f) st.alpha += density;
        if (pmul(tc.inv_max, density) > st.rng.next()) {
            if (kLean) {
                // single volume: the colour terms are two cached loads; updating the parked throughput here measured faster than a
                // deferred update (trace 7.1 ms against 7.8 ms on the headline frame)
                const float3 Cd = fmax3(f3(0.0f), volume_color(fs.vol0, st.wpos));
                const int index = int(floorf(fminf(fmaxf((density * tc.inv_max * 255.0f / kp.emission_pivot), 0.0f), 255.0f)));
                const float3 density_color = reinterpret_cast<const float3*>(kp.density_color_texture)[index];
                float3 b3 = f3(beta[0], beta[beta_stride], beta[2 * beta_stride]);
                b3 *= (ld3(kp.albedo) * Cd * density_color / ld3(kp.extinction)) * float(kp.energy_inject);
                beta[0] = b3.x; beta[beta_stride] = b3.y; beta[2 * beta_stride] = b3.z;
            } else {
                // instanced / emissive scenes: the per-leaf colour look-ups are deferred to scatter_event() in the service round, where
                // many rays run them at once instead of one or two lanes of a stepping warp (fireball: trace 212 -> 185 ms)
                st.trv = density;                                   // free during a delta walk (the ratio walk re-initialises it)
                st.aux.x = __int_as_float(leaf);                    // aux is free during a delta walk, too
            }
            st.op = OP_GLUE; st.exit_reason = EX_SCATTER;
        }
    } else {
        st.trv = pmul(st.trv, pfma(-tc.sigma_r_inv, psub(density, tc.sigma_c), 1.0f));   // 1 - (rho - sigma_c) * sigma_r_inv, fused as in the reference SASS
        if (length(f3(st.trv)) < VPT_EPS) { st.op = OP_GLUE; st.exit_reason = EX_TR_DONE; }
    }
}

// The throughput factor of an accepted collision: albedo * Cd * density_colour / extinction * energy_inject (reference `sample`,
// render_kernel.cu:1664-1675).  Cd is the leaf's colour at the collision point -- the leaf located at the step's START, as the reference
// does -- and the colour LUT is indexed by density / sigma_max; the reference evaluates both at every step and drops them unless accepted.
VPT_DEV void scatter_event(PathState& st, const FrameShared& fs, const FrameArgs& fa, const TraceConsts& tc)
{
    const vpt_kernel_params& kp = fa.kp;
    const float density = st.trv;
    const float3 Cd = leaf_color(fs.sc, fs.vol0, __float_as_int(st.aux.x), st.wpos);
    const int index = int(floorf(fminf(fmaxf((density * tc.inv_max * 255.0f / kp.emission_pivot), 0.0f), 255.0f)));
    const float3 density_color = reinterpret_cast<const float3*>(kp.density_color_texture)[index];
    st.beta *= (ld3(kp.albedo) * Cd * density_color / ld3(kp.extinction)) * float(kp.energy_inject);
}

// ---- OP_TRBEGIN: reference Tr prologue (:1150-1167); the walk starts from (st.wpos, st.wdir) -----------------
VPT_DEV void begin_ratio_walk(PathState& st, const FrameShared& fs, const TraceConsts& tc, const SphereRec& sph)
{
    const SceneTables& sc = fs.sc;
    float3 p = st.wpos; const float3 d = st.wdir;
    float t_min, t_max, geo_dist = .0f, distance = .0f;
    st.op = OP_GLUE;                                            // unless a walk is really needed
    st.T_c = 1.0f;
    if (!aabb_contains(sc.root_pmin, sc.root_pmax, p)) {
        if (aabb_intersect(sc.root_pmin, sc.root_pmax, p, d, t_min, t_max)) p = madd3(p, d, padd(t_min, VPT_EPS));
        else { st.trv = 1.0f; return; }                         // misses the volume box: transmittance 1
    }
    aabb_intersect(sc.root_pmin, sc.root_pmax, p, d, t_min, distance);
    st.sphere_free = line_misses_sphere(sph, p, d);
    if (!st.sphere_free && sphere_intersect(sph, p, d, geo_dist, t_max)) { st.trv = 0.0f; return; }   // sphere occludes: BLACK
    st.T_c = expf(-tc.sigma_c * distance);
    st.wpos = p; st.t = 0.0f; st.distance = distance; st.trv = 1.0f;
    st.mode = W_RATIO; st.op = OP_STEP;
}

VPT_DEV float finish_ratio_walk(const PathState& st) { return clampf(st.trv * st.T_c, .0f, 1.0f); }

// ---- the integrator's control flow between heavy operations -------------------------------------------------
template <bool kLean>
VPT_DEV void advance(PathState& st, const FrameShared& fs, const FrameArgs& fa, const TraceConsts& tc, const SphereRec& sph)
{
    const SceneTables& sc = fs.sc;
    const vpt_kernel_params& kp = fa.kp;

    for (;;) {
        switch (st.phase) {
        case PH_BOUNCE_TOP:
            if (st.rd > kp.ray_depth) { st.op = OP_FINISH; return; }
            if (!st.have_closest) { st.op = OP_CLOSEST; st.phase = PH_TOP_HAVE; return; }
            st.phase = PH_TOP_HAVE;
            break;
        case PH_TOP_HAVE:
            if (st.first_walk && st.obj_c != 1) {                   // depth pass without a volume walk (:1883-1888)
                if (st.obj_c == 2) st.depth = length(st.org - (st.pos + st.dir * st.tmin_c));
                st.first_walk = false;
            }
            if (st.obj_c == 0) { st.op = OP_FINISH; return; }       // nothing ahead: every later bounce is a no-op
            if (st.obj_c == 1) {
                st.pos = madd3(st.pos, st.dir, padd(st.tmin_c, VPT_EPS));
                st.have_closest = false;
                st.vd = 1;
                st.phase = PH_VOL_ITER;
            } else st.phase = PH_SPHERE;                            // (a) and (b) see the same ray
            break;
        case PH_VOL_ITER:
            if (st.vd > kp.volume_depth) { st.phase = PH_VOL_DONE; break; }
            st.mi = false;
            st.wpos = st.pos; st.wdir = st.dir;
            if (!aabb_contains(sc.root_pmin, sc.root_pmax, st.pos)) {
                // `sample` leaves at once, and so does every remaining volume_depth iteration (ray unchanged, no draw)
                st.exit_reason = EX_OUTSIDE; st.vd = kp.volume_depth; st.phase = PH_AFTER_DELTA;
                break;
            }
            st.sphere_free = line_misses_sphere(sph, st.pos, st.dir);
            st.t = 0.0f; st.distance = .0f; st.mode = W_DELTA; st.op = OP_STEP; st.phase = PH_AFTER_DELTA;
            return;
        case PH_AFTER_DELTA: {
            st.pos = st.wpos;                                       // `sample` advances the caller's ray_pos
            int obj = 1;
            if (st.exit_reason == EX_SCATTER) { if (!kLean) scatter_event(st, fs, fa, tc); st.mi = true; }
            if (st.exit_reason == EX_DISTANCE) obj = 2;             // compiled reference: obj = 2 on every distance exit (Q4)
            if (st.first_walk) {
                st.depth = st.mi ? length(st.org - st.pos) : .0f;
                // the reference runs this identical walk twice (depth pass + integrator) and accumulates `tr` in
                // both; the replay adds the same densities again while tr < 1
                if (st.alpha < 1.0f) st.alpha += st.alpha;
                st.first_walk = false;
            }
            if (is_black(st.beta) || obj == 2) { st.phase = PH_VOL_DONE; break; }
            if (st.mi) hg_sample(st.dir, st.rng, kp.phase_g1);
            st.vd++; st.phase = PH_VOL_ITER;
            break;
        }
        case PH_AFTER_HG:
            st.vd++; st.phase = PH_VOL_ITER;
            break;
        case PH_VOL_DONE:
            if (st.mi) {
                st.tr_kind = TR_SUN; st.wpos = st.pos; st.wdir = tc.sun_dir;
                st.op = OP_TRBEGIN; st.phase = PH_AFTER_TR;
                return;
            }
            st.phase = PH_EMISSION;
            break;
        case PH_AFTER_TR: {
            const float tr = finish_ratio_walk(st);
            if (st.tr_kind == TR_SUN) {                             // reference estimate_sun (:1478-1516)
                const float cos_theta = dot(st.dir, tc.sun_dir);
                const float phase_pdf = hg_phase(cos_theta, kp.phase_g1);
                const float3 Ld = f3(tr) * phase_pdf;
                st.L += Ld * ld3(kp.sun_color) * kp.sun_mult * st.beta;
                if (!kLean && fa.lights.num_lights > 0) { st.aux = f3(.0f); st.light_budget = 10; st.phase = PH_POINT_NEXT; }
                else st.phase = PH_EMISSION;
            } else if (!kLean && st.tr_kind == TR_POINT) {                    // reference point_light::Le, light.h:104-121
                if (st.light_budget < (int)fa.lights.num_lights) {
                    const vpt_point_light& pl = reinterpret_cast<const vpt_point_light*>(fa.lights.light_ptr)[st.light_index];
                    const float3 lpos = ld3(pl.pos);
                    const float3 wi = normalize(lpos - st.pos);
                    const float cos_theta = dot(st.dir, wi);
                    const float phase_pdf = hg_phase(cos_theta, kp.phase_g1);
                    const float sqr_dist = length(lpos * lpos - st.pos * st.pos);
                    const float falloff = 1 / sqr_dist;
                    st.aux += ld3(pl.color) * pl.power * f3(tr) * phase_pdf * falloff;
                }
                st.light_budget--;
                st.phase = PH_POINT_NEXT;
            } else {                                                // sphere branch tail (:1831-1833)
                st.L += ld3(kp.sun_color) * kp.sun_mult * f3(tr) * fmaxf(dot(tc.sun_dir, st.aux), .0f) * st.beta;
                if (fa.planeD) fa.planeD[(size_t)st.pass * fa.geom.n_local + ray_local_pixel(fa.geom, st.lp)] = make_float4(st.pos.x, st.pos.y, st.pos.z, 0.f);   // env_pos = ray_pos
                st.sphere_bounced = true;
                st.rd++; st.have_closest = false;
                st.phase = PH_BOUNCE_TOP;
            }
            break;
        }
        case PH_POINT_NEXT: {                                       // reference estimate_point_light (:1445-1475)
            if (kLean) { st.phase = PH_EMISSION; break; }
            if (st.light_budget < 0) { st.L += st.aux * st.beta; st.phase = PH_EMISSION; break; }
            const vpt_point_light* lp = reinterpret_cast<const vpt_point_light*>(fa.lights.light_ptr);
            st.light_index = int(floorf(st.rng.next() * fa.lights.num_lights));
            st.tr_kind = TR_POINT; st.wpos = st.pos; st.wdir = normalize(ld3(lp[st.light_index].pos) - st.pos);
            st.op = OP_TRBEGIN; st.phase = PH_AFTER_TR;
            return;
        }
        case PH_EMISSION:
            if (!kLean && kp.emission_scale > 0 && st.mi) {
                st.wpos = st.pos; st.wdir = st.dir; st.t = 0.0f; st.aux = f3(.0f);
                st.mode = W_EMIT; st.op = OP_STEP; st.phase = PH_AFTER_EMIT;
                return;
            }
            st.phase = PH_AFTER_VOLUME;
            break;
        case PH_AFTER_EMIT:
            st.L += st.aux;
            st.phase = PH_AFTER_VOLUME;
            break;
        case PH_AFTER_VOLUME: {                                     // the ray moved since the last test
            // Later bounces cannot change the sample unless the sphere comes into play (quirk Q21: a ray inside the box is
            // teleported to its exit, a ray outside it and leaving sees nothing).  If this ray's line provably misses the
            // sphere and the position is not exactly on the box boundary, retire the path here: bit-identical output.
            const bool inside_strict = st.pos.x > sc.root_pmin[0] && st.pos.x < sc.root_pmax[0] && st.pos.y > sc.root_pmin[1] &&
                                       st.pos.y < sc.root_pmax[1] && st.pos.z > sc.root_pmin[2] && st.pos.z < sc.root_pmax[2];
            const bool outside = !aabb_contains(sc.root_pmin, sc.root_pmax, st.pos);
            // outside here means the walk stepped out of the box along dir (or never re-entered it): the box lies behind
            if ((inside_strict || (outside && st.exit_reason == EX_OUTSIDE && !st.sphere_bounced)) && line_misses_sphere(sph, st.pos, st.dir)) { st.op = OP_FINISH; return; }
            st.op = OP_CLOSEST; st.phase = PH_AFTERVOL_HAVE;
            return;
        }
        case PH_AFTERVOL_HAVE:
            if (st.obj_c == 2) { st.phase = PH_SPHERE; break; }
            st.rd++; st.phase = PH_BOUNCE_TOP;                      // next bounce starts from the same ray: reuse the test
            break;
        case PH_SPHERE: {                                           // bounce off the reference sphere (:1807-1834)
            st.pos += st.dir * st.tmin_c;
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
            st.aux = normal;
            st.tr_kind = TR_SPHERE; st.wpos = st.pos; st.wdir = tc.sun_dir;
            st.op = OP_TRBEGIN; st.phase = PH_AFTER_TR;
            return;
        }
        default:
            st.op = OP_FINISH;
            return;
        }
    }
}

#include "vpt_trace_vol.cuh"

template <int kInteg>
VPT_DEV void write_sample(const PathState& st, const FrameArgs& fa)
{
    uint32_t lp = st.lp;
    if (fa.geom.n_ranks > 1) lp = ray_local_pixel(fa.geom, st.lp);
    const size_t o = (size_t)st.pass * fa.geom.n_local + lp;
    fa.planeA[o] = make_float4(st.dir.x, st.dir.y, st.dir.z, st.alpha);                 // final direction, tr
    fa.planeB[o] = make_float4(st.L.x, st.L.y, st.L.z, st.depth);                       // L, depth
    fa.planeC[o] = make_float4(st.beta.x, st.beta.y, st.beta.z, 1.f);                   // beta
    if (kInteg) {
        // vol_integrator evaluates the sky from env_pos while the throughput is still ~white, else from where the path ended (:1750)
        float3 e = length(st.beta) > 0.9999f ? st.org : st.pos;
        if (fa.debug_flags & 16) e = st.org;                                                // development probes
        if (fa.debug_flags & 32) e = st.pos;
        fa.planeD[o] = make_float4(e.x, e.y, e.z, 0.f);
    } else if (fa.planeD && !st.sphere_bounced) fa.planeD[o] = make_float4(st.org.x, st.org.y, st.org.z, 0.f);   // env_pos = camera origin unless the sphere branch moved it
}

// kInteg = 0: direct integrator (the tuned headline path); kInteg = 1: volumetric path integrator, which additionally needs the
// caller's AtmosphereParameters in the kernel (sky radiance decides whether a transmittance walk is run at all)
struct NoAtmo { int pad[4]; };

template <int kInteg, bool kLean, int kSlots, bool kCells = false>
__global__ void __launch_bounds__(kTraceThreads, trace_min_ctas(kSlots))
k_trace(const FrameArgs fa, const typename std::conditional<kInteg != 0, vpt_atmosphere, NoAtmo>::type atm)
{
    __shared__ FrameShared fs;
    extern __shared__ float pool_smem[];                         // kTraceWarps x kRayWords x kPool words (opt-in dynamic shared memory)
    load_frame_shared(fs, fa.scene);
    const SceneTables& sc = fs.sc;
    const vpt_kernel_params& kp = fa.kp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    TraceConsts tc;
    tc.inv_max = 1.0f / sc.max_extinction;
    tc.inv_mult = 1.0f / kp.density_mult;
    tc.sigma_c = sc.min_extinction;
    tc.sigma_r_inv = 1.0f / (sc.max_extinction - tc.sigma_c);
    tc.sun_dir = sun_direction(kp.azimuth, kp.elevation);
    const SphereRec sph = load_sphere(fa.sphere);

    constexpr int kPool = 32 * kSlots;
    PoolView<kPool> pv; pv.w = pool_smem + (size_t)warp * kRayWords * kPool + lane;
    constexpr int OP_NOSLOT = 15;                                 // third slot disabled when kSlots == 2
    int tag0 = OP_IDLE, tag1 = OP_IDLE, tag2 = (kSlots > 2) ? OP_IDLE : OP_NOSLOT;   // operation each of my rays waits for

    const unsigned q_count = *fa.queue_count;
    bool queue_dry = (q_count == 0);
    uint32_t nlook = 0, lane_steps = 0, warp_iters = 0, lane_ops = 0, warp_ops = 0, lane_rays = 0;   // statistics
    PathState st;

    for (;;) {
        // ---- refill free slots from the global queue: one atomic per warp, one record per lane that has room ----------
        const bool has_idle = (tag0 == OP_IDLE) | (tag1 == OP_IDLE) | (tag2 == OP_IDLE);
        const unsigned idle_lanes = __ballot_sync(0xffffffffu, has_idle);
        if (!queue_dry && __popc(idle_lanes) >= fa.sched_min_lanes) {
            unsigned base = 0;
            const int leader = __ffs(idle_lanes) - 1;
            if (lane == leader) base = atomicAdd(fa.queue_head, (unsigned)__popc(idle_lanes));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + __popc(idle_lanes) >= q_count) queue_dry = true;
            const unsigned slot = base + __popc(idle_lanes & lt_mask);
            if (has_idle && slot < q_count) {
                const int j = (tag0 == OP_IDLE) ? 0 : (tag1 == OP_IDLE) ? 1 : 2;
                const float4 r0 = __ldg(fa.queue_dir + slot);
                const uint2 id = __ldg(fa.queue_id + slot);
                st.dir = f3(r0.x, r0.y, r0.z);
                st.qslot = slot;
                st.org = fa.thin_lens ? f3(__ldg(fa.queue_aux + slot).x, __ldg(fa.queue_aux + slot).y, __ldg(fa.queue_aux + slot).z) : ld3(fa.cam.origin);
                st.lp = id.x; st.pass = id.y & 63u;
                st.rng.k = (id.y >> 6) & 1023u;
                st.pos = st.org;
                st.beta = f3(1.0f); st.L = f3(.0f); st.alpha = .0f; st.depth = .0f;
                st.wpos = st.pos; st.wdir = st.dir; st.aux = f3(0.f); st.t = 0.f; st.distance = 0.f; st.trv = 1.f; st.T_c = 1.f;
                st.mi = false; st.first_walk = true; st.sphere_bounced = false; st.rd = 1; st.vd = 1; st.light_budget = 0; st.light_index = 0;
                st.mode = W_DELTA; st.exit_reason = EX_NONE; st.tr_kind = TR_SUN;
                st.tmin_c = r0.w; st.obj_c = (int)((id.y >> 16) & 3u); st.have_closest = true;       // k_generate already ran the first test
                st.phase = kInteg ? (int)VP_START : (int)PH_BOUNCE_TOP; st.sphere_free = false; st.op = OP_GLUE;
                if ((id.y >> 18) & 1u) {
                    // k_generate already hopped over the empty nodes in front of this ray: it is a delta walker at `wpos`
                    const float4 ws = __ldg(fa.queue_aux + slot);
                    st.wpos = f3(ws.x, ws.y, ws.z); st.pos = st.wpos;
                    st.have_closest = false; st.sphere_free = true;
                    st.phase = kInteg ? (int)VP_AFTER_DELTA : (int)PH_AFTER_DELTA; st.op = OP_STEP;
                }
                store_ray(pv, j, st);
                if (j == 0) tag0 = st.op; else if (j == 1) tag1 = st.op; else tag2 = st.op;
                lane_rays++;
            }
        }

        // ---- vote: step, or service the rays that wait for bookkeeping / a closest-object test / a transmittance set-up? ----
        const bool has_step = (tag0 == OP_STEP) | (tag1 == OP_STEP) | (tag2 == OP_STEP);
        const bool svc0 = (tag0 != OP_STEP) & (tag0 != OP_IDLE), svc1 = (tag1 != OP_STEP) & (tag1 != OP_IDLE), svc2 = (tag2 != OP_STEP) & (tag2 != OP_IDLE) & (tag2 != OP_NOSLOT);
        const int nS = __popc(__ballot_sync(0xffffffffu, has_step));
        const int nV = __popc(__ballot_sync(0xffffffffu, svc0 | svc1 | svc2));
        if ((nS | nV) == 0) {
            if (queue_dry) break;                                  // nothing parked, nothing left to fetch
            continue;                                              // everything idle: the refill above takes records next round
        }

        if (nV > nS) {
            // ---- service round: one parked ray per participating lane runs [closest-object test] -> integrator control flow
            //      (incl. HG resampling, NEE bookkeeping, retirement) -> [transmittance set-up]; each block has one code site ----
            const int j = svc0 ? 0 : svc1 ? 1 : svc2 ? 2 : -1;
            if (j >= 0) {
                load_ray(pv, j, st, fa);
                st.op = (j == 0) ? tag0 : (j == 1) ? tag1 : tag2;
                if (st.op == OP_CLOSEST) {
                    st.obj_c = closest_object(sc, sph, st.pos, st.dir, st.tmin_c);
                    st.have_closest = true; st.op = OP_GLUE;
                }
                if (st.op == OP_GLUE) {
                    if constexpr (kInteg != 0) advance_vol(st, fs, fa, atm, tc, sph);
                    else advance<kLean>(st, fs, fa, tc, sph);
                }
                if (st.op == OP_TRBEGIN) begin_ratio_walk(st, fs, tc, sph);
                if (st.op == OP_FINISH) { write_sample<kInteg>(st, fa); st.op = OP_IDLE; }
                else store_ray(pv, j, st);
                if (j == 0) tag0 = st.op; else if (j == 1) tag1 = st.op; else tag2 = st.op;
                lane_ops++;
            }
            warp_ops++;
            continue;
        }

        // ---- stepping loop: a lane steps one of its walkers; when that walk ends it parks the ray and swaps in its next walker ----
        {
            int cur = -1;
            int parked_lanes = nV;                                // lanes with something to service (grows as walks end)
            for (;;) {
                if (cur < 0) {
                    cur = (tag0 == OP_STEP) ? 0 : (tag1 == OP_STEP) ? 1 : (tag2 == OP_STEP) ? 2 : -1;
                    if (cur >= 0) { load_walk(pv, cur, st, fa); st.op = OP_STEP; }
                }
                const unsigned walking = __ballot_sync(0xffffffffu, cur >= 0);
                if (walking == 0u) break;
                // few walkers left and a fuller group is waiting: park and let the vote pick it
                if (__popc(walking) < fa.sched_min_lanes && parked_lanes > __popc(walking)) {
                    if (cur >= 0) { store_walk(pv, cur, st); cur = -1; }
                    break;
                }
                if (cur >= 0) {
                    walk_step<kLean, kCells>(st, fs, fa, tc, sph, nlook, &pv.f(9, cur), kPool); lane_steps++;
                    if (st.op != OP_STEP) {                        // walk ended: park the ray with its new tag
                        store_walk(pv, cur, st);
                        if (cur == 0) tag0 = st.op; else if (cur == 1) tag1 = st.op; else tag2 = st.op;
                        cur = -1;
                    }
                }
                parked_lanes = __popc(__ballot_sync(0xffffffffu, ((tag0 != OP_STEP) & (tag0 != OP_IDLE)) | ((tag1 != OP_STEP) & (tag1 != OP_IDLE)) | ((tag2 != OP_STEP) & (tag2 != OP_IDLE) & (tag2 != OP_NOSLOT))));
                warp_iters++;
            }
        }
    }

    if (fa.counters) {                                        // optional statistics (one atomic set per warp)
        unsigned long long a = nlook, b = lane_steps, c = lane_ops, d = lane_rays;
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o); d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if (lane == 0) {
            atomicAdd(fa.counters + 0, a); atomicAdd(fa.counters + 1, b); atomicAdd(fa.counters + 2, (unsigned long long)warp_iters);
            atomicAdd(fa.counters + 3, c); atomicAdd(fa.counters + 4, (unsigned long long)warp_ops); atomicAdd(fa.counters + 5, d);
        }
    }
}
