// vpt_trace.cuh -- k_trace: persistent wavefront over the hit queue, scheduled per warp by operation.
//
// Every lane owns one ray and carries an `op` = the next heavy operation its path needs:
//     OP_STEP     one unified tracking step (delta / residual-ratio / emission walk)
//     OP_CLOSEST  nearest of {octree root box, sphere}            (reference get_closest_object)
//     OP_HG       Henyey-Greenstein direction resampling          (reference sample_hg)
//     OP_TRBEGIN  set-up of a residual-ratio transmittance walk   (reference Tr prologue)
//     OP_FINISH   write the sample record
//     OP_GLUE     cheap bookkeeping between the above (the integrator's control flow)
// Each operation has exactly ONE code site.  Every round the warp votes (ballot + popc) and executes
// the operation most lanes are waiting for, so divergent estimator code runs with a majority of the
// lanes active instead of one lane at a time, and the kernel's instruction footprint stays small enough
// for the instruction cache (the first version inlined the transitions: 15 k SASS instructions, 37 %
// of stall samples were instruction-fetch misses, 10 of 32 lanes active on average -- profiles/).
// Finished lanes refill from the ray queue with one atomic per warp.
//
// Control flow restated from the reference direct integrator (render_kernel.cu:1760-1857) with three
// bit-exact shortcuts: (1) the depth pass replays the integrator's first walk on a copy of the RNG
// (:1859-1889), so it is taken from that walk instead of being run again; (2) the closest-object test
// at the end of a bounce and the one at the top of the next have identical inputs unless the sphere
// branch ran, so the result is reused; (3) a bounce that finds nothing ahead makes every later bounce
// a no-op, so the path retires there.
#pragma once
// (included from vpt_kernels.cu inside `namespace vpt`)

enum WalkMode : int { W_DELTA = 1, W_RATIO = 2, W_EMIT = 3 };
enum Op : int { OP_IDLE = 0, OP_GLUE, OP_STEP, OP_CLOSEST, OP_HG, OP_TRBEGIN, OP_FINISH };
enum Phase : int {
    PH_BOUNCE_TOP = 0, PH_TOP_HAVE, PH_VOL_ITER, PH_AFTER_DELTA, PH_AFTER_HG, PH_VOL_DONE, PH_AFTER_TR,
    PH_POINT_NEXT, PH_EMISSION, PH_AFTER_EMIT, PH_AFTER_VOLUME, PH_AFTERVOL_HAVE, PH_SPHERE
};
enum ExitReason : int { EX_NONE = 0, EX_OUTSIDE, EX_DISTANCE, EX_SCATTER, EX_TR_DONE };
enum TrKind : int { TR_SUN = 0, TR_POINT = 1, TR_SPHERE = 2 };

struct PathState {
    float3 pos, dir;      // the integrator's ray
    float3 org;           // camera-ray origin (depth reference, default env_pos)
    float3 env_pos;
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
    bool   mi, first_walk, have_closest;
    uint32_t lp, pass;
    uint32_t nlook;
    Rng    rng;
};

struct TraceConsts {
    float inv_max, inv_mult, sigma_c, sigma_r_inv;
    float3 sun_dir;
};

// ---- OP_STEP ---------------------------------------------------------------------------------------------
VPT_DEV void walk_step(PathState& st, const FrameShared& fs, const FrameArgs& fa, const TraceConsts& tc, const SphereRec& sph)
{
    const SceneTables& sc = fs.sc;
    const vpt_kernel_params& kp = fa.kp;
    const int leaf = oct_locate_or_skip(fs.oct, sc, st.wpos, st.wdir);
    if (leaf == -2) return;                                   // skipped an empty node, no draw consumed
    if (leaf == -1) { st.op = OP_GLUE; st.exit_reason = EX_OUTSIDE; return; }

    if (st.mode == W_DELTA) {
        // distance to the box exit (or to the sphere) from the CURRENT position, every step (:1647-1651)
        float t_min, t_max, geo_dist = .0f;
        aabb_intersect(sc.root_pmin, sc.root_pmax, st.wpos, st.wdir, t_min, st.distance);
        if (sphere_intersect(sph, st.wpos, st.wdir, geo_dist, t_max)) st.distance = geo_dist;
    }
    const float u = st.rng.next();
    // t -= log(1-u) * a * b, as the reference build evaluates it: fma(b, a * (lg2(1-u) * -ln2), t)
    const float l2 = __log2f(psub(1.0f, u));
    if (st.mode == W_DELTA)      st.t = pfma(tc.inv_mult, pmul(tc.inv_max, pmul(l2, -0.693147182f)), st.t);
    else if (st.mode == W_RATIO) st.t = pfma(kp.tr_depth, pmul(tc.sigma_r_inv, pmul(l2, -0.693147182f)), st.t);
    else st.t = psub(st.t, __fdividef(pmul(kp.tr_depth, pmul(tc.inv_max, pmul(l2, 0.693147182f))), kp.extinction.x));
    if (st.mode != W_EMIT && st.t >= st.distance) { st.op = OP_GLUE; st.exit_reason = EX_DISTANCE; return; }

    st.wpos = madd3(st.wpos, st.wdir, st.t);                  // cumulative t, never reset (quirk Q2)
    if (!aabb_contains(sc.root_pmin, sc.root_pmax, st.wpos)) { st.op = OP_GLUE; st.exit_reason = EX_OUTSIDE; return; }

    st.nlook++;
    if (st.mode == W_EMIT) {
        st.aux += leaf_emission(sc, fs.vol0, leaf, st.wpos, reinterpret_cast<const float3*>(kp.emission_texture), kp.emission_pivot, kp.emission_scale);
        return;
    }
    const float density = leaf_density(sc, fs.vol0, leaf, st.wpos);
    if (st.mode == W_DELTA) {
        const float3 Cd = leaf_color(sc, fs.vol0, leaf, st.wpos);
        const int index = int(floorf(fminf(fmaxf((density * tc.inv_max * 255.0f / kp.emission_pivot), 0.0f), 255.0f)));
        const float3 density_color = reinterpret_cast<const float3*>(kp.density_color_texture)[index];
        if (st.alpha < 1.0f) st.alpha += density;
        if (pmul(tc.inv_max, density) > st.rng.next()) {
            st.beta *= (ld3(kp.albedo) * Cd * density_color / ld3(kp.extinction)) * float(kp.energy_inject);
            st.op = OP_GLUE; st.exit_reason = EX_SCATTER;
        }
    } else {
        st.trv = pmul(st.trv, psub(1.0f, pmul(tc.sigma_r_inv, psub(density, tc.sigma_c))));
        if (length(f3(st.trv)) < VPT_EPS) { st.op = OP_GLUE; st.exit_reason = EX_TR_DONE; }
    }
}

// ---- OP_TRBEGIN: reference Tr prologue (:1150-1167); the walk starts from (st.wpos, st.wdir) -------------
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
    if (sphere_intersect(sph, p, d, geo_dist, t_max)) { st.trv = 0.0f; return; }   // sphere occludes: BLACK
    st.T_c = expf(-tc.sigma_c * distance);
    st.wpos = p; st.t = 0.0f; st.distance = distance; st.trv = 1.0f;
    st.mode = W_RATIO; st.op = OP_STEP;
}

VPT_DEV float finish_ratio_walk(const PathState& st) { return clampf(st.trv * st.T_c, .0f, 1.0f); }

// ---- OP_GLUE: the integrator's control flow between heavy operations -------------------------------------
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
            st.phase = PH_AFTER_DELTA;
            if (!(fa.debug_flags & 2) && !aabb_contains(sc.root_pmin, sc.root_pmax, st.pos)) { st.exit_reason = EX_OUTSIDE; break; }   // walk would leave at once
            st.t = 0.0f; st.distance = .0f; st.mode = W_DELTA; st.op = OP_STEP;
            return;
        case PH_AFTER_DELTA: {
            st.pos = st.wpos;                                       // `sample` advances the caller's ray_pos
            int obj = 1;
            if (st.exit_reason == EX_SCATTER) st.mi = true;
            if (st.exit_reason == EX_DISTANCE) obj = 2;             // compiled reference: obj = 2 on every distance exit (Q4)
            if (st.first_walk) {
                st.depth = st.mi ? length(st.org - st.pos) : .0f;
                // the reference runs this identical walk twice (depth pass + integrator) and accumulates `tr` in
                // both; the replay adds the same densities again while tr < 1
                if (st.alpha < 1.0f) st.alpha += st.alpha;
                st.first_walk = false;
            }
            if (is_black(st.beta) || obj == 2) { st.phase = PH_VOL_DONE; break; }
            if (st.mi) { st.op = OP_HG; st.phase = PH_AFTER_HG; return; }
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
                if (fa.lights.num_lights > 0) { st.aux = f3(.0f); st.light_budget = 10; st.phase = PH_POINT_NEXT; }
                else st.phase = PH_EMISSION;
            } else if (st.tr_kind == TR_POINT) {                    // reference point_light::Le, light.h:104-121
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
                st.env_pos = st.pos;
                st.rd++; st.have_closest = false;
                st.phase = PH_BOUNCE_TOP;
            }
            break;
        }
        case PH_POINT_NEXT: {                                       // reference estimate_point_light (:1445-1475)
            if (st.light_budget < 0) { st.L += st.aux * st.beta; st.phase = PH_EMISSION; break; }
            const vpt_point_light* lp = reinterpret_cast<const vpt_point_light*>(fa.lights.light_ptr);
            st.light_index = int(floorf(st.rng.next() * fa.lights.num_lights));
            st.tr_kind = TR_POINT; st.wpos = st.pos; st.wdir = normalize(ld3(lp[st.light_index].pos) - st.pos);
            st.op = OP_TRBEGIN; st.phase = PH_AFTER_TR;
            return;
        }
        case PH_EMISSION:
            if (kp.emission_scale > 0 && st.mi) {
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
        case PH_AFTER_VOLUME:                                       // the ray moved since the last test
            st.op = OP_CLOSEST; st.phase = PH_AFTERVOL_HAVE;
            return;
        case PH_AFTERVOL_HAVE:
            if (st.obj_c == 2) { st.phase = PH_SPHERE; break; }
            st.rd++; st.phase = PH_BOUNCE_TOP;                      // next bounce starts from the same ray: reuse the test
            if (fa.debug_flags & 4) st.have_closest = false;
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

__global__ void __launch_bounds__(kTraceThreads, kTraceMinCtas)
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
    st.op = OP_IDLE; st.nlook = 0;
    bool queue_dry = (q_count == 0);
    uint32_t lane_steps = 0, warp_iters = 0, lane_ops = 0, warp_ops = 0;   // statistics

    for (;;) {
        // ---- retire finished paths (three coalescible stores) ----
        if (__ballot_sync(0xffffffffu, st.op == OP_FINISH)) {
            if (st.op == OP_FINISH) {
                const size_t s = (size_t)st.pass * g.n_local + st.lp;
                fa.planeA[s] = make_float4(st.dir.x, st.dir.y, st.dir.z, st.alpha);
                fa.planeB[s] = make_float4(st.L.x, st.L.y, st.L.z, st.depth);
                fa.planeC[s] = make_float4(st.beta.x, st.beta.y, st.beta.z, 1.f);
                if (fa.planeD) fa.planeD[s] = make_float4(st.env_pos.x, st.env_pos.y, st.env_pos.z, 0.f);
                st.op = OP_IDLE;
            }
        }
        // ---- refill idle lanes from the ray queue: one atomic per warp, amortised over >= 8 lanes ----
        const unsigned idle = __ballot_sync(0xffffffffu, st.op == OP_IDLE);
        if (idle == 0xffffffffu && queue_dry) break;
        if (idle && !queue_dry && (__popc(idle) >= kRefillLanes || idle == 0xffffffffu ||
                                   __ballot_sync(0xffffffffu, st.op == OP_STEP) == 0u)) {
            unsigned base = 0;
            const int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(fa.queue_head, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + __popc(idle) >= q_count) queue_dry = true;
            if (st.op == OP_IDLE) {
                const unsigned slot = base + __popc(idle & ((1u << lane) - 1u));
                if (slot < q_count) {
                    const float4 r0 = __ldg(fa.queue_dir + slot);
                    const uint2 id = __ldg(fa.queue_id + slot);
                    st.dir = f3(r0.x, r0.y, r0.z);
                    st.org = fa.queue_org ? f3(__ldg(fa.queue_org + slot).x, __ldg(fa.queue_org + slot).y, __ldg(fa.queue_org + slot).z) : ld3(fa.cam.origin);
                    st.lp = id.x;
                    st.pass = id.y & 63u;
                    uint32_t idx = st.lp;
                    if (g.n_ranks > 1) {
                        const uint32_t lr = st.lp / (uint32_t)g.width, x = st.lp - lr * (uint32_t)g.width;
                        idx = (uint32_t)global_row(g, (int)lr) * (uint32_t)g.width + x;
                    }
                    st.rng.init(idx, kp.iteration + st.pass, (id.y >> 6) & 1023u);
                    st.pos = st.org; st.env_pos = st.org;
                    st.beta = f3(1.0f); st.L = f3(.0f); st.alpha = .0f; st.depth = .0f;
                    st.mi = false; st.first_walk = true; st.rd = 1;
                    st.tmin_c = r0.w; st.obj_c = (int)(id.y >> 16); st.have_closest = !(fa.debug_flags & 1);   // k_generate already ran the first test
                    st.op = OP_GLUE; st.phase = PH_BOUNCE_TOP;
                }
            }
        }
        // ---- bookkeeping for lanes between heavy operations ----
        if (st.op == OP_GLUE) advance(st, fs, fa, tc, sph);

        // ---- vote: one operation per round.  Operations other than STEP run once per lane and unblock walkers, so
        //      the largest such group runs as soon as it has gathered `sched_min_lanes` lanes (or nobody is stepping);
        //      otherwise every walker takes one step.
        const unsigned mS = __ballot_sync(0xffffffffu, st.op == OP_STEP);
        const unsigned mC = __ballot_sync(0xffffffffu, st.op == OP_CLOSEST);
        const unsigned mH = __ballot_sync(0xffffffffu, st.op == OP_HG);
        const unsigned mT = __ballot_sync(0xffffffffu, st.op == OP_TRBEGIN);
        const int nS = __popc(mS), nC = __popc(mC), nH = __popc(mH), nT = __popc(mT);
        if ((nS | nC | nH | nT) == 0) continue;                 // only finishes / refills pending
        const int others = max(nC, max(nH, nT));
        if (nS > 0 && (others < fa.sched_min_lanes && !(others > 0 && nS < others))) {
            if (st.op == OP_STEP) { walk_step(st, fs, fa, tc, sph); lane_steps++; }
            warp_iters++;
        } else if (nC >= nH && nC >= nT) {
            if (st.op == OP_CLOSEST) {
                st.obj_c = closest_object(sc, sph, st.pos, st.dir, st.tmin_c);
                st.have_closest = true; st.op = OP_GLUE; lane_ops++;
            }
            warp_ops++;
        } else if (nT >= nH) {
            if (st.op == OP_TRBEGIN) { begin_ratio_walk(st, fs, tc, sph); lane_ops++; }
            warp_ops++;
        } else {
            if (st.op == OP_HG) { hg_sample(st.dir, st.rng, kp.phase_g1); st.op = OP_GLUE; lane_ops++; }
            warp_ops++;
        }
    }

    if (fa.counters) {                                        // optional statistics (one atomic set per warp)
        unsigned long long a = st.nlook, b = lane_steps, c = lane_ops;
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            atomicAdd(fa.counters + 0, a); atomicAdd(fa.counters + 1, b); atomicAdd(fa.counters + 2, (unsigned long long)warp_iters);
            atomicAdd(fa.counters + 3, c); atomicAdd(fa.counters + 4, (unsigned long long)warp_ops);
        }
    }
}

