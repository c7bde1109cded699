// vpt_context.cpp -- C-ABI implementation (include/vpt_b200.h): context, render-pass scheduling,
// scene-table cache, and the host-side scene helpers.
//
// Scheduling of one vpt_render_passes(n) call on the caller's stream (no host synchronisation):
//   [once per (volumes, octree) pair]  k_prepare_scene
//   for each chunk of <= passes_per_chunk sampled passes:
//        memset(queue counters) -> k_bn_prepare -> k_generate -> k_trace<integrator, lean> (persistent) -> k_resolve<env>
//   passes beyond max_interactions / render == false: k_resolve (re-tonemap) + k_bn_advance
// which leaves every buffer named by Kernel_params in the state n reference launches would.
#include "vpt_host.h"
#include "vdb_reader.h"
#include "image_io.h"

#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>
#include <string>
#include <algorithm>
#include <vector>

static thread_local std::string g_last_error;
namespace vpt { int fail_global(int code, const std::string& msg) { g_last_error = msg; return code; } }

static int fail(vpt_context* ctx, int code, const std::string& msg) {
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}
namespace vpt { int fail_ctx(vpt_context* ctx, int code, const std::string& msg) { return fail(ctx, code, msg); } }
#define VPT_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(ctx, VPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

namespace vpt { FrameGeom make_frame_geom(const vpt_context* c, unsigned w, unsigned h); }
static vpt::FrameGeom make_geom(const vpt_context* c, unsigned w, unsigned h) { return vpt::make_frame_geom(c, w, h); }
vpt::FrameGeom vpt::make_frame_geom(const vpt_context* c, unsigned w, unsigned h) {
    vpt::FrameGeom g;
    g.width = (int)w; g.height = (int)h; g.n_ranks = c->n_ranks; g.rank = c->rank;
    if (c->n_ranks == 1) { g.stripe_h = (int)h > 0 ? (int)h : 1; g.local_rows = (int)h; }
    else {
        g.stripe_h = c->stripe_rows;
        const int stripes = ((int)h + g.stripe_h - 1) / g.stripe_h;
        const int per_rank = (stripes + c->n_ranks - 1) / c->n_ranks;
        g.local_rows = per_rank * g.stripe_h;
    }
    g.n_local = g.local_rows * g.width;
    return g;
}

extern "C" {

const char* vpt_version(void) { return "vpt-b200 0.1 (sm_100a wavefront build)"; }

// sizeof of every boundary struct, in VPT_ARG order then the auxiliary ones (no CUDA call: usable on a CPU-only host)
int vpt_abi_sizes(size_t* out, int n) {
    const size_t s[] = { sizeof(vpt_camera), sizeof(vpt_light_list), sizeof(vpt_gpu_vdb), sizeof(vpt_sphere), sizeof(vpt_geometry_list),
                         sizeof(vpt_bvhnode), sizeof(vpt_octnode), sizeof(vpt_atmosphere), sizeof(vpt_kernel_params),
                         sizeof(vpt_point_light), sizeof(vpt_vdb_info), sizeof(vpt_aabb) };
    const int m = (int)(sizeof(s) / sizeof(s[0]));
    for (int i = 0; i < n && i < m; ++i) out[i] = s[i];
    return m;
}

const char* vpt_last_error(const vpt_context* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

static void free_context(vpt_context* c) {
    if (!c) return;
    cudaFree(c->d_scene); cudaFree(c->d_internal); cudaFree(c->d_leaf_list); cudaFree(c->d_leaf_indices); cudaFree(c->d_vrec);
    cudaFree(c->d_stats);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    vpt_comm_destroy(c);
    for (auto& e : c->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaFree(c->d_queue_id); cudaFree(c->d_queue_org); cudaFree(c->d_bn_table);
    cudaFree(c->d_counters); cudaFree(c->d_queue); cudaFree(c->d_planeA); cudaFree(c->d_planeB); cudaFree(c->d_planeC); cudaFree(c->d_planeD);
    delete c;
}

int vpt_create(vpt_context** out) {
    if (!out) return fail(nullptr, VPT_ERR_INVALID, "vpt_create: null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VPT_ERR_CUDA, std::string("vpt_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    vpt_context* c = new vpt_context();
    // every failure below releases what was allocated so far (the context never escapes half-built)
#define VPT_CREATE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        const std::string m_ = std::string("vpt_create: " #call ": ") + cudaGetErrorString(e_); free_context(c); return fail(nullptr, VPT_ERR_CUDA, m_); } } while (0)
    VPT_CREATE(cudaGetDevice(&c->device));
    cudaDeviceProp prop;
    VPT_CREATE(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 10) { std::string m = "vpt_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ", kernels are built for sm_100a only"; free_context(c); return fail(nullptr, VPT_ERR_UNSUPPORTED, m); }
    c->num_sms = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    VPT_CREATE(cudaMalloc(&c->d_scene, sizeof(vpt::SceneTables)));
    VPT_CREATE(cudaMalloc(&c->d_counters, sizeof(unsigned) * 4));
    VPT_CREATE(cudaMalloc(&c->d_stats, sizeof(unsigned long long) * 8));
    VPT_CREATE(cudaMemset(c->d_stats, 0, sizeof(unsigned long long) * 8));
    VPT_CREATE(cudaHostAlloc((void**)&c->h_pinned, sizeof(int) * 4, cudaHostAllocDefault));
    // per-DEVICE kernel attributes (dynamic shared memory opt-in) and the occupancy the persistent grid is sized with
    VPT_CREATE(vpt::trace_kernels_init(c->max_ctas));
#undef VPT_CREATE
    *out = c;
    return VPT_OK;
}

void vpt_destroy(vpt_context* c) { free_context(c); }

int vpt_set_option(vpt_context* c, const char* key, int value) {
    if (!c || !key) return fail(c, VPT_ERR_INVALID, "vpt_set_option: null argument");
    const std::string k(key);
    if (k == "passes_per_chunk") { if (value < 0 || value > 64) return fail(c, VPT_ERR_INVALID, "passes_per_chunk must be 0 (automatic) or 1..64"); c->chunk_auto = value == 0; if (value) c->passes_per_chunk = value; }
    else if (k == "max_scratch_mb") { if (value < 64) return fail(c, VPT_ERR_INVALID, "max_scratch_mb must be >= 64"); c->max_scratch_bytes = (size_t)value << 20; }
    else if (k == "gather_async") { c->gather_async = value ? 1 : 0; }
    else if (k == "l2_sector_fetch") { c->l2_sector_fetch = value ? 1 : 0; }
    else if (k == "trace_slots") { if (value != 0 && value != 2 && value != 3) return fail(c, VPT_ERR_INVALID, "trace_slots must be 0 (by grid size), 2 or 3"); c->trace_slots = value; }
    else if (k == "ctas_per_sm") { if (value < 0 || value > 8) return fail(c, VPT_ERR_INVALID, "ctas_per_sm must be 0..8"); c->ctas_per_sm = value; }
    else if (k == "sched_min_lanes") { if (value < 0 || value > 32) return fail(c, VPT_ERR_INVALID, "sched_min_lanes must be 0 (by kernel) or 1..32"); c->sched_min_lanes = value; }
    else if (k == "debug_flags") { c->debug_flags = value; }
    else if (k == "generic_kernel") { c->force_generic = value ? 1 : 0; }
    else if (k == "count_stats") { c->count_stats = value ? 1 : 0; }
    else if (k == "profile") { c->profile = value ? 1 : 0; }
    else return fail(c, VPT_ERR_INVALID, "unknown option " + k);
    return VPT_OK;
}

int vpt_set_partition(vpt_context* c, int rank, int n_ranks, int stripe_rows) {
    if (!c || n_ranks < 1 || rank < 0 || rank >= n_ranks || stripe_rows < 1) return fail(c, VPT_ERR_INVALID, "vpt_set_partition: bad arguments");
    c->rank = rank; c->n_ranks = n_ranks; c->stripe_rows = stripe_rows;
    return VPT_OK;
}

long long vpt_local_pixels(const vpt_context* c, unsigned width, unsigned height) {
    if (!c) return -1;
    return make_geom(c, width, height).n_local;
}

int vpt_unpermute(vpt_context* c, const void* d_gathered, void* d_full, unsigned width, unsigned height, int elem_bytes, void* stream) {
    if (!c || !d_gathered || !d_full || elem_bytes <= 0 || (elem_bytes & 3)) return fail(c, VPT_ERR_INVALID, "vpt_unpermute: bad arguments");
    const vpt::FrameGeom g = make_geom(c, width, height);
    VPT_CUDA(c, vpt::launch_unpermute(d_gathered, d_full, g, elem_bytes, (cudaStream_t)stream));
    c->launches++;
    return VPT_OK;
}

int vpt_invalidate_scene(vpt_context* c) {
    if (!c) return VPT_ERR_INVALID;
    c->cached_volumes = 0; c->cached_root = 0; c->cached_generation = 0;
    return VPT_OK;
}

int vpt_get_stats(vpt_context* c, unsigned long long* launches, unsigned* last_queue_count) {
    if (!c) return VPT_ERR_INVALID;
    if (launches) *launches = c->launches;
    if (last_queue_count) VPT_CUDA(c, cudaMemcpy(last_queue_count, c->d_counters, sizeof(unsigned), cudaMemcpyDeviceToHost));
    return VPT_OK;
}

static int ensure_frame_buffers(vpt_context* c, size_t n_samples, bool need_planeD) {
    if (n_samples > c->cap_samples) {
        cudaFree(c->d_queue); cudaFree(c->d_planeA); cudaFree(c->d_planeB); cudaFree(c->d_planeC); cudaFree(c->d_planeD);
        cudaFree(c->d_queue_id); cudaFree(c->d_queue_org);
        c->d_queue = c->d_planeA = c->d_planeB = c->d_planeC = c->d_planeD = c->d_queue_org = nullptr; c->d_queue_id = nullptr;
        c->cap_samples = 0; c->cap_planeD = false;
        VPT_CUDA(c, cudaMalloc(&c->d_queue, n_samples * sizeof(float4)));
        VPT_CUDA(c, cudaMalloc(&c->d_queue_id, n_samples * sizeof(uint2)));
        VPT_CUDA(c, cudaMalloc(&c->d_queue_org, n_samples * sizeof(float4)));
        VPT_CUDA(c, cudaMalloc(&c->d_planeA, n_samples * sizeof(float4)));
        VPT_CUDA(c, cudaMalloc(&c->d_planeB, n_samples * sizeof(float4)));
        VPT_CUDA(c, cudaMalloc(&c->d_planeC, n_samples * sizeof(float4)));
        c->cap_samples = n_samples;
    }
    if (need_planeD && !c->cap_planeD) {
        VPT_CUDA(c, cudaMalloc(&c->d_planeD, c->cap_samples * sizeof(float4)));
        c->cap_planeD = true;
    }
    return VPT_OK;
}

int vpt_render_passes(vpt_context* c, void* const params[VPT_NUM_ARGS], unsigned n_passes, void* stream_) {
    if (!c || !params) return fail(c, VPT_ERR_INVALID, "vpt_render_passes: null argument");
    for (int i = 0; i < VPT_NUM_ARGS; ++i) if (!params[i]) return fail(c, VPT_ERR_INVALID, "vpt_render_passes: params[" + std::to_string(i) + "] is null");
    if (n_passes == 0) return VPT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;

    vpt::FrameArgs fa;
    memcpy(&fa.cam, params[VPT_ARG_CAMERA], sizeof(vpt_camera));
    memcpy(&fa.lights, params[VPT_ARG_LIGHTS], sizeof(vpt_light_list));
    memcpy(&fa.kp, params[VPT_ARG_KERNEL_PARAMS], sizeof(vpt_kernel_params));
    vpt_devptr_t d_volumes, d_sphere, d_root;
    memcpy(&d_volumes, params[VPT_ARG_VOLUMES], 8);
    memcpy(&d_sphere, params[VPT_ARG_SPHERE], 8);
    memcpy(&d_root, params[VPT_ARG_OCTREE], 8);
    const vpt_kernel_params& kp = fa.kp;

    // environment_type == 0: the caller's AtmosphereParameters (scalars + the four precomputed look-up textures)
    vpt_atmosphere atmo;
    memcpy(&atmo, params[VPT_ARG_ATMOSPHERE], sizeof(vpt_atmosphere));
    if (c->n_ranks > 1 && (kp.resolution.x > 65535u || kp.resolution.y > 65535u)) return fail(c, VPT_ERR_UNSUPPORTED, "a partitioned frame is limited to 65535 x 65535 pixels");
    const bool vol_integ = (kp.integrator != 0);             // vol_integrator always ends on the precomputed sky (:1752)
    const bool sky_env = (kp.environment_type == 0) || vol_integ;
    const bool samples = kp.render && kp.iteration < kp.max_interactions;
    if (sky_env && samples &&
        (!atmo.transmittance_texture || !atmo.scattering_texture || !atmo.irradiance_texture || !atmo.single_mie_scattering_texture))
        return fail(c, VPT_ERR_INVALID, "environment_type == 0 / integrator != 0 need the four precomputed atmosphere textures in AtmosphereParameters");
    if (vol_integ && samples && kp.environment_type == 0 && kp.sky_mult > 0.0f &&
        (!kp.env_func_tex || !kp.env_cdf_tex || !kp.env_marginal_func_tex || !kp.env_marginal_cdf_tex || kp.env_sample_tex_res < 2))
        return fail(c, VPT_ERR_INVALID, "integrator != 0 with environment_type == 0 needs the env sampling tables (env_*_tex, env_sample_tex_res) in Kernel_params");
    if (vol_integ && samples && kp.environment_type != 0 && !kp.env_tex)
        return fail(c, VPT_ERR_INVALID, "environment_type != 0 needs Kernel_params.env_tex");
    const vpt_atmosphere* sky = sky_env ? &atmo : nullptr;
    if (!d_volumes || !d_sphere || !d_root) return fail(c, VPT_ERR_INVALID, "volumes / sphere / octree device pointer is null");
    if (!kp.accum_buffer || !kp.depth_buffer || !kp.cost_buffer || !kp.display_buffer || !kp.raw_buffer || !kp.blue_noise_buffer)
        return fail(c, VPT_ERR_INVALID, "Kernel_params output buffer pointer is null");
    if (kp.resolution.x == 0 || kp.resolution.y == 0) return fail(c, VPT_ERR_INVALID, "resolution is zero");

    // ---- scene tables ---------------------------------------------------------------------------------------
    // Octrees built by vpt_octree_build are registered with their flat tables and instance count: nothing is read
    // back and nothing blocks.  A foreign (reference-built, pointer-linked) octree is flattened on the device and
    // its instance count read back (4 bytes, one stream synchronisation) when the pointers change or a frame starts.
    {
        vpt::SceneEntry ent;
        const bool registered = vpt::scene_registry_find(d_root, &ent);
        if (registered && ent.device != c->device) return fail(c, VPT_ERR_INVALID, "octree was built on another device");
        const bool key_changed = c->cached_volumes != d_volumes || c->cached_root != d_root || c->cached_generation != (registered ? ent.generation : 0ull);
        // iteration == 0 starts a new frame: every scene edit of the reference application resets it (main.cpp:1667-1780), so
        // refreshing the cheap tables there also covers buffers that were rewritten or reallocated at the same address
        if (key_changed || kp.iteration == 0) {
            if (registered) {
                if ((size_t)ent.n > c->cap_vrec) {
                    cudaFree(c->d_vrec); c->d_vrec = nullptr; c->cap_vrec = 0;
                    VPT_CUDA(c, cudaMalloc(&c->d_vrec, sizeof(vpt::VolumeRec) * (size_t)ent.n));
                    c->cap_vrec = (size_t)ent.n;
                }
                vpt::SceneTables hdr;
                for (int a = 0; a < 3; ++a) { hdr.root_pmin[a] = ent.root6[a]; hdr.root_pmax[a] = ent.root6[3 + a]; }
                hdr.max_extinction = ent.max_extinction; hdr.min_extinction = ent.min_extinction;
                hdr.num_volumes = ent.n; hdr.single_volume = ent.n == 1 ? 1 : 0;
                hdr.internal = ent.d_internal; hdr.leaf_list = ent.d_leaf_list; hdr.leaf_indices = ent.d_leaf_indices; hdr.volumes = c->d_vrec; hdr.leaf_nodes = nullptr;
                VPT_CUDA(c, vpt::launch_prepare_volumes(reinterpret_cast<const vpt_gpu_vdb*>(d_volumes), hdr, c->d_scene, c->d_vrec, stream));
                c->scene_single_volume = ent.n == 1;
            } else {
                if (!c->d_internal) {
                    VPT_CUDA(c, cudaMalloc(&c->d_internal, sizeof(vpt::OctInternal) * vpt::kOctInternalNodes));
                    VPT_CUDA(c, cudaMalloc(&c->d_leaf_list, sizeof(uint2) * vpt::kOctLeaves));
                    VPT_CUDA(c, cudaMalloc(&c->d_leaf_indices, sizeof(int) * vpt::kOctLeaves * VPT_OCT_MAX_VOLUMES));
                }
                if (c->cap_vrec < VPT_OCT_MAX_VOLUMES) {
                    cudaFree(c->d_vrec); c->d_vrec = nullptr; c->cap_vrec = 0;
                    VPT_CUDA(c, cudaMalloc(&c->d_vrec, sizeof(vpt::VolumeRec) * VPT_OCT_MAX_VOLUMES));
                    c->cap_vrec = VPT_OCT_MAX_VOLUMES;
                }
                VPT_CUDA(c, vpt::launch_prepare_scene(reinterpret_cast<const vpt_gpu_vdb*>(d_volumes), reinterpret_cast<const vpt_octnode*>(d_root),
                                                      c->d_scene, c->d_internal, c->d_leaf_list, c->d_leaf_indices, c->d_vrec, VPT_OCT_MAX_VOLUMES, stream));
                VPT_CUDA(c, cudaMemcpyAsync(c->h_pinned, reinterpret_cast<const char*>(c->d_scene) + offsetof(vpt::SceneTables, single_volume), sizeof(int),
                                            cudaMemcpyDeviceToHost, stream));
                VPT_CUDA(c, cudaStreamSynchronize(stream));
                c->scene_single_volume = c->h_pinned[0] != 0;
            }
            c->launches++;
            c->cached_volumes = d_volumes; c->cached_root = d_root; c->cached_generation = registered ? ent.generation : 0ull;
        }
    }
    if (c->brick_pool || c->cell_table) {
        // brick / cell mode read volume 0 from their own copy of the grid: only the lean case of the direct integrator is built for them
        vpt::SceneEntry ent;
        const bool ok = vpt::scene_registry_find(d_root, &ent) && ent.n == 1 && !(ent.any_flags & 1u) && !(kp.emission_scale > 0.0f) &&
                        fa.lights.num_lights == 0 && kp.integrator == 0 && !c->force_generic;
        if (!ok) return fail(c, VPT_ERR_UNSUPPORTED, "brick / cell mode needs a vpt_octree_build scene of ONE volume without colour grid, the direct integrator, no emission and no point lights");
        if (c->brick_pool && c->cell_table) return fail(c, VPT_ERR_INVALID, "brick mode and cell mode are both set");
    }
    // "lean" = nothing but one volume, sun and environment in play (the headline configuration)
    const bool lean = c->scene_single_volume && !(kp.emission_scale > 0.0f) && fa.lights.num_lights == 0 && kp.integrator == 0 && !c->force_generic;

    fa.sphere = reinterpret_cast<const vpt_sphere*>(d_sphere);
    fa.scene = c->d_scene;
    fa.geom = make_geom(c, kp.resolution.x, kp.resolution.y);

    // how many of the requested passes actually sample (render_kernel.cu:2254)
    unsigned n_sampled = 0;
    if (kp.render && kp.iteration < kp.max_interactions) {
        const unsigned left = kp.max_interactions - kp.iteration;
        n_sampled = n_passes < left ? n_passes : left;
    }

    // passes fused per generate/trace/resolve round: 32 by default; a small local frame (multi-GPU shard, low resolution)
    // takes up to 64 so the fixed per-round cost and the persistent kernel's tail are paid half as often; always capped so
    // the per-round scratch (ray queue + sample planes, 88 B or 104 B per sample) stays under max_scratch_bytes
    const bool planeD = sky_env;
    int chunk = c->passes_per_chunk;
    if (c->chunk_auto) chunk = (size_t)fa.geom.n_local <= ((size_t)1 << 20) ? 64 : 32;
    {
        const size_t per_sample = 88 + (planeD ? 16 : 0);
        const size_t fit = c->max_scratch_bytes / (per_sample * (size_t)(fa.geom.n_local > 0 ? fa.geom.n_local : 1));
        if ((size_t)chunk > fit) chunk = fit < 1 ? 1 : (int)fit;
    }
    if (n_sampled) {
        const size_t per_chunk = (size_t)fa.geom.n_local * (size_t)(n_sampled < (unsigned)chunk ? n_sampled : (unsigned)chunk);
        int rc = ensure_frame_buffers(c, per_chunk, planeD);
        if (rc != VPT_OK) return rc;
    }
    if (c->cap_chunk < chunk) {
        cudaFree(c->d_bn_table); c->d_bn_table = nullptr;
        VPT_CUDA(c, cudaMalloc(&c->d_bn_table, sizeof(float2) * 65536 * (size_t)chunk));
        c->cap_chunk = chunk;
    }
    fa.queue_dir = c->d_queue; fa.queue_id = c->d_queue_id; fa.queue_aux = c->d_queue_org; fa.thin_lens = (fa.cam.lens_radius != 0.0f) ? 1 : 0;
    fa.bn_table = c->d_bn_table; fa.debug_flags = c->debug_flags;     // fa.sched_min_lanes: set below, once the trace kernel is chosen
    fa.queue_count = c->d_counters; fa.queue_head = c->d_counters + 1;
    fa.planeA = c->d_planeA; fa.planeB = c->d_planeB; fa.planeC = c->d_planeC; fa.planeD = planeD ? c->d_planeD : nullptr;

    // rays per lane: 3 by default; 2 (one more CTA per SM) when a look-up is a chain of cache misses -- a density grid that cannot live
    // in L2 (every look-up a DRAM round trip) or long per-leaf instance lists (dependent loads): warps in flight then matter more than
    // rays per warp.  Measured: 1024^3 grid 158 -> 104 ms, 512^3 92 -> 77 ms, 1000 instances 94 -> 82 ms; the
    // L2-resident scenes lose with it (dragon 7.1 -> 8.2 ms, fireball 185 -> 282 ms).
    int slots = 3;
    if (c->cell_table) slots = 2;
    else if (!vol_integ && !c->brick_pool) {
        if (c->trace_slots) slots = c->trace_slots;
        else {
            vpt::SceneEntry ent;
            if (vpt::scene_registry_find(d_root, &ent) && (ent.max_grid_bytes > 2ull * c->l2_bytes || ent.n >= 64)) slots = 2;
        }
    }
    // lanes an operation must gather before it pre-empts stepping: measured per kernel (headline 6.83 -> 6.80 ms with 26; fireball, 1024^3 grid
    // and 1000 instances lose 0.6 - 2.7 % with it)
    fa.sched_min_lanes = c->sched_min_lanes > 0 ? c->sched_min_lanes : ((lean && slots == 3 && !vol_integ && !c->brick_pool && !c->cell_table) ? 26 : 20);
    int ctas_per_sm = c->ctas_per_sm > 0 ? c->ctas_per_sm : c->max_ctas[c->brick_pool ? 3 : c->cell_table ? 6 : vol_integ ? 2 : (slots == 2 ? (lean ? 5 : 4) : (lean ? 1 : 0))];
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int trace_ctas = c->num_sms * ctas_per_sm;

    fa.counters = c->count_stats ? c->d_stats : nullptr;
    fa.cell_table = c->cell_table; fa.cell_nx = c->cell_dims[0]; fa.cell_ny = c->cell_dims[1]; fa.cell_nz = c->cell_dims[2];
    auto timed = [&](int kind, auto&& fn) -> cudaError_t {
        if (!c->profile) return fn();
        vpt_context::Ev ev; ev.kind = kind;
        cudaEventCreate(&ev.a); cudaEventCreate(&ev.b);
        cudaEventRecord(ev.a, stream);
        cudaError_t e = fn();
        cudaEventRecord(ev.b, stream);
        c->events.push_back(ev);
        return e;
    };
    if (c->p2p_on && ((int)kp.resolution.x != c->p2p_w || (int)kp.resolution.y != c->p2p_h))
        return fail(c, VPT_ERR_INVALID, "the peer exchange block was exported for another frame size (vpt_comm_p2p_export)");
    { int rc = vpt::comm_p2p_begin(c, stream); if (rc != VPT_OK) return rc; }      // multi-GPU peer exchange: this rank's frame may be overwritten from here on
    const uint32_t it0 = kp.iteration;
    const unsigned long long frame_px = (unsigned long long)kp.resolution.x * kp.resolution.y;
    const int bn_limit = frame_px < 65536ull ? (int)frame_px : 65536;
    unsigned done = 0;
    while (done < n_sampled) {
        const unsigned np = (n_sampled - done) < (unsigned)chunk ? (n_sampled - done) : (unsigned)chunk;
        fa.kp.iteration = it0 + done;
        VPT_CUDA(c, timed(3, [&] { return vpt::launch_bn_prepare((void*)kp.blue_noise_buffer, c->d_bn_table, (int)np, bn_limit, c->d_counters, stream); }));   // jitter table + advance + queue reset
        VPT_CUDA(c, timed(0, [&] { return vpt::launch_generate(fa, (int)np, stream); }));
        if (c->brick_pool) VPT_CUDA(c, timed(1, [&] { return vpt::launch_trace_brick(fa, c->brick_pool, c->brick_dims, trace_ctas, stream); }));
        else VPT_CUDA(c, timed(1, [&] { return vpt::launch_trace(fa, vol_integ ? &atmo : nullptr, lean, slots, trace_ctas, stream); }));
        const bool last = (done + np == n_passes);
        if (c->gather_pending) { int rc = vpt::comm_before_accum_write(c, stream); if (rc != VPT_OK) return rc; }
        vpt::PeerFrames peers;
        if (last) { int rc = vpt::comm_p2p_peers(c, stream, &peers); if (rc != VPT_OK) return rc; }
        VPT_CUDA(c, timed(2, [&] { return vpt::launch_resolve(fa, sky, (int)np, 1, last ? 1 : 0, last ? &peers : nullptr, stream); }));
        c->launches += 4;
        done += np;
    }
    if (n_sampled < n_passes) {
        // passes that no longer sample: WHITE / re-tonemap semantics of the kernel tail
        fa.kp.iteration = it0 + n_sampled;
        if (c->gather_pending) { int rc = vpt::comm_before_accum_write(c, stream); if (rc != VPT_OK) return rc; }
        vpt::PeerFrames peers;
        { int rc = vpt::comm_p2p_peers(c, stream, &peers); if (rc != VPT_OK) return rc; }
        VPT_CUDA(c, vpt::launch_resolve(fa, sky, (int)(n_passes - n_sampled), 0, 1, &peers, stream));
        VPT_CUDA(c, vpt::launch_bn_advance((void*)kp.blue_noise_buffer, (int)(n_passes - n_sampled), bn_limit, stream));
        c->launches += 2;
    }
    // multi-GPU: the one exchange of the path.  Peer-memory mode: the last resolve above already stored every pixel into every rank's
    // frame; publish and wait for the peers.  NCCL mode: all-gather of the rank-local frame into the caller's full-frame buffers.
    { int rc = vpt::comm_p2p_end(c, stream); if (rc != VPT_OK) return rc; }
    if (c->nccl_comm && (c->d_full_accum || c->d_full_display)) {
        int rc = vpt::comm_gather_frame(c, fa.geom, (const void*)kp.accum_buffer, (const void*)kp.display_buffer, stream);
        if (rc != VPT_OK) return rc;
    }
    return VPT_OK;
}

// Statistics accumulated since the last call (needs option "count_stats"): out[0..4] = volume lookups, lane-steps,
// warp step iterations, lane transitions, warp transition rounds.  Synchronises the device.
int vpt_get_counters(vpt_context* c, unsigned long long out[8], int reset) {
    if (!c || !out) return VPT_ERR_INVALID;
    VPT_CUDA(c, cudaDeviceSynchronize());
    VPT_CUDA(c, cudaMemcpy(out, c->d_stats, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost));
    if (reset) VPT_CUDA(c, cudaMemset(c->d_stats, 0, sizeof(unsigned long long) * 8));
    return VPT_OK;
}

// Per-kernel device time accumulated since the last call (needs option "profile"):
// ms[0..3] = generate, trace, resolve, bn_advance; n[0..3] = launches of each.  Synchronises the device.
int vpt_get_kernel_times(vpt_context* c, float ms[4], int n[4]) {
    if (!c || !ms || !n) return VPT_ERR_INVALID;
    VPT_CUDA(c, cudaDeviceSynchronize());
    for (int i = 0; i < 4; ++i) { ms[i] = 0.f; n[i] = 0; }
    for (auto& e : c->events) {
        float t = 0.f; cudaEventElapsedTime(&t, e.a, e.b);
        if (e.kind >= 0 && e.kind < 4) { ms[e.kind] += t; n[e.kind]++; }
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    c->events.clear();
    return VPT_OK;
}

int vpt_render_pass(vpt_context* c, void* const params[VPT_NUM_ARGS], void* stream) {
    return vpt_render_passes(c, params, 1, stream);
}

// ------------------------------------------------------------------------------------------------------
// host-side scene helpers
// ------------------------------------------------------------------------------------------------------
int vpt_texture_create_3d(const float* host, int channels, int dx, int dy, int dz, vpt_tex_t* tex_out, void** array_out) {
    if (!host || !tex_out || !array_out || (channels != 1 && channels != 4) || dx < 1 || dy < 1 || dz < 1)
        return fail(nullptr, VPT_ERR_INVALID, "vpt_texture_create_3d: bad arguments");
    cudaChannelFormatDesc desc = channels == 1 ? cudaCreateChannelDesc<float>() : cudaCreateChannelDesc<float4>();
    cudaExtent ext = make_cudaExtent((size_t)dx, (size_t)dy, (size_t)dz);
    cudaArray_t arr = nullptr;
    VPT_CUDA(nullptr, cudaMalloc3DArray(&arr, &desc, ext));
    cudaMemcpy3DParms cp; memset(&cp, 0, sizeof(cp));
    const size_t esz = sizeof(float) * (size_t)channels;
    cp.srcPtr = make_cudaPitchedPtr((void*)host, (size_t)dx * esz, (size_t)dx, (size_t)dy);
    cp.dstArray = arr; cp.extent = ext; cp.kind = cudaMemcpyHostToDevice;
    VPT_CUDA(nullptr, cudaMemcpy3D(&cp));
    cudaResourceDesc res; memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray; res.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.normalizedCoords = 1; td.filterMode = cudaFilterModeLinear;
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    VPT_CUDA(nullptr, cudaCreateTextureObject(&tex, &res, &td, NULL));
    *tex_out = (vpt_tex_t)tex; *array_out = (void*)arr;
    return VPT_OK;
}

int vpt_texture_create_3d_from_device(const float* d_data, int channels, int dx, int dy, int dz, vpt_tex_t* tex_out, void** array_out) {
    if (!d_data || !tex_out || !array_out || (channels != 1 && channels != 4) || dx < 1 || dy < 1 || dz < 1)
        return fail(nullptr, VPT_ERR_INVALID, "vpt_texture_create_3d_from_device: bad arguments");
    cudaChannelFormatDesc desc = channels == 1 ? cudaCreateChannelDesc<float>() : cudaCreateChannelDesc<float4>();
    cudaExtent ext = make_cudaExtent((size_t)dx, (size_t)dy, (size_t)dz);
    cudaArray_t arr = nullptr;
    VPT_CUDA(nullptr, cudaMalloc3DArray(&arr, &desc, ext));
    cudaMemcpy3DParms cp; memset(&cp, 0, sizeof(cp));
    const size_t esz = sizeof(float) * (size_t)channels;
    cp.srcPtr = make_cudaPitchedPtr((void*)d_data, (size_t)dx * esz, (size_t)dx, (size_t)dy);
    cp.dstArray = arr; cp.extent = ext; cp.kind = cudaMemcpyDeviceToDevice;
    cudaError_t e = cudaMemcpy3D(&cp);
    cudaResourceDesc res; memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray; res.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.normalizedCoords = 1; td.filterMode = cudaFilterModeLinear;
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    if (e == cudaSuccess) e = cudaCreateTextureObject(&tex, &res, &td, NULL);
    if (e != cudaSuccess) { cudaFreeArray(arr); return fail(nullptr, VPT_ERR_CUDA, std::string("vpt_texture_create_3d_from_device: ") + cudaGetErrorString(e)); }
    *tex_out = (vpt_tex_t)tex; *array_out = (void*)arr;
    return VPT_OK;
}

int vpt_procedural_fill(float* d_buffer, int dx, int dy, int dz, int noise_type, float scale, int seed, void* stream) {
    if (!d_buffer || dx < 1 || dy < 1 || dz < 1) return fail(nullptr, VPT_ERR_INVALID, "vpt_procedural_fill: bad arguments");
    if (noise_type != 0) return fail(nullptr, VPT_ERR_UNSUPPORTED, "vpt_procedural_fill: only noise type 0 (Perlin gradient noise, the reference's default) is built");
    VPT_CUDA(nullptr, vpt::launch_fill_perlin(d_buffer, dx, dy, dz, scale, seed, (cudaStream_t)stream));
    return VPT_OK;
}

int vpt_bricks_create(const float* d_dense, int dx, int dy, int dz, vpt_devptr_t* d_pool_out, unsigned long long* bytes_out) {
    if (!d_dense || !d_pool_out || dx < 1 || dy < 1 || dz < 1) return fail(nullptr, VPT_ERR_INVALID, "vpt_bricks_create: bad arguments");
    const size_t nb = (size_t)((dx + 3) / 4) * ((dy + 3) / 4) * ((dz + 3) / 4);
    if (nb > (size_t)0x7fffffff) return fail(nullptr, VPT_ERR_UNSUPPORTED, "vpt_bricks_create: more than 2^31 bricks");
    float* pool = nullptr;
    VPT_CUDA(nullptr, cudaMalloc(&pool, nb * 512));
    cudaError_t e = vpt::launch_build_bricks(d_dense, dx, dy, dz, pool, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(pool); return fail(nullptr, VPT_ERR_CUDA, std::string("vpt_bricks_create: ") + cudaGetErrorString(e)); }
    *d_pool_out = (vpt_devptr_t)(uintptr_t)pool;
    if (bytes_out) *bytes_out = (unsigned long long)nb * 512ull;
    return VPT_OK;
}

int vpt_bricks_read(vpt_devptr_t d_pool, unsigned long long first_brick, unsigned long long n_bricks, float* h_out) {
    if (!d_pool || !h_out) return fail(nullptr, VPT_ERR_INVALID, "vpt_bricks_read: null argument");
    VPT_CUDA(nullptr, cudaMemcpy(h_out, reinterpret_cast<const char*>((uintptr_t)d_pool) + first_brick * 512ull, n_bricks * 512ull, cudaMemcpyDeviceToHost));
    return VPT_OK;
}

int vpt_debug_sampler_compare(vpt_tex_t tex, vpt_devptr_t d_pool, int dx, int dy, int dz, int n_points, unsigned seed, double out12[12]) {
    if (!tex || !d_pool || !out12 || n_points < 1) return fail(nullptr, VPT_ERR_INVALID, "vpt_debug_sampler_compare: bad arguments");
    double* d_out = nullptr;
    VPT_CUDA(nullptr, cudaMalloc(&d_out, sizeof(double) * 12));
    cudaError_t e = cudaMemset(d_out, 0, sizeof(double) * 12);
    const int dims[3] = { dx, dy, dz };
    if (e == cudaSuccess) e = vpt::launch_sampler_compare(tex, reinterpret_cast<const float*>((uintptr_t)d_pool), dims, n_points, seed, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out12, d_out, sizeof(double) * 12, cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(nullptr, VPT_ERR_CUDA, std::string("vpt_debug_sampler_compare: ") + cudaGetErrorString(e));
    for (int m = 0; m < 3; ++m) out12[m * 4 + 3] = (double)n_points;
    return VPT_OK;
}

int vpt_cells_create(const float* d_dense, int dx, int dy, int dz, vpt_devptr_t* d_cells_out, unsigned long long* bytes_out) {
    if (!d_dense || !d_cells_out || dx < 1 || dy < 1 || dz < 1) return fail(nullptr, VPT_ERR_INVALID, "vpt_cells_create: bad arguments");
    const size_t n = (size_t)dx * dy * dz;
    float4* cells = nullptr;
    cudaError_t e = cudaMalloc(&cells, n * 32);
    if (e != cudaSuccess) return fail(nullptr, VPT_ERR_CUDA, std::string("vpt_cells_create: ") + std::to_string(n * 32 >> 20) + " MiB: " + cudaGetErrorString(e));
    e = vpt::launch_build_cells(d_dense, dx, dy, dz, cells, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(cells); return fail(nullptr, VPT_ERR_CUDA, std::string("vpt_cells_create: ") + cudaGetErrorString(e)); }
    *d_cells_out = (vpt_devptr_t)(uintptr_t)cells;
    if (bytes_out) *bytes_out = (unsigned long long)n * 32ull;
    return VPT_OK;
}

int vpt_cells_read(vpt_devptr_t d_cells, unsigned long long first_cell, unsigned long long n_cells, float* h_out) {
    if (!d_cells || !h_out) return fail(nullptr, VPT_ERR_INVALID, "vpt_cells_read: null argument");
    VPT_CUDA(nullptr, cudaMemcpy(h_out, reinterpret_cast<const char*>((uintptr_t)d_cells) + first_cell * 32ull, n_cells * 32ull, cudaMemcpyDeviceToHost));
    return VPT_OK;
}

int vpt_cells_destroy(vpt_devptr_t d_cells) { if (d_cells) cudaFree((void*)(uintptr_t)d_cells); return VPT_OK; }

int vpt_set_cell_volume(vpt_context* c, vpt_devptr_t d_cells, int dx, int dy, int dz) {
    if (!c) return VPT_ERR_INVALID;
    if (d_cells && (dx < 1 || dy < 1 || dz < 1)) return fail(c, VPT_ERR_INVALID, "vpt_set_cell_volume: bad dimensions");
    c->cell_table = reinterpret_cast<const float4*>((uintptr_t)d_cells);
    c->cell_dims[0] = dx; c->cell_dims[1] = dy; c->cell_dims[2] = dz;
    // a cell is one 32-byte sector: option "l2_sector_fetch" asks the L2 not to widen a miss to a larger DRAM fetch while this mode is on
    // (cudaLimitMaxL2FetchGranularity, a DEVICE-wide hint: the streaming kernels lose ~4 % with it, so it is not the default)
    if (d_cells && c->l2_sector_fetch) {
        if (!c->l2_fetch_default) cudaDeviceGetLimit(&c->l2_fetch_default, cudaLimitMaxL2FetchGranularity);
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    } else if (c->l2_fetch_default) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, c->l2_fetch_default);
    cudaGetLastError();
    return VPT_OK;
}

int vpt_bricks_destroy(vpt_devptr_t d_pool) { if (d_pool) cudaFree((void*)(uintptr_t)d_pool); return VPT_OK; }

int vpt_set_brick_volume(vpt_context* c, vpt_devptr_t d_pool, int dx, int dy, int dz) {
    if (!c) return VPT_ERR_INVALID;
    if (d_pool && (dx < 1 || dy < 1 || dz < 1)) return fail(c, VPT_ERR_INVALID, "vpt_set_brick_volume: bad dimensions");
    c->brick_pool = reinterpret_cast<const float*>((uintptr_t)d_pool);
    c->brick_dims[0] = dx; c->brick_dims[1] = dy; c->brick_dims[2] = dz;
    return VPT_OK;
}

int vpt_texture_create_env(const float* rgba, unsigned w, unsigned h, vpt_tex_t* tex_out, void** array_out) {
    if (!rgba || !tex_out || !array_out || !w || !h) return fail(nullptr, VPT_ERR_INVALID, "vpt_texture_create_env: bad arguments");
    const cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    cudaArray_t arr = nullptr;
    VPT_CUDA(nullptr, cudaMallocArray(&arr, &desc, w, h));
    VPT_CUDA(nullptr, cudaMemcpy2DToArray(arr, 0, 0, rgba, (size_t)w * sizeof(float4), (size_t)w * sizeof(float4), h, cudaMemcpyHostToDevice));
    cudaResourceDesc res; memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray; res.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeWrap; td.addressMode[1] = cudaAddressModeClamp; td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
    cudaTextureObject_t tex = 0;
    VPT_CUDA(nullptr, cudaCreateTextureObject(&tex, &res, &td, NULL));
    *tex_out = (vpt_tex_t)tex; *array_out = (void*)arr;
    return VPT_OK;
}

// point-sampled, unnormalised float table (h == 0: 1-D array), the descriptor main.cpp:788-867 uses for the env sampling tables
static int create_table_texture(const float* data, unsigned w, unsigned h, vpt_tex_t* tex_out, void** array_out) {
    const cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    cudaArray_t arr = nullptr;
    VPT_CUDA(nullptr, cudaMallocArray(&arr, &desc, w, h));
    VPT_CUDA(nullptr, cudaMemcpy2DToArray(arr, 0, 0, data, (size_t)w * sizeof(float), (size_t)w * sizeof(float), h ? h : 1, cudaMemcpyHostToDevice));
    cudaResourceDesc res; memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray; res.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeWrap; td.addressMode[1] = h ? cudaAddressModeClamp : cudaAddressModeWrap; td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex = 0;
    VPT_CUDA(nullptr, cudaCreateTextureObject(&tex, &res, &td, NULL));
    *tex_out = (vpt_tex_t)tex; *array_out = (void*)arr;
    return VPT_OK;
}

// Host arithmetic of create_cdf (main.cpp:681-757), pointer walk restated with indices.  The reference reads one element
// BEFORE two of its arrays (quirk Q20): `*(func_p - 1)` at y == 0, x == 0 and `*(marginal_cdf_p - 1)` at y == 0; those two reads
// are taken as 0 here (in the reference they hit the allocator's bookkeeping word: a denormal).  Everything else is literal:
//   * row y > 0 starts from func[y-1][res-1] / res, because `*(func_p - 1)` at x == 0 is the LAST element of the previous row
//     (in range), and that offset is carried through the row's cdf and into marginal_func[y];
//   * `*(cdf_p - 1) = 0` at each row start zeroes the previous row's last cdf entry, which the normalisation then sets to 1;
//   * the "total" that selects the uniform fallback is res x marginal_func[0] (the loop never advances its pointer).
int vpt_env_tables_compute(const float* func, unsigned res, float* cdf, float* marginal_func, float* marginal_cdf, float* marginal_int_out) {
    if (!func || res < 2 || !cdf || !marginal_func || !marginal_cdf || !marginal_int_out) return fail(nullptr, VPT_ERR_INVALID, "vpt_env_tables_compute: bad arguments");
    const float fres = (float)res;
    for (unsigned y = 0; y < res; ++y) {
        float prev_cdf = 0.0f;                                          // *(cdf_p - 1) = .0f
        if (y > 0) cdf[(size_t)y * res - 1] = 0.0f;
        for (unsigned x = 0; x < res; ++x) {
            const float prev_func = (x == 0 && y == 0) ? 0.0f : func[(size_t)y * res + x - 1];
            prev_cdf = prev_cdf + prev_func / fres;
            cdf[(size_t)y * res + x] = prev_cdf;
        }
        marginal_func[y] = prev_cdf;
    }
    float total_int = 0.0f;
    for (unsigned j = 0; j < res; ++j) total_int += marginal_func[0];
    if (total_int == 0.0f) {
        for (unsigned y = 0; y < res; ++y) for (unsigned x = 0; x < res; ++x) cdf[(size_t)y * res + x] = ((float)x / fres) * ((float)y / fres);
    } else {
        for (unsigned y = 0; y < res; ++y) for (unsigned x = 0; x < res; ++x) {
            float& c = cdf[(size_t)y * res + x];
            c /= marginal_func[y];
            if (x == res - 1) c = 1.0f;
        }
    }
    float run = 0.0f;                                                   // *(marginal_cdf_p - 1) at y == 0: out of range, taken as 0
    for (unsigned y = 0; y < res; ++y) { run = run + marginal_func[y] / fres; marginal_cdf[y] = run; }
    const float marginal_int = run;
    if (marginal_int > 0.0f) for (unsigned y = 0; y < res; ++y) marginal_cdf[y] /= std::max(.000001f, marginal_int);
    else marginal_cdf[0] = 1.0f;                                        // the reference's trailing `*marginal_cdf_p = 1.0f` lands here when the loop is skipped
    *marginal_int_out = marginal_int;
    return VPT_OK;
}

int vpt_env_tables_create(const float* func, unsigned res, vpt_tex_t tex_out[4], void* arrays_out[4], float* marginal_int_out) {
    if (!func || res < 2 || !tex_out || !arrays_out || !marginal_int_out) return fail(nullptr, VPT_ERR_INVALID, "vpt_env_tables_create: bad arguments");
    const size_t n = (size_t)res * res;
    std::vector<float> cdf(n), mfunc(res), mcdf(res);
    int rc0 = vpt_env_tables_compute(func, res, cdf.data(), mfunc.data(), mcdf.data(), marginal_int_out);
    if (rc0 != VPT_OK) return rc0;
    const float* src[4] = { func, cdf.data(), mfunc.data(), mcdf.data() };
    const unsigned heights[4] = { res, res, 0u, 0u };
    for (int i = 0; i < 4; ++i) { tex_out[i] = 0; arrays_out[i] = nullptr; }
    for (int i = 0; i < 4; ++i) {
        const int rc = create_table_texture(src[i], res, heights[i], &tex_out[i], &arrays_out[i]);
        if (rc != VPT_OK) {                                       // leave nothing half-built behind
            for (int j = 0; j < i; ++j) { vpt_texture_destroy(tex_out[j], arrays_out[j]); tex_out[j] = 0; arrays_out[j] = nullptr; }
            return rc;
        }
    }
    return VPT_OK;
}

int vpt_texture_destroy(vpt_tex_t tex, void* array) {
    if (tex) cudaDestroyTextureObject((cudaTextureObject_t)tex);
    if (array) cudaFreeArray((cudaArray_t)array);
    return VPT_OK;
}

void vpt_free(void* p) { free(p); }

int vpt_ins_load(const char* path, vpt_ins_header** out) {
    if (!path || !out) return fail(nullptr, VPT_ERR_INVALID, "vpt_ins_load: null argument");
    *out = nullptr;
    std::ifstream f(path);
    if (!f) return fail(nullptr, VPT_ERR_IO, std::string("vpt_ins_load: cannot open ") + path);
    auto bad = [&](const std::string& what) { return fail(nullptr, VPT_ERR_IO, std::string("vpt_ins_load: ") + path + ": " + what); };
    auto numbers = [](const std::string& line, double* v, int n) { std::istringstream is(line); for (int i = 0; i < n; ++i) if (!(is >> v[i])) return false; return true; };
    std::string line;
    if (!std::getline(f, line)) return bad("empty file");
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
    std::vector<vpt_ins_file_entry> files;
    std::vector<double> recs;
    int kind = 0;
    if (line == "light") {
        kind = 1;
        double n = 0;
        if (!std::getline(f, line) || !numbers(line, &n, 1) || n < 0 || n > 1e7) return bad("light count missing");
        for (int i = 0; i < (int)n; ++i) {
            double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (!std::getline(f, line) || !numbers(line, v, 7)) return bad("light record " + std::to_string(i) + " needs 7 numbers: px py pz r g b power");
            recs.insert(recs.end(), v, v + 8);
        }
    } else {
        double n = 0;
        if (!numbers(line, &n, 1) || n < 0 || n > 1e6) return bad("first line must be \"light\" or the number of .vdb entries");
        for (int i = 0; i < (int)n; ++i) {
            vpt_ins_file_entry e; memset(&e, 0, sizeof(e));
            if (!std::getline(f, line)) return bad("missing .vdb path of entry " + std::to_string(i));
            while (!line.empty() && (line.back() == '\r')) line.pop_back();
            if (line.size() >= sizeof(e.path)) return bad("path too long");
            memcpy(e.path, line.c_str(), line.size());
            double cnt = 0;
            if (!std::getline(f, line) || !numbers(line, &cnt, 1) || cnt < 0 || cnt > 1e7) return bad("missing instance count of entry " + std::to_string(i));
            e.first_record = (int32_t)(recs.size() / 8); e.n_instances = (int32_t)cnt;
            for (int x = 0; x < (int)cnt; ++x) {
                double v[8];
                if (!std::getline(f, line) || !numbers(line, v, 8)) return bad("instance " + std::to_string(x) + " of entry " + std::to_string(i) + " needs 8 numbers: px py pz qx qy qz qw scale");
                recs.insert(recs.end(), v, v + 8);
            }
            files.push_back(e);
        }
    }
    const size_t bytes = sizeof(vpt_ins_header) + files.size() * sizeof(vpt_ins_file_entry) + recs.size() * sizeof(double);
    char* blk = (char*)malloc(bytes);
    if (!blk) return fail(nullptr, VPT_ERR_IO, "vpt_ins_load: out of memory");
    vpt_ins_header h; h.kind = kind; h.n_files = (int32_t)files.size(); h.n_records = (int32_t)(recs.size() / 8); h.reserved = 0;
    memcpy(blk, &h, sizeof(h));
    if (!files.empty()) memcpy(blk + sizeof(h), files.data(), files.size() * sizeof(vpt_ins_file_entry));
    if (!recs.empty()) memcpy(blk + sizeof(h) + files.size() * sizeof(vpt_ins_file_entry), recs.data(), recs.size() * sizeof(double));
    *out = reinterpret_cast<vpt_ins_header*>(blk);
    return VPT_OK;
}

int vpt_vdb_load(const char* path, const char* grid_name, float** values_out, int info[13], float xform16[16], float stats[4]) {
    if (!path || !grid_name || !values_out) return fail(nullptr, VPT_ERR_INVALID, "vpt_vdb_load: null argument");
    *values_out = nullptr;
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(nullptr, VPT_ERR_IO, std::string("vpt_vdb_load: cannot open ") + path);
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    vpt::VdbDenseGrid g;
    try {
        if (!vpt::vdb_read_dense(data.data(), data.size(), grid_name, g)) return 1;
    } catch (const std::exception& e) { return fail(nullptr, VPT_ERR_IO, std::string("vpt_vdb_load: ") + e.what()); }
    // cross-check against the file's own metadata (the only pin the format offers)
    if (g.meta.has_file_bbox)
        for (int a = 0; a < 3; ++a)
            if (g.meta.file_bbox_min[a] != g.bbox_min[a] || g.meta.file_bbox_max[a] != g.bbox_max[a])
                return fail(nullptr, VPT_ERR_IO, "vpt_vdb_load: decoded active bbox disagrees with file_bbox metadata");
    if (g.meta.file_voxel_count >= 0 && (uint64_t)g.meta.file_voxel_count != g.active_leaf_voxels + g.active_tile_voxels)
        return fail(nullptr, VPT_ERR_IO, "vpt_vdb_load: decoded active voxel count disagrees with file_voxel_count metadata");
    const size_t nfl = g.values.size();
    float* out = (float*)malloc(nfl * sizeof(float));
    if (!out) return fail(nullptr, VPT_ERR_IO, "vpt_vdb_load: out of memory");
    memcpy(out, g.values.data(), nfl * sizeof(float));
    *values_out = out;
    if (info) {
        for (int a = 0; a < 3; ++a) { info[a] = g.dim[a]; info[3 + a] = g.bbox_min[a]; info[6 + a] = g.bbox_max[a]; }
        info[9] = g.channels; info[10] = (int)g.leaf_count;
        const uint64_t act = g.active_leaf_voxels + g.active_tile_voxels;
        info[11] = (int)(act & 0xffffffffu); info[12] = (int)(act >> 32);
    }
    if (xform16) {
        // convert_to_mat4: xform[i][j] = float(M(j, i)) with M the OpenVDB row-vector Mat4 => memory image m[c][r] = M[r][c]
        for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) xform16[i * 4 + j] = float(g.index_to_world[j * 4 + i]);
    }
    if (stats) {
        float mx = 0.0f, mn = FLT_MAX;                      // VDB_INFO defaults then the scan of gpu_vdb.cpp:206-207 (first channel)
        const size_t ch = (size_t)g.channels;
        for (size_t i = 0; i < nfl; i += ch) { const float v = g.values[i]; mx = fmaxf(mx, v); mn = fminf(fmaxf(FLT_EPSILON, v), mn); }
        stats[0] = mx; stats[1] = mn; stats[2] = (float)g.voxel_size[0]; stats[3] = g.background[0];
    }
    return VPT_OK;
}

int vpt_hdr_load(const char* path, float** rgba_out, unsigned* w, unsigned* h) {
    if (!path || !rgba_out || !w || !h) return fail(nullptr, VPT_ERR_INVALID, "vpt_hdr_load: null argument");
    std::vector<float> px; std::string err;
    if (!vpt::load_hdr_float4(path, px, *w, *h, err)) return fail(nullptr, VPT_ERR_IO, "vpt_hdr_load: " + err);
    *rgba_out = (float*)malloc(px.size() * sizeof(float)); memcpy(*rgba_out, px.data(), px.size() * sizeof(float));
    return VPT_OK;
}

int vpt_bmp_load_rbg(const char* path, float** xyz_out, int* w, int* h) {
    if (!path || !xyz_out || !w || !h) return fail(nullptr, VPT_ERR_INVALID, "vpt_bmp_load_rbg: null argument");
    std::vector<float> px; std::string err;
    if (!vpt::load_bmp_float3_rbg(path, px, *w, *h, err)) return fail(nullptr, VPT_ERR_IO, "vpt_bmp_load_rbg: " + err);
    *xyz_out = (float*)malloc(px.size() * sizeof(float)); memcpy(*xyz_out, px.data(), px.size() * sizeof(float));
    return VPT_OK;
}

int vpt_exr_load_rgb(const char* path, float** rgb_out, int* w, int* h) {
    if (!path || !rgb_out || !w || !h) return fail(nullptr, VPT_ERR_INVALID, "vpt_exr_load_rgb: null argument");
    std::vector<float> px; std::string err;
    if (!vpt::load_exr_float3(path, px, *w, *h, err)) return fail(nullptr, VPT_ERR_IO, "vpt_exr_load_rgb: " + err);
    *rgb_out = (float*)malloc(px.size() * sizeof(float)); memcpy(*rgb_out, px.data(), px.size() * sizeof(float));
    return VPT_OK;
}

void vpt_camera_look_at(vpt_camera* cam, const float lookfrom[3], const float lookat[3], const float vup[3], float vfov, float aspect, float aperture) {
    auto sub = [](const float* a, const float* b, float* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; };
    auto len = [](const float* a) { return sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
    auto norm = [&](float* a) { const float inv = 1.0f / sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); a[0] *= inv; a[1] *= inv; a[2] *= inv; };
    auto cross = [](const float* a, const float* b, float* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
    memset(cam, 0, sizeof(*cam));
    cam->time0 = .0f; cam->time1 = 1.0f;
    float d[3]; sub(lookfrom, lookat, d);
    const float focus = len(d);
    cam->focus_dist = focus;
    cam->lens_radius = aperture / 2.0f;
    const float theta = vfov * float(M_PI) / 180.0f;
    const float half_height = (float)tan(theta / 2.0f);
    const float half_width = aspect * half_height;
    float w[3] = { d[0], d[1], d[2] }; norm(w);
    float u[3]; cross(vup, w, u); norm(u);
    float v[3]; cross(w, u, v);
    for (int a = 0; a < 3; ++a) {
        (&cam->origin.x)[a] = lookfrom[a];
        (&cam->u.x)[a] = u[a]; (&cam->v.x)[a] = v[a]; (&cam->w.x)[a] = w[a];
        (&cam->lower_left_corner.x)[a] = lookfrom[a] - half_width * focus * u[a] - half_height * focus * v[a] - focus * w[a];
        (&cam->horizontal.x)[a] = 2.0f * half_width * focus * u[a];
        (&cam->vertical.x)[a] = 2.0f * half_height * focus * v[a];
    }
    cam->viz_dof = 0;
}

void vpt_kernel_params_defaults(vpt_kernel_params* kp) {
    memset(kp, 0, sizeof(*kp));
    kp->render = 1; kp->iteration = 0; kp->max_interactions = 100; kp->exposure_scale = 1.0f;
    kp->environment_type = 0; kp->ray_depth = 50; kp->volume_depth = 1;
    kp->phase_g1 = 0.0f; kp->phase_g2 = 0.0f; kp->phase_f = 1.0f; kp->tr_depth = 1.0f; kp->density_mult = 1.0f;
    kp->albedo = {1.0f, 1.0f, 1.0f}; kp->extinction = {1.0f, 1.0f, 1.0f};
    kp->azimuth = 120.0f; kp->elevation = 30.0f;            // ImGui values that overwrite 150/30 every frame (main.cpp:1420-1421, 1538-1539)
    kp->sun_color = {1.0f, 1.0f, 1.0f}; kp->sun_mult = 1.0f;
    kp->energy_inject = 1.0;                                 // main.cpp:1543 (energy == 0)
    kp->sky_color = {1.0f, 1.0f, 1.0f}; kp->sky_mult = 1.0f;
    kp->env_sample_tex_res = 360; kp->integrator = 0; kp->emission_scale = 0.0f; kp->emission_pivot = 1.0f;
}

} // extern "C"
