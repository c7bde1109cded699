/*
 * ref_harness.cu -- TEST INFRASTRUCTURE ONLY (oracle side).
 *
 * Thin C-ABI shim that drives the *reference's own* device code, compiled unmodified from
 * /root/reference at build time (see oracle/Makefile):
 *   - vptref_load_kernel / vptref_launch : the loader + launch of source/main.cpp:1215-1236 and
 *     :1823-1829 (cuModuleLoad -> "volume_rt_kernel", grid (w/16+1, h/16+1) x (16,16), sync).
 *   - vptref_build_octree : the host root set-up of source/bvh/bvh_builder.cpp:61-96 followed by
 *     the reference `build_octree` (source/bvh/bvh_kernels.cu:582-604, device-heap recursion).
 *   - vptref_build_bvh    : the reference LBVH (`BuildBVH`, bvh_kernels.cu:460-580), N >= 2 only
 *     (for N == 1 the reference dereferences an unwritten parent pointer).
 *   - vptref_fill_volume  : the reference's fill_volume_buffer (texture_kernels.cu, compiled unmodified into
 *     texture_kernels_ref.cubin) with the launch shape of gpu_vdb.cpp:547-552.
 *   - vptref_bn_advance   : launches `bn_advance_ref`, a kernel generated at build time from the
 *     reference's own blue-noise update statements (render_kernel.cu:2321-2324), used together
 *     with the "nobn" oracle build that has that block compiled out (race-free protocol, Q6).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs may load this.
 * The product library (libvpt_b200.so) never links or calls it.
 */
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define _USE_MATH_DEFINES
#include <cmath>
#include <curand_kernel.h>
#include "helper_math.h"
#include "kernel_params.h"
#include "atmosphere/definitions.h"
#include "gpu_vdb.h"
#include "camera.h"
#include "light.h"
#include "bvh/bvh.h"
#include "geometry/geometry.h"

// The three GPU_VDB special members live in gpu_vdb.cpp, which needs OpenVDB and cannot be built
// here; they are trivial there (gpu_vdb.cpp:53-59), so the harness supplies equivalents to link.
GPU_VDB::GPU_VDB() {}
GPU_VDB::GPU_VDB(const GPU_VDB& o) : vdb_info(o.vdb_info), xform(o.xform) {}
GPU_VDB::~GPU_VDB() {}

extern "C" void BuildBVH(BVH& bvh, GPU_VDB* volumes, int numVolumes, AABB& sceneBounds, bool debug_bvh);
extern "C" void build_octree(OCTNode* root, GPU_VDB* volumes, int num_volumes, int depth, bool debug);

static CUmodule g_mod[3] = {nullptr, nullptr, nullptr};
static CUfunction g_fn[3] = {nullptr, nullptr, nullptr};   // [0] unmodified kernel, [1] "nobn" build, [2] a candidate drop-in module (level (A) test)
static CUmodule g_bn_mod = nullptr;
static CUfunction g_bn_fn = nullptr;

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); \
    fprintf(stderr, "[vptref] %s failed: %s\n", #x, s_ ? s_ : "?"); return (int)r_; } } while (0)
#define CKR(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { \
    fprintf(stderr, "[vptref] %s failed: %s\n", #x, cudaGetErrorString(r_)); return (int)r_; } } while (0)

extern "C" {

int vptref_sizes(size_t* out, int n) {
    size_t s[] = { sizeof(camera), sizeof(light_list), sizeof(GPU_VDB), sizeof(sphere), sizeof(geometry_list),
                   sizeof(BVHNode), sizeof(OCTNode), sizeof(AtmosphereParameters), sizeof(Kernel_params),
                   sizeof(point_light), sizeof(VDB_INFO), sizeof(AABB) };
    int m = (int)(sizeof(s) / sizeof(s[0]));
    for (int i = 0; i < n && i < m; ++i) out[i] = s[i];
    return m;
}

// which: 0 = unmodified reference kernel, 1 = build with the blue-noise tail compiled out, 2 = any other module that claims to be
// a drop-in: it goes through the very same loader calls and the very same launch below.
int vptref_load_kernel(const char* cubin_path, int which) {
    CKR(cudaFree(0));
    if (which < 0 || which > 2) return -1;
    if (g_mod[which]) { cuModuleUnload(g_mod[which]); g_mod[which] = nullptr; g_fn[which] = nullptr; }
    CK(cuModuleLoad(&g_mod[which], cubin_path));
    CK(cuModuleGetFunction(&g_fn[which], g_mod[which], "volume_rt_kernel"));
    return 0;
}

int vptref_load_bn_kernel(const char* cubin_path) {
    CKR(cudaFree(0));
    CK(cuModuleLoad(&g_bn_mod, cubin_path));
    CK(cuModuleGetFunction(&g_bn_fn, g_bn_mod, "bn_advance_ref"));
    return 0;
}

// The reference's procedural density fill (source/texture_kernels.cu:76-128) launched as GPU_PROC_VOL::create_volume does
// (gpu_vdb.cpp:547-552: block 8x8x8, grid dim/8 + 1).  Its sub-voxel jitter comes from an UNINITIALISED curand state (quirk Q14).
static CUmodule g_tex_mod = nullptr;
static CUfunction g_fill_fn = nullptr;
int vptref_fill_volume(const char* cubin_path, void* d_buffer, int dx, int dy, int dz, float scale, int noise_type) {
    CKR(cudaFree(0));
    if (!g_fill_fn) { CK(cuModuleLoad(&g_tex_mod, cubin_path)); CK(cuModuleGetFunction(&g_fill_fn, g_tex_mod, "fill_volume_buffer")); }
    int3 dims = make_int3(dx, dy, dz);
    void* params[] = { &d_buffer, &dims, &scale, &noise_type };
    CK(cuLaunchKernel(g_fill_fn, dx / 8 + 1, dy / 8 + 1, dz / 8 + 1, 8, 8, 8, 0, NULL, params, NULL));
    CKR(cudaDeviceSynchronize());
    return 0;
}

// One progressive pass exactly as the reference frame loop issues it (main.cpp:1823-1829).
int vptref_launch(void** params, unsigned width, unsigned height, int which, int sync) {
    if (!g_fn[which]) return -2;
    unsigned bx = 16, by = 16;
    unsigned gx = (unsigned)(int(width / bx) + 1), gy = (unsigned)(int(height / by) + 1);
    CK(cuLaunchKernel(g_fn[which], gx, gy, 1, bx, by, 1, 0, NULL, params, NULL));
    if (sync) CKR(cudaDeviceSynchronize());
    return 0;
}

// n_entries: how many blue-noise entries the pass advances.  The reference updates entry idx = y*W + x from the thread of pixel
// (x, y) when idx < 65536, i.e. min(W*H, 65536) entries; the generated kernel keeps only the `idx < 65536` guard, so it is launched
// with exactly that many single-thread blocks.
int vptref_bn_advance(const void* kernel_params, int sync, unsigned n_entries) {
    if (!g_bn_fn) return -2;
    void* params[] = { (void*)kernel_params };
    if (n_entries > 65536u) n_entries = 65536u;
    if (n_entries == 0) return 0;
    CK(cuLaunchKernel(g_bn_fn, n_entries, 1, 1, 1, 1, 1, 0, NULL, params, NULL));
    if (sync) CKR(cudaDeviceSynchronize());
    return 0;
}

// h_volumes: host array of n GPU_VDB (144 B each).  Returns the device root in *d_root_out.
int vptref_build_octree(const void* h_volumes, int n, void** d_root_out) {
    const GPU_VDB* vdbs = reinterpret_cast<const GPU_VDB*>(h_volumes);
    if (n < 1 || n > 600) { fprintf(stderr, "[vptref] octree: n=%d outside the reference's 1..600 range\n", n); return -3; }
    // every reference octree takes 585 x 2520 B from the device heap and is never freed by the reference;
    // the default 8 MB heap is exhausted after five builds (then `new` returns NULL and the build kernel faults)
    static bool heap_set = false;
    if (!heap_set) { cudaDeviceSetLimit(cudaLimitMallocHeapSize, (size_t)512 << 20); cudaGetLastError(); heap_set = true; }
    GPU_VDB* d_vols = nullptr;
    CKR(cudaMalloc(&d_vols, n * sizeof(GPU_VDB)));
    CKR(cudaMemcpy(d_vols, vdbs, n * sizeof(GPU_VDB), cudaMemcpyHostToDevice));

    OCTNode* root_h = new OCTNode;
    root_h->depth = 4;
    for (int i = 0; i < n; ++i) {
        AABB b = vdbs[i].Bounds();
        root_h->bbox.pmax = fmaxf(root_h->bbox.pmax, b.pmax);
        root_h->bbox.pmin = fminf(root_h->bbox.pmin, b.pmin);
        root_h->vol_indices[i] = i;
        root_h->num_volumes++;
        root_h->max_extinction = fmaxf(root_h->max_extinction, vdbs[i].vdb_info.max_density);
        root_h->min_extinction = fminf(root_h->min_extinction, vdbs[i].vdb_info.min_density);
        root_h->has_children = true;
    }
    root_h->bbox.pmax += make_float3(1.0f);
    root_h->bbox.pmin -= make_float3(1.0f);

    OCTNode* d_root = nullptr;
    CKR(cudaMalloc(&d_root, sizeof(OCTNode)));
    CKR(cudaMemcpy(d_root, root_h, sizeof(OCTNode), cudaMemcpyHostToDevice));
    build_octree(d_root, d_vols, n, root_h->depth - 1, false);
    CKR(cudaDeviceSynchronize());
    delete root_h;
    cudaFree(d_vols);
    *d_root_out = d_root;
    return 0;
}

// Reference LBVH over n >= 2 instances; returns device arrays (internal nodes, leaves).
int vptref_build_bvh(const void* h_volumes, int n, void** d_nodes_out, void** d_leaves_out, float* scene_bounds6) {
    if (n < 2) return -3;
    const GPU_VDB* vdbs = reinterpret_cast<const GPU_VDB*>(h_volumes);
    GPU_VDB* d_vols = nullptr;
    CKR(cudaMalloc(&d_vols, n * sizeof(GPU_VDB)));
    CKR(cudaMemcpy(d_vols, vdbs, n * sizeof(GPU_VDB), cudaMemcpyHostToDevice));
    BVH bvh;
    AABB sb(make_float3(.0f), make_float3(.0f));
    BuildBVH(bvh, d_vols, n, sb, false);
    CKR(cudaDeviceSynchronize());
    *d_nodes_out = bvh.BVHNodes;
    *d_leaves_out = bvh.BVHLeaves;
    if (scene_bounds6) { scene_bounds6[0] = sb.pmin.x; scene_bounds6[1] = sb.pmin.y; scene_bounds6[2] = sb.pmin.z;
                         scene_bounds6[3] = sb.pmax.x; scene_bounds6[4] = sb.pmax.y; scene_bounds6[5] = sb.pmax.z; }
    cudaFree(d_vols);
    return 0;
}

// Host-side reference math the harness *uses* rather than restates.
void vptref_update_camera(void* cam_out, const float* lookfrom, const float* lookat, const float* vup,
                          float vfov, float aspect, float aperture) {
    camera c;
    c.update_camera(make_float3(lookfrom[0], lookfrom[1], lookfrom[2]), make_float3(lookat[0], lookat[1], lookat[2]),
                    make_float3(vup[0], vup[1], vup[2]), vfov, aspect, aperture);
    c.viz_dof = false;
    memcpy(cam_out, &c, sizeof(camera));
}

void vptref_bounds(const void* h_volume, float* out6) {
    AABB b = reinterpret_cast<const GPU_VDB*>(h_volume)->Bounds();
    out6[0] = b.pmin.x; out6[1] = b.pmin.y; out6[2] = b.pmin.z; out6[3] = b.pmax.x; out6[4] = b.pmax.y; out6[5] = b.pmax.z;
}

// Instance transform of main.cpp:1066-1097 evaluated with the reference's own mat4 algebra.
void vptref_instance_xform(const float* base_xform16, const float* pos3, const float* quat4, float scale, float* out16) {
    mat4 xform;
    memcpy(&xform, base_xform16, sizeof(mat4));
    // (restated call sequence; the arithmetic is the reference's mat4/quaternion code)
    mat4 rot = quaternion_to_mat4(quat4[0], quat4[1], quat4[2], quat4[3]);
    mat4 x0 = xform;
    x0[0][3] = 0.0f; x0[1][3] = 0.0f; x0[2][3] = 0.0f;
    x0.scale(make_float3(scale));
    mat4 r = rot * x0;
    r.translate(make_float3(pos3[0], pos3[1], pos3[2]));
    memcpy(out16, &r, sizeof(mat4));
}

} // extern "C"
