// vpt_host.h -- declarations shared by the host translation units of libvpt_b200.so (not part of the public ABI).
#pragma once
#include "../../../include/vpt_b200.h"
#include "../device/vpt_kernels.h"

#include <cuda_runtime.h>
#include <string>

#include <vector>

struct vpt_context {
    int device = 0;
    int num_sms = 0;
    std::string err;
    // options
    int passes_per_chunk = 32;
    int ctas_per_sm = 0;
    // partition
    int rank = 0, n_ranks = 1, stripe_rows = 16;
    // scene cache: keyed by the two device pointers AND the registry generation of the octree (a rebuilt octree that got the
    // same address back is a different scene); the cheap per-volume records are refreshed whenever a new frame starts
    vpt_devptr_t cached_volumes = 0, cached_root = 0;
    unsigned long long cached_generation = 0;
    bool   scene_single_volume = false;
    size_t cap_vrec = 0;              // VolumeRec capacity (grows with the instance count)
    int*   h_pinned = nullptr;        // one pinned word for the 4-byte read-back a foreign octree needs
    int    max_ctas[7] = {0, 0, 0, 0, 0, 0, 0};   // trace kernel occupancy: [0] generic, [1] lean, [2] volumetric path, [3] brick mode, [4] generic / [5] lean at 2 rays per lane, [6] lean + cell table
    size_t l2_fetch_default = 0; int l2_sector_fetch = 0;     // option "l2_sector_fetch" (cell mode)
    const float4* cell_table = nullptr; int cell_dims[3] = {0, 0, 0};   // vpt_set_cell_volume
    int    trace_slots = 0;              // option "trace_slots": rays per lane of k_trace, 0 = by grid size (2 when the grid is larger than twice the L2)
    size_t l2_bytes = 0;
    const float* brick_pool = nullptr; int brick_dims[3] = {0, 0, 0};   // fast mode: density of volume 0 as a brick pool (vpt_set_brick_volume)
    int    force_generic = 0;        // option "generic_kernel": 1 = never use the lean trace instantiation (A/B and tests)
    vpt::SceneTables* d_scene = nullptr;
    vpt::OctInternal* d_internal = nullptr;
    uint2* d_leaf_list = nullptr;
    int* d_leaf_indices = nullptr;
    vpt::VolumeRec* d_vrec = nullptr;
    // frame buffers
    size_t cap_samples = 0;           // n_local * chunk
    bool   cap_planeD = false;
    uint2* d_queue_id = nullptr; float4* d_queue_org = nullptr; float2* d_bn_table = nullptr; int cap_chunk = 0;
    int sched_min_lanes = 0;          // option "sched_min_lanes": 0 = by kernel (26 for the lean kernel at 3 rays per lane, else 20)
    int debug_flags = 0;
    float4 *d_queue = nullptr, *d_planeA = nullptr, *d_planeB = nullptr, *d_planeC = nullptr, *d_planeD = nullptr;
    unsigned* d_counters = nullptr;   // [0] queue_count, [1] queue_head
    // stats
    unsigned long long launches = 0;
    int count_stats = 0;              // option "count_stats": accumulate trace counters
    unsigned long long* d_stats = nullptr;   // 8 counters
    int profile = 0;                  // option "profile": CUDA-event timing of every kernel (debug / bench breakdown)
    struct Ev { int kind; cudaEvent_t a, b; };
    std::vector<Ev> events;
    // multi-GPU (vpt_comm.cpp): NCCL communicator + gather targets; null for a single rank
    void*  nccl_comm = nullptr;
    void*  d_gathered = nullptr; size_t cap_gathered = 0;      // [rank][local pixel] staging of the all-gather
    void*  d_full_accum = nullptr;                              // caller-owned full frame float3 [H][W] (gather target), may be null
    void*  d_full_display = nullptr;                            // caller-owned full frame u32 [H][W], may be null
    void*  d_gathered_disp = nullptr; size_t cap_gathered_disp = 0;
    cudaStream_t comm_stream = nullptr; cudaEvent_t ev_render_done = nullptr, ev_gather_done = nullptr;
    // peer-memory exchange (vpt_comm_p2p_*): one cudaMalloc'd block per rank {flags, full accum, full display}, IPC-mapped into every rank
    void*  p2p_block = nullptr; size_t p2p_bytes = 0; int p2p_w = 0, p2p_h = 0;
    void*  p2p_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [rank] -> that rank's block as mapped here (own: p2p_block)
    int    p2p_n = 0; bool p2p_on = false, p2p_display = false, p2p_local = false; unsigned long long p2p_epoch = 0;
    int    gather_async = 0;                                    // option "gather_async": gather on the side stream (see vpt_comm_wait)
    bool   gather_pending = false;
    size_t max_scratch_bytes = (size_t)12 << 30;                // cap of the per-chunk ray queue + sample planes
    int    chunk_auto = 1;                                      // passes_per_chunk not set by the caller: pick by local frame size
};


namespace vpt {

// ---- device-side builders (vpt_octree.cu) ---------------------------------------------------------------
cudaError_t octree_build_device(vpt_octnode* d_nodes, const vpt_gpu_vdb* d_vols, int n, cudaStream_t s);
void        instance_bounds_host(const vpt_gpu_vdb& g, float out6[6]);
cudaError_t octree_snapshot(const vpt_octnode* d_root, vpt_octnode* h_nodes, int* h_exists);
cudaError_t octree_flat_counts(const vpt_gpu_vdb* d_vols, int n, const float root6[6], void** d_bounds_out, int h_counts[585], cudaStream_t s);
cudaError_t octree_flat_fill(const void* d_bounds, int n, const float root6[6], const unsigned* d_offsets, int* d_indices, cudaStream_t s);
cudaError_t bvh_build_device(const vpt_gpu_vdb* d_vols, int n, vpt_bvhnode* d_nodes, vpt_bvhnode* d_leaves, float scene6[6],
                             unsigned long long* h_codes, int* h_ids, cudaStream_t s);

// ---- scene registry (vpt_scene_build.cpp) ------------------------------------------------------------------
// Every octree built by vpt_octree_build is registered under its root device pointer together with the flat
// tables the render kernels read.  vpt_render_passes looks the root pointer of params[VPT_ARG_OCTREE] up here;
// a pointer that is not registered is a foreign (reference-built, pointer-linked) octree and is flattened by
// k_prepare_scene instead.
struct SceneEntry {
    int          device = 0;
    int          n = 0;                      // instances
    unsigned long long generation = 0;       // unique per build: detects a rebuilt octree that got the same address back
    float        root6[6] = {0, 0, 0, 0, 0, 0};
    float        max_extinction = 0.f, min_extinction = 0.f;
    OctInternal* d_internal = nullptr;       // [73]
    uint2*       d_leaf_list = nullptr;      // [512] (offset, count)
    int*         d_leaf_indices = nullptr;   // CSR payload (at least 1 element)
    size_t       total_indices = 0;
    int          max_leaf_count = 0;
    unsigned     any_flags = 0;              // bit0: some instance has a colour grid, bit1: some instance has an emission grid
    unsigned long long max_grid_bytes = 0;   // density grid of the largest instance (float voxels): picks the trace kernel's rays per lane
};
bool scene_registry_find(vpt_devptr_t d_root, SceneEntry* out);

int  fail_global(int code, const std::string& msg);     // records the text for vpt_last_error(NULL)
int  fail_ctx(vpt_context* ctx, int code, const std::string& msg);
FrameGeom make_frame_geom(const vpt_context* c, unsigned w, unsigned h);
// after the last resolve of a vpt_render_passes call: all-gather + un-permute into the caller's full-frame buffers (vpt_comm.cpp)
int  comm_p2p_begin(vpt_context* c, cudaStream_t stream);                                        // call start: tell the peers this rank's frame may be overwritten
int  comm_p2p_peers(vpt_context* c, cudaStream_t stream, PeerFrames* out);                       // before the call's last resolve: wait for the peers, hand out their frames
int  comm_p2p_end(vpt_context* c, cudaStream_t stream);                                          // after it: publish, and wait until every peer has published
int  comm_gather_frame(vpt_context* c, const FrameGeom& g, const void* d_local_accum, const void* d_local_display, cudaStream_t stream);
int  comm_before_accum_write(vpt_context* c, cudaStream_t stream);   // async mode: the previous gather must have read accum

} // namespace vpt
