// vpt_kernels.h -- launch interface between the host context (csrc/host) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../../include/vpt_abi.h"
#include "vpt_scene.cuh"

namespace vpt {

constexpr int kTraceThreads = 128;
#ifndef VPT_TRACE_MIN_CTAS
#define VPT_TRACE_MIN_CTAS 4
#endif
constexpr int kTraceMinCtas = VPT_TRACE_MIN_CTAS;     // __launch_bounds__ hint: 4 CTAs x (51 KB ray slots + 4 KB tables) and <= 128 registers per thread

// Which pixels this rank renders: rows are dealt to ranks in interleaved stripes of `stripe_h` rows
// (rank r owns stripes r, r+R, r+2R, ...).  One rank: identity mapping, local == global indices.
struct FrameGeom {
    int width, height;      // full frame
    int local_rows;         // rows stored locally (padded so every rank holds the same count)
    int n_local;            // local_rows * width
    int stripe_h, n_ranks, rank;
};

// Everything a per-frame kernel needs, passed by value (about 600 B of kernel parameter space).
struct FrameArgs {
    vpt_camera        cam;
    vpt_light_list    lights;
    vpt_kernel_params kp;          // buffers inside are indexed by LOCAL pixel
    const vpt_sphere* sphere;
    const SceneTables* scene;
    FrameGeom         geom;
    // ray queue (hits), SoA: direction + entry distance | (local pixel, pass | draws<<6 | obj<<16) | origin (thin lens only)
    float4*   queue_dir;
    uint2*    queue_id;
    float4*   queue_aux;           // thin lens: per-ray origin; pre-stepped ray (pinhole only): position where stepping starts
    int       thin_lens;           // cam.lens_radius != 0
    const float2* bn_table;        // [pass][65536] blue-noise jitter of the chunk
    int       n_passes;            // passes in this chunk (k_generate loops over them)
    int       passes_per_block;    // k_generate: passes handled by one block (blockIdx.z selects the run)
    int       debug_flags;         // development switches (0 in production)
    int       sched_min_lanes;     // trace scheduler: lanes an operation must gather before it pre-empts stepping
    unsigned* queue_count;
    unsigned* queue_head;
    // sample planes, [pass][local pixel]
    float4 *planeA, *planeB, *planeC, *planeD;   // planeD (env_pos) only for environment_type == 0, else null
    // optional statistics: [0] volume lookups, [1] lane-steps, [2] warp step iterations, [3] lane transitions, [4] warp transition rounds
    unsigned long long* counters;
    // cell table of volume 0 (vpt_cells_create; null = texture path): per texel cell its eight corner values, 32 bytes = ONE sector per look-up
    const float4* cell_table;
    int       cell_nx, cell_ny, cell_nz;
    int       tiles_per_block;     // k_generate: consecutive 32x4 tiles handled by one block (blockIdx.x selects the run)
};

cudaError_t launch_prepare_scene(const vpt_gpu_vdb* vols, const vpt_octnode* root, SceneTables* out, OctInternal* internal,
                                 uint2* leaf_list, int* leaf_indices, VolumeRec* vrec, int max_volumes, cudaStream_t s);
cudaError_t launch_prepare_volumes(const vpt_gpu_vdb* vols, const SceneTables& hdr, SceneTables* out, VolumeRec* vrec, cudaStream_t s);
cudaError_t launch_generate(const FrameArgs& fa, int n_passes, cudaStream_t s);
// atm != null selects the volumetric path integrator variant of the trace kernel
cudaError_t launch_trace(const FrameArgs& fa, const vpt_atmosphere* atm, bool lean, int slots, int n_ctas, cudaStream_t s);   // slots: rays per lane, 2 or 3; fa.cell_table != null: the lean 2-slot cell-table instantiation
// per device: dynamic shared memory opt-in of the trace kernels + CTAs per SM of [generic, lean, volumetric path, brick]
cudaError_t trace_kernels_init(int max_ctas[7]);      // [6]: lean, 2 rays per lane, cell table      // [0..2] generic / lean / volumetric path at 3 rays per lane, [3] k_trace_brick, [4..5] generic / lean at 2 rays per lane
// fast mode: lean direct integrator reading the density from a brick pool (vpt_trace_brick.cuh); dims = voxels per axis
cudaError_t launch_trace_brick(const FrameArgs& fa, const float* pool, const int dims[3], int n_ctas, cudaStream_t s);
cudaError_t launch_sampler_compare(unsigned long long tex, const float* pool, const int dims[3], int n, unsigned seed, double* d_out12, cudaStream_t s);
cudaError_t launch_fill_perlin(float* d_buffer, int dx, int dy, int dz, float scale, int seed, cudaStream_t s);
cudaError_t launch_build_bricks(const float* d_dense, int dx, int dy, int dz, float* d_bricks, cudaStream_t s);
cudaError_t launch_build_cells(const float* d_dense, int dx, int dy, int dz, float4* d_cells, cudaStream_t s);
// sky != null selects the environment_type == 0 variant (host copy of the caller's AtmosphereParameters)
// Peer-memory exchange (multi-GPU, csrc/host/vpt_comm.cpp): the call's LAST resolve kernel stores every finished pixel straight into the
// full-frame buffer of every rank (its own included) at the pixel's global position, over NVLink peer mappings -- the all-gather and
// the stripe un-permutation fused into the kernel that produces the values.
constexpr int kMaxPeers = 8;
struct PeerFrames { int n = 0; float* accum[kMaxPeers] = {}; unsigned int* display[kMaxPeers] = {}; };
cudaError_t launch_resolve(const FrameArgs& fa, const vpt_atmosphere* sky, int n_passes, int sampled, int write_display, const PeerFrames* peers, cudaStream_t s);
// flags: 64-bit epoch counters in peer-mapped memory.  signal: store `epoch` into slot `rank` of flag array `which` (0 ready, 1 done) of every
// peer; wait: spin until slots 0..n-1 of the LOCAL array have all reached `epoch` (gives up after ~10 s and raises the block's error word).
cudaError_t launch_peer_signal(unsigned long long* const peer_flags[kMaxPeers], int n, int rank, int which, unsigned long long epoch, cudaStream_t s);
cudaError_t launch_peer_wait(unsigned long long* local_flags, int n, int which, unsigned long long epoch, cudaStream_t s);
constexpr int kPeerFlagStride = 16;                   // u64 slots per flag array; layout of a block header: ready[16] done[16] error[1]
// limit = min(W*H, 65536) entries are advanced (the reference updates entry y*W + x from pixel (x, y))
cudaError_t launch_bn_prepare(void* bn, float2* table, int np, int limit, unsigned* queue_counters, cudaStream_t s);   // also zeroes queue_counters[0..1]
cudaError_t launch_bn_advance(void* bn, int n, int limit, cudaStream_t s);
cudaError_t launch_unpermute(const void* gathered, void* full, const FrameGeom& g, int elem_bytes, cudaStream_t s);

} // namespace vpt
