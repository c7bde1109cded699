// vpt_host.h -- declarations shared by the host translation units of libvpt_b200.so (not part of the public ABI).
#pragma once
#include "../../../include/vpt_b200.h"
#include "../device/vpt_kernels.h"

#include <cuda_runtime.h>
#include <string>

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
};
bool scene_registry_find(vpt_devptr_t d_root, SceneEntry* out);

int  fail_global(int code, const std::string& msg);     // records the text for vpt_last_error(NULL)

} // namespace vpt
