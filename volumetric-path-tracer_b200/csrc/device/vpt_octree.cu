// vpt_octree.cu -- parallel build of the fixed depth-3 instance octree, reference node layout.
//
// Replaces the reference's serial device-heap recursion `pass_octree<<<1,1>>>`
// (source/bvh/bvh_kernels.cu:204-246, 455, 582-604) and the host root set-up of
// source/bvh/bvh_builder.cpp:61-96: one thread per node of a contiguous 585-node array
// (1 + 8 + 64 + 512), every node deriving its box from the root by the same three halvings the
// recursion performs, then collecting the instances whose AABB overlaps it in ascending order.
// The output is a pointer-linked vpt_octnode tree, so either kernel (this library's or the
// reference's) can consume it.  Compiled WITHOUT --use_fast_math, like the reference's bvh object
// (source/CompileObj.cmake:23), so instance bounds round the same way.
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../../include/vpt_abi.h"

namespace vpt {

constexpr int kOctNodes = 585;

struct Box { float3 pmin, pmax; };

__host__ __device__ inline Box instance_bounds(const vpt_gpu_vdb& g)      // GPU_VDB::Bounds, gpu_vdb.h:131-146
{
    const float3 bmax = make_float3(g.vdb_info.bmax.x, g.vdb_info.bmax.y, g.vdb_info.bmax.z);
    const float3 bmin = make_float3(g.vdb_info.bmin.x, g.vdb_info.bmin.y, g.vdb_info.bmin.z);
    const float3 c = make_float3((bmax.x + bmin.x) * 0.5f, (bmax.y + bmin.y) * 0.5f, (bmax.z + bmin.z) * 0.5f);
    const float3 e = make_float3((bmax.x - bmin.x) * 0.5f, (bmax.y - bmin.y) * 0.5f, (bmax.z - bmin.z) * 0.5f);
    const float (*X)[4] = g.xform;
    float3 nc, ne;
    nc.x = X[0][0] * c.x + X[0][1] * c.y + X[0][2] * c.z + X[0][3] * 1.0f;
    nc.y = X[1][0] * c.x + X[1][1] * c.y + X[1][2] * c.z + X[1][3] * 1.0f;
    nc.z = X[2][0] * c.x + X[2][1] * c.y + X[2][2] * c.z + X[2][3] * 1.0f;
    ne.x = fabsf(X[0][0]) * e.x + fabsf(X[0][1]) * e.y + fabsf(X[0][2]) * e.z + fabsf(X[0][3]) * 0.0f;
    ne.y = fabsf(X[1][0]) * e.x + fabsf(X[1][1]) * e.y + fabsf(X[1][2]) * e.z + fabsf(X[1][3]) * 0.0f;
    ne.z = fabsf(X[2][0]) * e.x + fabsf(X[2][1]) * e.y + fabsf(X[2][2]) * e.z + fabsf(X[2][3]) * 0.0f;
    Box b;
    b.pmin = make_float3(nc.x - ne.x, nc.y - ne.y, nc.z - ne.z);
    b.pmax = make_float3(nc.x + ne.x, nc.y + ne.y, nc.z + ne.z);
    return b;
}

__device__ inline Box child_box(int idx, float3 pmin, float3 pmax)        // divide_bbox, bvh_kernels.cu:150-202
{
    const float hx = (pmin.x + pmax.x) * 0.5, hy = (pmin.y + pmax.y) * 0.5, hz = (pmin.z + pmax.z) * 0.5;
    const bool xp = idx & 1, ym = idx & 2, zp = idx & 4;      // 0:(x-,y+,z-) 1:(x+,y+,z-) 2:(x-,y-,z-) 3:(x+,y-,z-) 4..7: z+
    Box b;
    b.pmin = make_float3(xp ? hx : pmin.x, ym ? pmin.y : hy, zp ? hz : pmin.z);
    b.pmax = make_float3(xp ? pmax.x : hx, ym ? hy : pmax.y, zp ? pmax.z : hz);
    return b;
}

__device__ inline bool overlaps(const Box& a, const Box& b)                // Overlaps, AABB.h:135-140
{
    const bool x = (a.pmax.x >= b.pmin.x) && (a.pmin.x <= b.pmax.x);
    const bool y = (a.pmax.y >= b.pmin.y) && (a.pmin.y <= b.pmax.y);
    const bool z = (a.pmax.z >= b.pmin.z) && (a.pmin.z <= b.pmax.z);
    return x && y && z;
}

__global__ void k_instance_bounds(const vpt_gpu_vdb* __restrict__ vols, int n, Box* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = instance_bounds(vols[i]);
}

// nodes[0] (the root) is filled by the host; this kernel fills nodes 1..584 and the child pointers.
__global__ void k_octree_build(vpt_octnode* nodes, const vpt_gpu_vdb* __restrict__ vols, const Box* __restrict__ bounds, int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= kOctNodes) return;
    vpt_octnode* nd = &nodes[j];
    const float3 rmin = make_float3(nodes[0].bbox.pmin.x, nodes[0].bbox.pmin.y, nodes[0].bbox.pmin.z);
    const float3 rmax = make_float3(nodes[0].bbox.pmax.x, nodes[0].bbox.pmax.y, nodes[0].bbox.pmax.z);

    // path from the root
    int level, c[3] = {0, 0, 0};
    if (j == 0) level = 0;
    else if (j < 9) { level = 1; c[0] = j - 1; }
    else if (j < 73) { level = 2; c[0] = (j - 9) >> 3; c[1] = (j - 9) & 7; }
    else { level = 3; c[0] = (j - 73) >> 6; c[1] = ((j - 73) >> 3) & 7; c[2] = (j - 73) & 7; }

    // ancestors' boxes and whether each ancestor holds any volume (children exist only under those)
    Box b; b.pmin = rmin; b.pmax = rmax;
    bool exists = true;
    for (int l = 0; l < level; ++l) {
        if (l > 0) {
            bool any = false;
            for (int v = 0; v < n && !any; ++v) any = overlaps(b, bounds[v]);
            if (!any) { exists = false; break; }
        }
        b = child_box(c[l], b.pmin, b.pmax);
    }

    if (j == 0) {
        for (int k = 0; k < 8; ++k) nd->children[k] = (vpt_devptr_t)(uintptr_t)&nodes[1 + k];
        return;
    }
    if (!exists) return;                                        // never allocated by the reference; stays zero

    nd->num_volumes = 0;
    nd->max_extinction = .0f;
    nd->min_extinction = 3.402823466e+38F;
    nd->voxel_size = 3.402823466e+38F;
    nd->depth = 4 - level;
    nd->has_children = 0;
    nd->bbox.pmin.x = b.pmin.x; nd->bbox.pmin.y = b.pmin.y; nd->bbox.pmin.z = b.pmin.z;
    nd->bbox.pmax.x = b.pmax.x; nd->bbox.pmax.y = b.pmax.y; nd->bbox.pmax.z = b.pmax.z;
    const int parent = (level == 1) ? 0 : (level == 2) ? 1 + c[0] : 9 + c[0] * 8 + c[1];
    nd->parent = (vpt_devptr_t)(uintptr_t)&nodes[parent];
    int idx = 0;
    for (int v = 0; v < n; ++v) {
        if (overlaps(b, bounds[v])) {
            nd->vol_indices[idx] = v;
            nd->max_extinction = fmaxf(nd->max_extinction, vols[v].vdb_info.max_density);
            nd->min_extinction = fminf(nd->min_extinction, vols[v].vdb_info.min_density);
            nd->voxel_size = fminf(nd->voxel_size, vols[v].vdb_info.voxelsize);
            idx++;
        }
    }
    nd->num_volumes = idx;
    if (idx > 0) {
        nd->has_children = 1;
        if (level < 3) {
            const int first = (level == 1) ? 9 + c[0] * 8 : 73 + c[0] * 64 + c[1] * 8;
            for (int k = 0; k < 8; ++k) nd->children[k] = (vpt_devptr_t)(uintptr_t)&nodes[first + k];
        }
    }
}

cudaError_t octree_build_device(vpt_octnode* d_nodes, const vpt_gpu_vdb* d_vols, int n, cudaStream_t s)
{
    Box* d_bounds = nullptr;
    cudaError_t e = cudaMalloc(&d_bounds, sizeof(Box) * (size_t)n);
    if (e != cudaSuccess) return e;
    k_instance_bounds<<<(n + 127) / 128, 128, 0, s>>>(d_vols, n, d_bounds);
    k_octree_build<<<(kOctNodes + 63) / 64, 64, 0, s>>>(d_nodes, d_vols, d_bounds, n);
    e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(d_bounds);
    return e != cudaSuccess ? e : e2;
}

// Copy a pointer-linked octree (either builder's, including device-heap nodes that cudaMemcpy cannot read)
// into a flat 585-entry array in the canonical numbering; exists[j] = 0 for nodes that were never allocated.
__global__ void k_octree_snapshot(const vpt_octnode* root, vpt_octnode* out, int* exists)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= kOctNodes) return;
    int level, c[3] = {0, 0, 0};
    if (j == 0) level = 0;
    else if (j < 9) { level = 1; c[0] = j - 1; }
    else if (j < 73) { level = 2; c[0] = (j - 9) >> 3; c[1] = (j - 9) & 7; }
    else { level = 3; c[0] = (j - 73) >> 6; c[1] = ((j - 73) >> 3) & 7; c[2] = (j - 73) & 7; }
    const vpt_octnode* n = root;
    bool ok = true;
    for (int l = 0; l < level; ++l) {
        if (n->num_volumes <= 0) { ok = false; break; }
        n = reinterpret_cast<const vpt_octnode*>((uintptr_t)n->children[c[l]]);
    }
    exists[j] = ok ? 1 : 0;
    if (ok) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(n); uint32_t* dst = reinterpret_cast<uint32_t*>(&out[j]);
        for (int w = 0; w < (int)(sizeof(vpt_octnode) / 4); ++w) dst[w] = src[w];
    }
}

cudaError_t octree_snapshot(const vpt_octnode* d_root, vpt_octnode* h_nodes, int* h_exists)
{
    vpt_octnode* d_out = nullptr; int* d_ex = nullptr;
    cudaError_t e = cudaMalloc(&d_out, sizeof(vpt_octnode) * kOctNodes);
    if (e == cudaSuccess) e = cudaMemset(d_out, 0, sizeof(vpt_octnode) * kOctNodes);
    if (e == cudaSuccess) e = cudaMalloc(&d_ex, sizeof(int) * kOctNodes);
    if (e == cudaSuccess) { k_octree_snapshot<<<(kOctNodes + 63) / 64, 64>>>(d_root, d_out, d_ex); e = cudaDeviceSynchronize(); }
    if (e == cudaSuccess) e = cudaMemcpy(h_nodes, d_out, sizeof(vpt_octnode) * kOctNodes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_exists, d_ex, sizeof(int) * kOctNodes, cudaMemcpyDeviceToHost);
    cudaFree(d_out); cudaFree(d_ex);
    return e;
}

void instance_bounds_host(const vpt_gpu_vdb& g, float out6[6])
{
    Box b = instance_bounds(g);
    out6[0] = b.pmin.x; out6[1] = b.pmin.y; out6[2] = b.pmin.z; out6[3] = b.pmax.x; out6[4] = b.pmax.y; out6[5] = b.pmax.z;
}

} // namespace vpt
