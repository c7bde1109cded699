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
#include <vector>
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

// ---------------------------------------------------------------------------------------------------------
// Flat scene tables for ANY number of instances (the reference's OCTNode holds
// 600 indices and its root lists every instance, so > 600 instances overflow the node on the device heap).
// Same node semantics (child boxes by the three halvings, inclusive overlap test, ascending instance order),
// stored as what the render kernels actually read: 73 internal nodes (box + split planes + empty-child mask)
// and 512 leaf lists in CSR form.  One thread per node counts, a host-side scan sizes the index array, one
// thread per leaf fills its list.
// ---------------------------------------------------------------------------------------------------------
__device__ inline void node_path(int j, int& level, int c[3])
{
    c[0] = c[1] = c[2] = 0;
    if (j == 0) level = 0;
    else if (j < 9) { level = 1; c[0] = j - 1; }
    else if (j < 73) { level = 2; c[0] = (j - 9) >> 3; c[1] = (j - 9) & 7; }
    else { level = 3; c[0] = (j - 73) >> 6; c[1] = ((j - 73) >> 3) & 7; c[2] = (j - 73) & 7; }
}

__device__ inline Box node_box(const Box& root, int level, const int c[3])
{
    Box b = root;
    for (int l = 0; l < level; ++l) b = child_box(c[l], b.pmin, b.pmax);
    return b;
}

// counts[j] = number of instances whose AABB overlaps node j (canonical numbering, 585 nodes)
__global__ void k_node_counts(const Box* __restrict__ bounds, int n, Box root, int* __restrict__ counts)
{
    // one warp per node: lanes stride over the instances
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= kOctNodes) return;
    int level, c[3]; node_path(j, level, c);
    const Box b = node_box(root, level, c);
    int cnt = 0;
    for (int v = lane; v < n; v += 32) cnt += overlaps(b, bounds[v]) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) counts[j] = (j == 0) ? n : cnt;
}

// leaf l (0..511) writes its instances, ascending, at offsets[l]; one warp per leaf, order kept by a ballot scan
__global__ void k_leaf_fill(const Box* __restrict__ bounds, int n, Box root, const unsigned* __restrict__ offsets, int* __restrict__ indices)
{
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (l >= 512) return;
    int level, c[3]; node_path(73 + l, level, c);
    const Box b = node_box(root, level, c);
    unsigned at = offsets[l];
    for (int v0 = 0; v0 < n; v0 += 32) {
        const int v = v0 + lane;
        const bool in = v < n && overlaps(b, bounds[v]);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (in) indices[at + __popc(m & ((1u << lane) - 1u))] = v;
        at += __popc(m);
    }
}

cudaError_t octree_flat_counts(const vpt_gpu_vdb* d_vols, int n, const float root6[6], void** d_bounds_out, int h_counts[585], cudaStream_t s)
{
    Box* d_bounds = nullptr; int* d_counts = nullptr;
    cudaError_t e = cudaMalloc(&d_bounds, sizeof(Box) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&d_counts, sizeof(int) * kOctNodes);
    if (e == cudaSuccess) {
        Box root; root.pmin = make_float3(root6[0], root6[1], root6[2]); root.pmax = make_float3(root6[3], root6[4], root6[5]);
        k_instance_bounds<<<(n + 127) / 128, 128, 0, s>>>(d_vols, n, d_bounds);
        k_node_counts<<<(kOctNodes * 32 + 127) / 128, 128, 0, s>>>(d_bounds, n, root, d_counts);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_counts, d_counts, sizeof(int) * kOctNodes, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_counts);
    if (e != cudaSuccess) { cudaFree(d_bounds); d_bounds = nullptr; }
    *d_bounds_out = d_bounds;
    return e;
}

cudaError_t octree_flat_fill(const void* d_bounds, int n, const float root6[6], const unsigned* d_offsets, int* d_indices, cudaStream_t s)
{
    Box root; root.pmin = make_float3(root6[0], root6[1], root6[2]); root.pmax = make_float3(root6[3], root6[4], root6[5]);
    k_leaf_fill<<<(512 * 32 + 127) / 128, 128, 0, s>>>(reinterpret_cast<const Box*>(d_bounds), n, root, d_offsets, d_indices);
    return cudaGetLastError();
}

// child boxes of a node for the host-side assembly of the 73 internal-node records (same halving arithmetic)
void octree_child_halves_host(const float pmin[3], const float pmax[3], float half[3])
{
    half[0] = (pmin[0] + pmax[0]) * 0.5; half[1] = (pmin[1] + pmax[1]) * 0.5; half[2] = (pmin[2] + pmax[2]) * 0.5;
}

// ---------------------------------------------------------------------------------------------------------
// LBVH over the instance AABBs (Karras 2012), restating the reference's BuildBVH (bvh_kernels.cu:460-580):
//   bounds -> scene box (fmin/fmax union) -> 30-bit Morton code of each centroid (:137-148, :320-339)
//   -> stable sort by code (thrust::sort_by_key is a stable radix sort: equal codes keep ascending ids)
//   -> radix tree with the (code, id) tie-break of LongestCommonPrefix (:106-123, :380-453)
//   -> bottom-up refit with one atomic counter per internal node (:341-378).
// Node layout is the reference's BVHNode (64 B, child / parent POINTERS into two arrays: n-1 internal nodes,
// n leaves), so either kernel can take it as `root_node`.  Differences, all on purpose: the arrays are
// zero-initialised (the reference leaves leaf child pointers and the root's parent unwritten: IsLeaf() relies on
// fresh memory being zero), n == 1 is handled (quirk Q18), the refit fences its box writes, and the sort is an
// O(n^2 / threads) stable rank sort (n is the number of instances: thousands).
// ---------------------------------------------------------------------------------------------------------
typedef unsigned long long Morton;

__device__ inline Morton expand_bits(Morton i)                           // bitExpansion, bvh_kernels.cu:126-132
{
    i = (i * 0x00010001u) & 0xFF0000FFu;
    i = (i * 0x00000101u) & 0x0F00F00Fu;
    i = (i * 0x00000011u) & 0xC30C30C3u;
    i = (i * 0x00000005u) & 0x49249249u;
    return i;
}

__global__ void k_bvh_morton(const Box* __restrict__ bounds, int n, Box scene, Morton* __restrict__ codes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Box b = bounds[i];
    const float3 c = make_float3(0.5f * (b.pmax.x + b.pmin.x), 0.5f * (b.pmax.y + b.pmin.y), 0.5f * (b.pmax.z + b.pmin.z));
    float x = (c.x - scene.pmin.x) / (scene.pmax.x - scene.pmin.x);
    float y = (c.y - scene.pmin.y) / (scene.pmax.y - scene.pmin.y);
    float z = (c.z - scene.pmin.z) / (scene.pmax.z - scene.pmin.z);
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    codes[i] = expand_bits((Morton)x) * 4 + expand_bits((Morton)y) * 2 + expand_bits((Morton)z);
}

// stable rank sort: position of i = #{ j : code_j < code_i  or  (code_j == code_i and j < i) }
__global__ void k_bvh_rank_sort(const Morton* __restrict__ codes, int n, Morton* __restrict__ sorted_codes, int* __restrict__ sorted_ids)
{
    __shared__ Morton tile[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const Morton mine = i < n ? codes[i] : 0;
    int rank = 0;
    for (int base = 0; base < n; base += 256) {
        const int j = base + threadIdx.x;
        tile[threadIdx.x] = j < n ? codes[j] : ~0ull;
        __syncthreads();
        const int lim = min(256, n - base);
        for (int k = 0; k < lim; ++k) {
            const Morton o = tile[k];
            rank += (o < mine || (o == mine && base + k < i)) ? 1 : 0;
        }
        __syncthreads();
    }
    if (i < n) { sorted_codes[rank] = mine; sorted_ids[rank] = i; }
}

__device__ inline int bvh_lcp(int i, int j, int n, const Morton* codes, const int* ids)   // LongestCommonPrefix, :106-123
{
    if (i < 0 || i > n - 1 || j < 0 || j > n - 1) return -1;
    const Morton mi = codes[i], mj = codes[j];
    if (mi == mj) return __clzll((long long)(mi ^ mj)) + __clzll((long long)(ids[i] ^ ids[j]));
    return __clzll((long long)(mi ^ mj));
}

__global__ void k_bvh_radix_tree(vpt_bvhnode* nodes, vpt_bvhnode* leaves, const Morton* __restrict__ codes, const int* __restrict__ ids, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (bvh_lcp(i, i + 1, n, codes, ids) - bvh_lcp(i, i - 1, n, codes, ids)) >= 0 ? 1 : -1;
    const int delta_min = bvh_lcp(i, i - d, n, codes, ids);
    int lmax = 2;
    while (bvh_lcp(i, i + lmax * d, n, codes, ids) > delta_min) lmax *= 2;
    int l = 0, divider = 2;
    for (int t = lmax / divider; t >= 1; divider *= 2) {
        if (bvh_lcp(i, i + (l + t) * d, n, codes, ids) > delta_min) l += t;
        if (t == 1) break;
        t = lmax / divider;
    }
    const int j = i + l * d;
    const int delta_node = bvh_lcp(i, j, n, codes, ids);
    int sp = 0; divider = 2;
    for (int t = (l + (divider - 1)) / divider; t >= 1; divider *= 2) {
        if (bvh_lcp(i, i + (sp + t) * d, n, codes, ids) > delta_node) sp += t;
        if (t == 1) break;
        t = (l + (divider - 1)) / divider;
    }
    const int gamma = i + sp * d + min(d, 0);
    vpt_bvhnode* cur = nodes + i;
    vpt_bvhnode* lc = (min(i, j) == gamma) ? leaves + gamma : nodes + gamma;
    vpt_bvhnode* rc = (max(i, j) == gamma + 1) ? leaves + gamma + 1 : nodes + gamma + 1;
    cur->leftChild = (vpt_devptr_t)(uintptr_t)lc;  lc->parent = (vpt_devptr_t)(uintptr_t)cur;
    cur->rightChild = (vpt_devptr_t)(uintptr_t)rc; rc->parent = (vpt_devptr_t)(uintptr_t)cur;
    cur->minId = min(i, j); cur->maxId = max(i, j);
}

__global__ void k_bvh_refit(vpt_bvhnode* nodes, vpt_bvhnode* leaves, int* counters, const Box* __restrict__ bounds, const int* __restrict__ ids, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vpt_bvhnode* leaf = leaves + i;
    const int vol = ids[i];
    const Box b = bounds[vol];
    leaf->volIndex = vol;
    leaf->boundingBox.pmin.x = b.pmin.x; leaf->boundingBox.pmin.y = b.pmin.y; leaf->boundingBox.pmin.z = b.pmin.z;
    leaf->boundingBox.pmax.x = b.pmax.x; leaf->boundingBox.pmax.y = b.pmax.y; leaf->boundingBox.pmax.z = b.pmax.z;
    if (n == 1) return;                                                  // a single leaf has no parent (quirk Q18)
    __threadfence();
    vpt_bvhnode* cur = reinterpret_cast<vpt_bvhnode*>((uintptr_t)leaf->parent);
    for (;;) {
        if (atomicAdd(counters + (int)(cur - nodes), 1) == 0) return;     // first child to arrive leaves; the second one merges
        __threadfence();
        const volatile vpt_bvhnode* L = reinterpret_cast<const vpt_bvhnode*>((uintptr_t)cur->leftChild);
        const volatile vpt_bvhnode* R = reinterpret_cast<const vpt_bvhnode*>((uintptr_t)cur->rightChild);
        cur->boundingBox.pmin.x = fminf(L->boundingBox.pmin.x, R->boundingBox.pmin.x);
        cur->boundingBox.pmin.y = fminf(L->boundingBox.pmin.y, R->boundingBox.pmin.y);
        cur->boundingBox.pmin.z = fminf(L->boundingBox.pmin.z, R->boundingBox.pmin.z);
        cur->boundingBox.pmax.x = fmaxf(L->boundingBox.pmax.x, R->boundingBox.pmax.x);
        cur->boundingBox.pmax.y = fmaxf(L->boundingBox.pmax.y, R->boundingBox.pmax.y);
        cur->boundingBox.pmax.z = fmaxf(L->boundingBox.pmax.z, R->boundingBox.pmax.z);
        if (cur == nodes) return;
        __threadfence();
        cur = reinterpret_cast<vpt_bvhnode*>((uintptr_t)cur->parent);
    }
}

// d_nodes: max(n-1, 1) zeroed BVHNodes, d_leaves: n zeroed BVHNodes.  scene6 (host) receives the union box
// (thrust::reduce over AABBUnion with the empty box +-M_INF as identity; fmin/fmax is order-independent).
// h_codes / h_ids (host, may be null) receive the sorted Morton codes and instance ids.
cudaError_t bvh_build_device(const vpt_gpu_vdb* d_vols, int n, vpt_bvhnode* d_nodes, vpt_bvhnode* d_leaves, float scene6[6],
                             unsigned long long* h_codes, int* h_ids, cudaStream_t s)
{
    Box* d_bounds = nullptr; Morton *d_codes = nullptr, *d_sorted = nullptr; int *d_ids = nullptr, *d_cnt = nullptr;
    std::vector<Box> hb((size_t)n);
    cudaError_t e = cudaMalloc(&d_bounds, sizeof(Box) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&d_codes, sizeof(Morton) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&d_sorted, sizeof(Morton) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&d_ids, sizeof(int) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&d_cnt, sizeof(int) * (size_t)n);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, sizeof(int) * (size_t)n, s);
    if (e == cudaSuccess) {
        k_instance_bounds<<<(n + 127) / 128, 128, 0, s>>>(d_vols, n, d_bounds);
        e = cudaMemcpyAsync(hb.data(), d_bounds, sizeof(Box) * (size_t)n, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) {
        Box sc; sc.pmin = make_float3(3.402823466e+38F, 3.402823466e+38F, 3.402823466e+38F); sc.pmax = make_float3(-3.402823466e+38F, -3.402823466e+38F, -3.402823466e+38F);
        for (int i = 0; i < n; ++i) {
            sc.pmin.x = fminf(sc.pmin.x, hb[i].pmin.x); sc.pmin.y = fminf(sc.pmin.y, hb[i].pmin.y); sc.pmin.z = fminf(sc.pmin.z, hb[i].pmin.z);
            sc.pmax.x = fmaxf(sc.pmax.x, hb[i].pmax.x); sc.pmax.y = fmaxf(sc.pmax.y, hb[i].pmax.y); sc.pmax.z = fmaxf(sc.pmax.z, hb[i].pmax.z);
        }
        if (scene6) { scene6[0] = sc.pmin.x; scene6[1] = sc.pmin.y; scene6[2] = sc.pmin.z; scene6[3] = sc.pmax.x; scene6[4] = sc.pmax.y; scene6[5] = sc.pmax.z; }
        const int g = (n + 127) / 128;
        k_bvh_morton<<<g, 128, 0, s>>>(d_bounds, n, sc, d_codes);
        k_bvh_rank_sort<<<(n + 255) / 256, 256, 0, s>>>(d_codes, n, d_sorted, d_ids);
        if (n > 1) k_bvh_radix_tree<<<g, 128, 0, s>>>(d_nodes, d_leaves, d_sorted, d_ids, n);
        k_bvh_refit<<<g, 128, 0, s>>>(d_nodes, d_leaves, d_cnt, d_bounds, d_ids, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && h_codes) e = cudaMemcpyAsync(h_codes, d_sorted, sizeof(Morton) * (size_t)n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && h_ids) e = cudaMemcpyAsync(h_ids, d_ids, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(d_bounds); cudaFree(d_codes); cudaFree(d_sorted); cudaFree(d_ids); cudaFree(d_cnt);
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
