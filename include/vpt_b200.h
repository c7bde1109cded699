/*
 * vpt_b200.h -- C ABI of libvpt_b200.so, the B200-native replacement for the reference render pass.
 *
 * Drop-in boundary.  The reference application issues one progressive pass as
 *
 *     void* params[] = { &cam, &l_list, &d_volume_ptr, &d_geo_ptr, &d_geo_list_ptr,
 *                        &bvh_builder.bvh.BVHNodes, &bvh_builder.root, atmos_params, &kernel_params };
 *     cuLaunchKernel(cuRaycastKernel, grid.x, grid.y, 1, 16, 16, 1, 0, NULL, params, NULL);   // main.cpp:1826-1827
 *
 * and the replacement is that one line:
 *
 *     vpt_render_pass(ctx, params, NULL);
 *
 * with the very same `params` array (layouts in vpt_abi.h).  All memory named by the parameters stays
 * owned by the caller, exactly as with the reference kernel; the context only owns its ray queue,
 * sample planes and the flattened scene tables it derives from GPU_VDB[] + the OCTNode tree.
 *
 * Every function returns 0 on success or a negative VPT_ERR_* code; vpt_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device every entry point fails with VPT_ERR_CUDA.
 */
#ifndef VPT_B200_H_
#define VPT_B200_H_

#include "vpt_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VPT_OK                 0
#define VPT_ERR_INVALID       -1   /* bad argument */
#define VPT_ERR_CUDA          -2   /* CUDA runtime / driver failure (text in vpt_last_error) */
#define VPT_ERR_UNSUPPORTED   -3   /* parameter combination this build does not implement */
#define VPT_ERR_IO            -4   /* file could not be read / decoded */

typedef struct vpt_context vpt_context;

/* ---- lifecycle ----------------------------------------------------------------------------------- */
int  vpt_create(vpt_context** out);                    /* binds to the calling thread's current CUDA device */
void vpt_destroy(vpt_context* ctx);
const char* vpt_last_error(const vpt_context* ctx);    /* ctx may be NULL: error of the last failed vpt_create / helper */
const char* vpt_version(void);
/* sizeof() of the boundary structs: camera, light_list, GPU_VDB, sphere, geometry_list, BVHNode, OCTNode,
 * AtmosphereParameters, Kernel_params, point_light, VDB_INFO, AABB.  Returns how many entries exist. No CUDA call. */
int  vpt_abi_sizes(size_t* out, int n);

/* Tunables: "passes_per_chunk" (1..64 passes fused per generate/trace/resolve round; 0 = automatic, the default: 32, or 64 when the
 * local frame has at most 2^20 pixels), "max_scratch_mb" (cap of the per-round ray queue + sample planes, default 12288),
 * "gather_async" (see vpt_comm_*),
 * "sched_min_lanes" (1..32, lanes an operation must gather in a warp before it pre-empts stepping; 0 = by kernel, the default: 26 for the
 * lean kernel at 3 rays per lane, else 20),
 * "trace_slots" (rays per lane of the trace kernel: 0 = by scene, the default -- 2 when the density grid is larger than twice the L2 or
 * the scene has 64 or more instances, else 3; 2 or 3 to force it),
 * "l2_sector_fetch" (0|1: while a cell table is set, ask the L2 for 32-byte DRAM fetches -- a device-wide hint, see vpt_set_cell_volume),
 * "ctas_per_sm" (0 = occupancy maximum), "generic_kernel" (1: never pick the lean instantiation), "debug_flags" (development),
 * "count_stats" / "profile" (0|1, see vpt_get_counters / vpt_get_kernel_times). */
int  vpt_set_option(vpt_context* ctx, const char* key, int value);

/* Multi-GPU partition of the frame: rank r of n renders the interleaved row stripes r, r+n, ... of
 * `stripe_rows` rows.  With n_ranks > 1 every buffer inside Kernel_params (accum, depth, cost, raw,
 * display) holds vpt_local_pixels() pixels in rank-local order; the blue-noise buffer stays a full
 * 256x256 replica.  Default: one rank, identity layout (bit-compatible with the reference buffers). */
int  vpt_set_partition(vpt_context* ctx, int rank, int n_ranks, int stripe_rows);
long long vpt_local_pixels(const vpt_context* ctx, unsigned width, unsigned height);
/* After an all-gather of rank-local buffers ([rank][local pixel], elem_bytes per pixel, multiple of 4)
 * scatter them back into a full row-major frame. */
int  vpt_unpermute(vpt_context* ctx, const void* d_gathered, void* d_full, unsigned width, unsigned height,
                   int elem_bytes, void* stream);

/* Multi-GPU exchange, issued by the library itself (one process per GPU, NCCL).  One rank obtains an id
 * (ncclGetUniqueId) and hands the 128 bytes to every rank by whatever means the application has (MPI, torch.distributed,
 * a file); every rank then calls vpt_comm_init (collective: ncclCommInitRank; also applies vpt_set_partition).  With a
 * gather target set, every vpt_render_pass(es) call ends with ONE ncclAllGather of the rank-local accumulators (and one of
 * the display words if asked) + the stripe un-permutation into the caller's full-frame buffers -- row-major [H][W] float3 /
 * uint32, allocated by the caller on every rank -- enqueued on the caller's stream behind the last kernel.  Option
 * "gather_async" = 1 moves that exchange to a side stream so the next call's sampling kernels overlap it; the library waits
 * for it before the next call overwrites the accumulator, and vpt_comm_wait(ctx, stream) makes `stream` wait for it.
 * NCCL is bound at run time (dlopen): without it these entry points fail with VPT_ERR_UNSUPPORTED, everything else works. */
#define VPT_COMM_ID_BYTES 128
int  vpt_comm_get_unique_id(unsigned char id_out[VPT_COMM_ID_BYTES]);
int  vpt_comm_init(vpt_context* ctx, const unsigned char id[VPT_COMM_ID_BYTES], int rank, int n_ranks, int stripe_rows);
int  vpt_comm_set_gather(vpt_context* ctx, void* d_full_accum_f3, void* d_full_display_u32);   /* either may be NULL */
int  vpt_comm_wait(vpt_context* ctx, void* stream);
int  vpt_comm_info(vpt_context* ctx, int* nccl_version, int* rank, int* n_ranks);
int  vpt_comm_destroy(vpt_context* ctx);

/* Peer-memory exchange: the same result without NCCL and without a separate gather step, for up to 8 GPUs of one NVLink / NVSwitch
 * domain (one process per GPU).  Every rank allocates ONE exchange block {flags, full-frame float3 accumulators, optional display words}
 * and exports it as a CUDA IPC handle (vpt_comm_p2p_export: also applies vpt_set_partition); the application hands all handles to all
 * ranks (64 bytes each, rank order) and every rank maps its peers (vpt_comm_p2p_import).  From then on the LAST resolve kernel of every
 * vpt_render_pass(es) call stores each finished pixel straight into every rank's frame at its global position (all-gather and stripe
 * un-permutation fused into the producing kernel, over peer mappings), bracketed by two flag exchanges in peer memory: "my previous frame
 * has been consumed" at the start of the call, "my stripes are in place" at the end; the call's stream work ends with a wait for all
 * peers, so after it the frame returned by vpt_comm_p2p_frame is complete on every rank, bit-identical to a single-GPU frame.
 * vpt_comm_p2p_enable(ctx, 0) suspends the exchange for rank-local work (calls that not every rank makes); vpt_comm_p2p_status reports
 * how many flag waits gave up (a wait abandons after 10 s instead of hanging the GPU: a rank died or made fewer calls). */
#define VPT_P2P_HANDLE_BYTES 64
int  vpt_comm_p2p_export(vpt_context* ctx, int rank, int n_ranks, int stripe_rows, unsigned width, unsigned height, int with_display,
                         unsigned char handle_out[VPT_P2P_HANDLE_BYTES]);
int  vpt_comm_p2p_import(vpt_context* ctx, const unsigned char* handles /* n_ranks x VPT_P2P_HANDLE_BYTES, rank order */);
/* Same-process peers (one process driving several contexts -- several GPUs with peer access enabled, or several shards on one GPU on
 * separate streams): exchange the block addresses (vpt_comm_p2p_block) instead of IPC handles. */
int  vpt_comm_p2p_block(vpt_context* ctx, vpt_devptr_t* d_block);
int  vpt_comm_p2p_import_local(vpt_context* ctx, const vpt_devptr_t* blocks /* n_ranks addresses, rank order */);
int  vpt_comm_p2p_frame(vpt_context* ctx, vpt_devptr_t* d_full_accum_f3, vpt_devptr_t* d_full_display_u32);
int  vpt_comm_p2p_enable(vpt_context* ctx, int on);
int  vpt_comm_p2p_status(vpt_context* ctx, unsigned long long* abandoned_waits);

/* ---- the hot path --------------------------------------------------------------------------------
 * vpt_render_pass   replaces ONE launch of the reference `volume_rt_kernel` (render_kernel.cu:2216):
 *                   same inputs, same buffer updates (accum/depth/cost running means, display, raw,
 *                   blue-noise advance).  The caller keeps doing `++kernel_params.iteration`.
 * vpt_render_passes is n consecutive such launches with iteration, iteration+1, ... fused into as few
 *                   kernel rounds as the chunk size allows; the buffers afterwards equal the state after
 *                   n reference launches (display/raw are written once, from the final accumulator).
 * `params` is the cuLaunchKernel argument array (VPT_ARG_* in vpt_abi.h). `stream` is a cudaStream_t.
 * Work is enqueued asynchronously; synchronise the stream (or the device) as the reference loop does. */
int  vpt_render_pass(vpt_context* ctx, void* const params[VPT_NUM_ARGS], void* stream);
int  vpt_render_passes(vpt_context* ctx, void* const params[VPT_NUM_ARGS], unsigned n_passes, void* stream);

/* The flattened scene tables are cached per (volumes pointer, octree root pointer, build generation of a vpt_octree_build
 * octree) and refreshed whenever a call arrives with Kernel_params.iteration == 0 (the reference application resets the
 * iteration on every scene edit, main.cpp:1667-1780).  Call this only if the GPU_VDB[] array or a foreign octree was
 * rewritten in place WITHOUT restarting the accumulation. */
int  vpt_invalidate_scene(vpt_context* ctx);

/* Counters of the last vpt_render_pass(es) call: kernels launched, and (after the stream has been
 * synchronised by the caller) rays pushed into the hit queue by the final chunk. */
int  vpt_get_stats(vpt_context* ctx, unsigned long long* kernel_launches_total, unsigned* last_queue_count);

/* Instrumentation.  Option "count_stats" = 1 makes the trace kernel accumulate out[0] volume lookups,
 * out[1] lane-steps, out[2] warp step-loop iterations, out[3] rays serviced, out[4] warp service rounds, out[5] rays
 * fetched from the queue, out[6] bricks staged by TMA (brick mode) (SIMT efficiency of the step loop = out[1] / (32 * out[2])).  Option "profile" = 1 brackets every kernel with
 * CUDA events: ms/n[0..3] = generate, trace, resolve, blue-noise advance.  Both calls synchronise the device. */
int  vpt_get_counters(vpt_context* ctx, unsigned long long out[8], int reset);
int  vpt_get_kernel_times(vpt_context* ctx, float ms[4], int n[4]);

/* ---- host-side scene helpers (what the reference does in main.cpp / gpu_vdb.cpp / bvh_builder.cpp) -- */

/* Dense grid -> 3-D texture, as GPU_VDB::loadVDB builds them (gpu_vdb.cpp:215-248: cudaArray, normalised
 * coordinates, linear filter, clamp).  channels = 1 (float) or 4 (float4).  Returns the texture object and
 * an opaque array handle for vpt_texture_destroy. */
int  vpt_texture_create_3d(const float* host_data, int channels, int dim_x, int dim_y, int dim_z,
                           vpt_tex_t* tex_out, void** array_out);
/* Same texture from a dense grid that already lives in device memory (x fastest), e.g. a procedural grid of several GiB. */
int  vpt_texture_create_3d_from_device(const float* d_data, int channels, int dim_x, int dim_y, int dim_z,
                                       vpt_tex_t* tex_out, void** array_out);
/* Procedural density grid: the reference's fill_volume_buffer (texture_kernels.cu:76-128) for noise_type 0 =
 * cudaNoise::perlinNoise(pos, scale, seed) at voxel (x, y, z), x fastest -- with ZERO sub-voxel jitter (the reference draws
 * its jitter from an uninitialised generator state, quirk Q14).  d_buffer: dx*dy*dz floats in device memory.  The matching
 * VDB_INFO is what GPU_PROC_VOL::create_volume fills (gpu_vdb.cpp:529-541): bmin = box min, bmax = bmin + dim, voxelsize = res,
 * max_density 1, min_density 0, xform = scale(res). */
int  vpt_procedural_fill(float* d_buffer, int dim_x, int dim_y, int dim_z, int noise_type, float scale, int seed, void* stream);

/* ---- brick mode: brick pool + TMA-staged software sampler (the look-up without the texture unit) ---------------------
 * vpt_bricks_create re-lays a dense device grid as 4x4x4-cell bricks stored with their +1 apron (5x5x5 texels + brick max /
 * min: 512 contiguous bytes each, edge texels clamped like the reference's clamp-addressed texture).  vpt_set_brick_volume
 * makes the context trace volume 0 from that pool: k_trace_brick stages the brick under each ray into shared memory with
 * one cp.async.bulk (TMA) and filters it in software with the texture unit's own arithmetic, measured on the device:
 * coordinate truncated to 21 fractional bits, eight integer corner weights that sum to 256 (split z -> x -> y), correctly
 * rounded sum.  99.6 % of the fetches are bit-identical to tex3D, the rest one ulp off; rendered frames meet the same
 * per-pixel tolerance as the texture path (tests/test_bricks_gpu.py).  Supported: a vpt_octree_build scene of one volume
 * without colour grid, direct integrator, no emission, no point lights; anything else is refused with VPT_ERR_UNSUPPORTED.
 * d_pool = 0 returns the context to the texture path. */
int  vpt_bricks_create(const float* d_dense, int dim_x, int dim_y, int dim_z, vpt_devptr_t* d_pool_out, unsigned long long* bytes_out);
int  vpt_bricks_read(vpt_devptr_t d_pool, unsigned long long first_brick, unsigned long long n_bricks, float* h_out);   /* 128 floats per brick */
/* Diagnostic: the software filter under three weight rules (m = 0 the texture unit's integer corner weights -- the production rule --,
 * 1 per-axis weights truncated to 1/256, 2 per-axis full fp32 fractions), reading the bricks from global memory, against tex3D on n pseudo-random points: out12[m*4 + 0..3] = max |d|, sum |d|,
 * number of bit-identical results, n. */
int  vpt_debug_sampler_compare(vpt_tex_t tex, vpt_devptr_t d_pool, int dim_x, int dim_y, int dim_z, int n_points, unsigned seed, double out12[12]);
int  vpt_bricks_destroy(vpt_devptr_t d_pool);
int  vpt_set_brick_volume(vpt_context* ctx, vpt_devptr_t d_pool, int dim_x, int dim_y, int dim_z);

/* ---- cell mode: one DRAM sector per look-up for grids no cache can hold -------------------------------------------
 * vpt_cells_create re-lays a dense device grid as a CELL TABLE: for every texel cell (i, j, k) the eight corner texels a
 * trilinear look-up in that cell blends (clamped at the upper faces), 32 contiguous, 32-byte aligned bytes, x fastest.  Eight
 * times the memory of the grid (1024^3: 32 GiB -- sized for 180 GB of HBM), and ONE sector per look-up where the 2x2x2
 * footprint costs four or five in the tiled array of the texture path.  vpt_set_cell_volume makes the trace kernel read volume
 * 0 from it (lean instantiation, 2 rays per lane, the texture unit's arithmetic in software as in brick mode: same per-pixel
 * tolerance, tests/test_bricks_gpu.py); same scene restrictions as brick mode; d_cells = 0 returns to the texture path. */
int  vpt_cells_create(const float* d_dense, int dim_x, int dim_y, int dim_z, vpt_devptr_t* d_cells_out, unsigned long long* bytes_out);
int  vpt_cells_read(vpt_devptr_t d_cells, unsigned long long first_cell, unsigned long long n_cells, float* h_out);       /* 8 floats per cell */
int  vpt_cells_destroy(vpt_devptr_t d_cells);
int  vpt_set_cell_volume(vpt_context* ctx, vpt_devptr_t d_cells, int dim_x, int dim_y, int dim_z);

/* Equirectangular environment map float4 (main.cpp:945-978: wrap / clamp, linear, normalised). */
int  vpt_texture_create_env(const float* host_rgba, unsigned width, unsigned height, vpt_tex_t* tex_out, void** array_out);
/* The four sky-sampling tables the volumetric path integrator reads when environment_type == 0 (reference create_cdf,
 * main.cpp:647-867): from a res x res table of the sky's luminous power over (azimuth = x, elevation = y) builds the
 * row-conditional cdf, the marginal function and its cdf, and wraps all four as point-sampled unnormalised float
 * textures (2-D, 2-D, 1-D, 1-D).  tex_out / arrays_out order: env_func_tex, env_cdf_tex, env_marginal_func_tex,
 * env_marginal_cdf_tex; *marginal_int_out is Kernel_params.env_marginal_int; env_sample_tex_res = res.  The arithmetic is
 * vpt_env_tables_compute's. */
int  vpt_env_tables_create(const float* func, unsigned res, vpt_tex_t tex_out[4], void* arrays_out[4], float* marginal_int_out);
/* The host arithmetic of create_cdf (main.cpp:681-757) on caller-provided arrays (cdf: res*res, marginal_func / marginal_cdf:
 * res), literally: row y > 0 of the conditional cdf starts from func[y-1][res-1] / res (the reference's `*(func_p - 1)` at x == 0
 * is the previous row's last element), the uniform fallback is selected by marginal_func[0] == 0, the last cdf entry of each
 * row is 1.  The only two reads the reference makes OUTSIDE its arrays (func[-1], marginal_cdf[-1]; quirk Q20) are taken as 0.
 * Pure host code, no device needed. */
int  vpt_env_tables_compute(const float* func, unsigned res, float* cdf, float* marginal_func, float* marginal_cdf, float* marginal_int_out);
/* The table create_cdf feeds into the above: luminous power of the reference's HOST-side analytic sky (single-scattering
 * Rayleigh + Mie march, 16 x 8 samples, main.cpp:242-301) over res x res directions, func[y*res + x] =
 * |sky(dir(az = x/(res-1)*2pi, el = y/(res-1)*pi)) * sky_color| (main.cpp:683-693; the reference uses res = 180 and
 * Kernel_params.azimuth / elevation / sky_color).  Pure host code, no device needed. */
int  vpt_env_sky_tabulate(float azimuth_deg, float elevation_deg, const float sky_color[3], unsigned res, float* func_out);
int  vpt_texture_destroy(vpt_tex_t tex, void* array);

/* Minimal OpenVDB (file format 224) reader: densify grid `grid_name` over its active-voxel bounding box
 * (gpu_vdb.cpp:133-212).  On success *values_out is malloc'd (x fastest, `channels` floats per voxel; free
 * with vpt_free) and info[] receives: dim[3], bbox_min[3], bbox_max[3], channels, leaf_count,
 * active_voxel_count (lo, hi 32 bits)  -> 13 ints; xform16 receives the reference's mat4 memory image
 * (gpu_vdb.cpp:81-92, convert_to_mat4); stats[] = { max_value, min_density per Q10, voxel_size, background }.
 * Returns 1 if the file has no such grid. */
int  vpt_vdb_load(const char* path, const char* grid_name, float** values_out, int info[13], float xform16[16], float stats[4]);
int  vpt_hdr_load(const char* path, float** rgba_out, unsigned* width, unsigned* height);   /* hdr_loader.h:249-278 */
int  vpt_bmp_load_rbg(const char* path, float** xyz_out, int* width, int* height);          /* fileIO.cpp:460-495 (Q16) */
int  vpt_exr_load_rgb(const char* path, float** rgb_out, int* width, int* height);          /* fileIO.cpp:356-390 */
void vpt_free(void* p);

/* Scene / light instance file (.ins) reader -- the text format of read_instance_file, main.cpp:980-1040, SURVEY 8(f) N4:
 *     "light" \n N \n { px py pz r g b power } x N                                  (point lights)
 *     N_vdbs \n { vdb_path \n N_inst \n { px py pz qx qy qz qw scale } x N_inst } x N_vdbs   (instanced volumes)
 * One record per line, whitespace separated, paths taken verbatim (the reference resolves them against its CWD).
 * *out is one malloc'd block (free with vpt_free): the header below, then n_files vpt_ins_file_entry, then n_records records
 * of 8 doubles each (volumes: px py pz qx qy qz qw scale; lights: px py pz r g b power 0).  Unlike the reference, a line
 * with too few numbers is an error (VPT_ERR_IO) instead of leaving the fields uninitialised.  The instance transform itself
 * (mat4 algebra of main.cpp:1059-1099) is applied by the caller; see volumetric-path-tracer_b200/scene.py:instance_xform. */
typedef struct vpt_ins_header { int32_t kind; int32_t n_files; int32_t n_records; int32_t reserved; } vpt_ins_header;   /* kind: 0 volumes, 1 lights */
typedef struct vpt_ins_file_entry { char path[1024]; int32_t first_record; int32_t n_instances; } vpt_ins_file_entry;
int  vpt_ins_load(const char* path, vpt_ins_header** out);

/* Instance acceleration build -- replaces BVH_Builder::build_bvh (bvh_builder.cpp:46-105).
 *
 * vpt_octree_build: depth-3 octree over the instance bounds (host root set-up bvh_builder.cpp:61-78, node semantics of
 * bvh_kernels.cu:204-246: child boxes by halving, inclusive overlap test, ascending instance order), built in parallel
 * with no device heap.  Any n >= 1:
 *   n <= VPT_OCT_MAX_VOLUMES  *d_root_out is node 0 of a contiguous 585-node tree in the reference's own OCTNode layout
 *                             (pointer-linked; the reference kernel can consume it as well);
 *   n >  VPT_OCT_MAX_VOLUMES  the reference layout cannot hold the lists (vol_indices[600]; the reference itself overflows
 *                             the node there, quirk Q11): *d_root_out points to the root record only (children null).
 * In both cases the octree is ALSO kept in the flat form the render kernels read (73 internal nodes + 512 leaf lists in
 * CSR form, no per-leaf limit) and registered under the returned root pointer: vpt_render_pass(es) recognises the pointer
 * in params[VPT_ARG_OCTREE] and uses those tables directly.  h_volumes: host array of n GPU_VDB; the GPU_VDB[] device
 * array passed at render time must describe the same instances.  Free with vpt_octree_destroy. */
int  vpt_octree_build(const vpt_gpu_vdb* h_volumes, int n, vpt_devptr_t* d_root_out);
int  vpt_octree_destroy(vpt_devptr_t d_root);
/* What vpt_octree_build registered for this root: instance count, whether reference-layout nodes exist (n <= 600), total
 * and largest leaf-list length.  Any output pointer may be NULL. */
int  vpt_octree_info(vpt_devptr_t d_root, int* n_instances, int* reference_layout, long long* total_leaf_entries, int* max_leaf_entries);
/* Copy any pointer-linked octree (this builder's n <= 600 tree or the reference's device-heap one) into 585 host nodes in
 * the canonical numbering 0 | 1+c1 | 9+c1*8+c2 | 73+c1*64+c2*8+c3; exists[j] = 0 for nodes never allocated. */
int  vpt_octree_read(vpt_devptr_t d_root, vpt_octnode* h_nodes585, int* h_exists585);
/* Flat leaf lists of a registered octree: leaf_offset_count[2*l] / [2*l+1] = offset / length of leaf l = c1*64+c2*8+c3 in
 * `indices` (ascending instance ids; capacity in ints, indices may be NULL to fetch the table only). */
int  vpt_octree_read_flat(vpt_devptr_t d_root, unsigned leaf_offset_count[1024], int* indices, long long capacity);

/* LBVH over the instance AABBs in the reference's BVHNode layout (BuildBVH, bvh_kernels.cu:460-580: Morton codes of the
 * centroids in the scene box, stable sort, Karras radix tree with the (code, id) tie-break, bottom-up refit).
 * *d_nodes_out: n-1 internal nodes (node 0 is the root; one zeroed node when n == 1), *d_leaves_out: n leaves in sorted
 * order; child / parent fields are device pointers into the two arrays, exactly what the reference passes as `root_node`.
 * scene_bounds6 (may be NULL): union of the instance boxes.  h_sorted_codes / h_sorted_ids (host, n entries, may be NULL):
 * the sorted 30-bit Morton codes and instance ids.  Unlike the reference the arrays are zero-initialised and n == 1 is
 * handled (quirk Q18).  The live integrators of the reference never traverse this tree; it is built for
 * callers that pass it on. */
int  vpt_bvh_build(const vpt_gpu_vdb* h_volumes, int n, vpt_devptr_t* d_nodes_out, vpt_devptr_t* d_leaves_out, float scene_bounds6[6],
                   unsigned long long* h_sorted_codes, int* h_sorted_ids);
/* Copy a BVH in that layout (either builder's) to the host, pointer fields rewritten as indices: internal node i -> i,
 * leaf i -> (n-1)+i, anything else -> (uint64)-1. */
int  vpt_bvh_read(vpt_devptr_t d_nodes, vpt_devptr_t d_leaves, int n, vpt_bvhnode* h_nodes, vpt_bvhnode* h_leaves);
int  vpt_bvh_destroy(vpt_devptr_t d_nodes, vpt_devptr_t d_leaves);
/* AABB of one instance (GPU_VDB::Bounds, gpu_vdb.h:131-146): out6 = pmin, pmax. */
void vpt_volume_bounds(const vpt_gpu_vdb* h_volume, float out6[6]);

/* ---- Bruneton sky tables (replaces atmosphere::init, source/atmosphere/atmosphere.cpp:1177-1291 + atmosphere_kernels.cu) ------------
 * Builds the Earth model of the reference (its spectra, density profiles, radii), reduces it to the AtmosphereParameters block the
 * render path takes by value, precomputes the transmittance / irradiance / scattering / single-Mie tables on the device (own sm_100a
 * kernels, four scattering orders) and wraps them as the four look-up textures with the reference's descriptors.  *out receives the
 * complete block (scalars + texture objects; scratch buffer pointers null); free with vpt_atmosphere_destroy(handle).
 * The tables reproduce the reference's build of the model, including the places where it departs from Bruneton's paper (later
 * orders overwrite instead of accumulate because the host passes a float4 where its kernels read an int: quirk Q17; mie_extinction
 * taken from the Mie scattering spectrum), except that table look-ups past the last texel are clamped where the reference reads
 * beyond its buffers (atmosphere_kernels.cu:366-372): the outermost texels of orders >= 2 differ by construction. */
typedef struct vpt_atmosphere_options {
    int   use_constant_solar_spectrum;   /* main.cpp:1433 default 1 */
    int   use_ozone;                     /* default 1 */
    int   luminance_mode;                /* 0 NONE (default), 1 APPROXIMATE, 2 PRECOMPUTED (15 wavelengths) */
    int   do_white_balance;              /* default 1 */
    float exposure;                      /* default 1 */
    int   num_scattering_orders;         /* default 4 */
} vpt_atmosphere_options;
void vpt_atmosphere_options_defaults(vpt_atmosphere_options* opt);
int  vpt_atmosphere_precompute(const vpt_atmosphere_options* opt, vpt_atmosphere* out, void** handle_out);
int  vpt_atmosphere_destroy(void* handle);
/* Texel centres of a float4 texture (w x h, or w x h x d when d > 0) into host memory, 4 floats per texel, x fastest. */
int  vpt_texture_read_f4(vpt_tex_t tex, int w, int h, int d, float* host_out);
/* Diagnostic: tex3D<float> at n arbitrary normalised coordinates (uvw: 3 floats per point). */
int  vpt_debug_texture_sample(vpt_tex_t tex, const float* uvw, int n, float* out);

/* Thin-lens camera set-up (camera::update_camera, camera.h:110-129). */
void vpt_camera_look_at(vpt_camera* cam, const float lookfrom[3], const float lookat[3], const float vup[3],
                        float vfov_deg, float aspect, float aperture);
/* Kernel_params defaults of main.cpp:1350-1376 with the frame-loop overrides (:1533-1544). */
void vpt_kernel_params_defaults(vpt_kernel_params* kp);

#ifdef __cplusplus
}
#endif
#endif /* VPT_B200_H_ */
