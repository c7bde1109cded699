/*
 * vpt_abi.h -- binary layout of the launch-parameter structs at the drop-in boundary.
 *
 * The reference hands `volume_rt_kernel` nine parameters (source/render_kernel.cu:2216-2225,
 * launched from source/main.cpp:1823-1827).  A replacement for that launch has to accept the very
 * same bytes, so every struct that crosses the boundary is re-declared here in plain C with the
 * layout nvcc 12.9 / gcc 13.3 (x86-64 Linux) gives the reference's own headers.  Sizes and offsets
 * are locked with static asserts; tests/test_abi.py re-measures them against the reference headers
 * when /root/reference is present.
 *
 * Nothing here is copied from the reference: these are independent declarations of the same
 * memory layout, each one citing the reference declaration it mirrors.
 */
#ifndef VPT_ABI_H_
#define VPT_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define VPT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define VPT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

#if defined(__CUDACC__) || defined(__GNUC__)
#define VPT_ALIGN(n) __attribute__((aligned(n)))
#else
#define VPT_ALIGN(n)
#endif

typedef struct vpt_f3 { float x, y, z; } vpt_f3;               /* CUDA float3: 12 B, align 4 */
typedef struct VPT_ALIGN(16) vpt_f4 { float x, y, z, w; } vpt_f4; /* CUDA float4: 16 B, align 16 */
typedef struct vpt_i3 { int32_t x, y, z; } vpt_i3;             /* CUDA int3 */
typedef struct VPT_ALIGN(8) vpt_u2 { uint32_t x, y; } vpt_u2;  /* CUDA uint2: 8 B, align 8 */
typedef uint64_t vpt_tex_t;                                    /* cudaTextureObject_t */
typedef uint64_t vpt_devptr_t;                                 /* device address */

/* ---- camera: source/gpu_vdb/camera.h:94-148 (104 B, align 4) ------------------------------- */
typedef struct vpt_camera {
    float   time1, time0;
    vpt_f3  origin;
    float   focus_dist;
    vpt_f3  lower_left_corner;
    vpt_f3  horizontal;
    vpt_f3  vertical;
    vpt_f3  u, v, w;
    float   lens_radius;
    uint8_t viz_dof;            /* bool */
    uint8_t _pad[3];
} vpt_camera;
VPT_STATIC_ASSERT(sizeof(vpt_camera) == 104, "camera is 104 B");
VPT_STATIC_ASSERT(offsetof(vpt_camera, origin) == 8 && offsetof(vpt_camera, focus_dist) == 20 &&
                  offsetof(vpt_camera, lower_left_corner) == 24 && offsetof(vpt_camera, horizontal) == 36 &&
                  offsetof(vpt_camera, vertical) == 48 && offsetof(vpt_camera, u) == 60 &&
                  offsetof(vpt_camera, lens_radius) == 96 && offsetof(vpt_camera, viz_dof) == 100,
                  "camera field offsets");

/* ---- point_light / light_list: source/light.h:93-167 ---------------------------------------- */
typedef struct vpt_point_light {      /* 48 B: vptr, pos, dir, power, color */
    uint64_t _vptr;
    vpt_f3   pos;
    vpt_f3   dir;
    float    power;
    vpt_f3   color;
} vpt_point_light;
VPT_STATIC_ASSERT(sizeof(vpt_point_light) == 48 && offsetof(vpt_point_light, pos) == 8 &&
                  offsetof(vpt_point_light, power) == 32 && offsetof(vpt_point_light, color) == 36,
                  "point_light layout");

typedef struct vpt_light_list {       /* 16 B, align 8 */
    uint32_t     num_lights;
    uint32_t     _pad;
    vpt_devptr_t light_ptr;           /* point_light[num_lights], managed memory in the reference */
} vpt_light_list;
VPT_STATIC_ASSERT(sizeof(vpt_light_list) == 16 && offsetof(vpt_light_list, light_ptr) == 8, "light_list layout");

/* ---- VDB_INFO / GPU_VDB: source/gpu_vdb/gpu_vdb.h:59-154 ------------------------------------ */
typedef struct VPT_ALIGN(16) vpt_vdb_info {   /* 80 B, align 16 */
    float     voxelsize;
    vpt_i3    dim;
    vpt_f3    bmin;
    vpt_f3    bmax;
    float     max_density;
    float     min_density;
    uint8_t   has_color, has_emission, matte;
    uint8_t   _pad[5];
    vpt_tex_t density_texture;
    vpt_tex_t emission_texture;
    vpt_tex_t color_texture;
} vpt_vdb_info;
VPT_STATIC_ASSERT(sizeof(vpt_vdb_info) == 80, "VDB_INFO is 80 B");
VPT_STATIC_ASSERT(offsetof(vpt_vdb_info, dim) == 4 && offsetof(vpt_vdb_info, bmin) == 16 &&
                  offsetof(vpt_vdb_info, bmax) == 28 && offsetof(vpt_vdb_info, max_density) == 40 &&
                  offsetof(vpt_vdb_info, min_density) == 44 && offsetof(vpt_vdb_info, has_color) == 48 &&
                  offsetof(vpt_vdb_info, density_texture) == 56 && offsetof(vpt_vdb_info, color_texture) == 72,
                  "VDB_INFO field offsets");

/* mat4 is float m[4][4] addressed m[col][row] (source/gpu_vdb/matrix_math.h:49-71);
 * GPU_VDB::xform.transpose() is the index->world matrix in column-vector form. */
typedef struct VPT_ALIGN(16) vpt_gpu_vdb {    /* 144 B */
    vpt_vdb_info vdb_info;
    float        xform[4][4];
} vpt_gpu_vdb;
VPT_STATIC_ASSERT(sizeof(vpt_gpu_vdb) == 144 && offsetof(vpt_gpu_vdb, xform) == 80, "GPU_VDB layout");

/* ---- AABB / OCTNode / BVHNode: source/bvh/AABB.h:46-234, source/bvh/bvh.h:11-24 -------------- */
typedef struct vpt_aabb { vpt_f3 pmin, pmax; } vpt_aabb;
VPT_STATIC_ASSERT(sizeof(vpt_aabb) == 24, "AABB is 24 B");

#define VPT_OCT_MAX_VOLUMES 600               /* OCTNode::vol_indices[600], AABB.h:222 */
typedef struct vpt_octnode {                  /* 2520 B, pointer-linked on the device heap */
    int32_t      num_volumes;
    int32_t      vol_indices[VPT_OCT_MAX_VOLUMES];
    float        max_extinction;
    float        min_extinction;
    float        voxel_size;
    int32_t      depth;
    uint8_t      has_children;
    uint8_t      _pad[3];
    vpt_devptr_t children[8];
    vpt_devptr_t parent;
    vpt_aabb     bbox;
} vpt_octnode;
VPT_STATIC_ASSERT(sizeof(vpt_octnode) == 2520, "OCTNode is 2520 B");
VPT_STATIC_ASSERT(offsetof(vpt_octnode, max_extinction) == 2404 && offsetof(vpt_octnode, min_extinction) == 2408 &&
                  offsetof(vpt_octnode, voxel_size) == 2412 && offsetof(vpt_octnode, depth) == 2416 &&
                  offsetof(vpt_octnode, has_children) == 2420 && offsetof(vpt_octnode, children) == 2424 &&
                  offsetof(vpt_octnode, parent) == 2488 && offsetof(vpt_octnode, bbox) == 2496,
                  "OCTNode field offsets");

typedef struct vpt_bvhnode {                  /* 64 B */
    int32_t      minId, maxId, volIndex, _pad;
    vpt_devptr_t leftChild, rightChild, parent;
    vpt_aabb     boundingBox;
} vpt_bvhnode;
VPT_STATIC_ASSERT(sizeof(vpt_bvhnode) == 64 && offsetof(vpt_bvhnode, leftChild) == 16 &&
                  offsetof(vpt_bvhnode, boundingBox) == 40, "BVHNode layout");

/* ---- sphere / geometry_list: source/geometry/geometry.h:82-172, 237-286 ---------------------- */
typedef struct vpt_sphere {                   /* 40 B: vptr (never dereferenced on device), centre, r, colour, roughness */
    uint64_t _vptr;
    vpt_f3   center;
    float    radius;
    vpt_f3   color;
    float    roughness;
} vpt_sphere;
VPT_STATIC_ASSERT(sizeof(vpt_sphere) == 40 && offsetof(vpt_sphere, center) == 8 && offsetof(vpt_sphere, radius) == 20 &&
                  offsetof(vpt_sphere, color) == 24 && offsetof(vpt_sphere, roughness) == 36, "sphere layout");

typedef struct vpt_geometry_list { vpt_devptr_t list; int32_t list_size; int32_t _pad; } vpt_geometry_list;
VPT_STATIC_ASSERT(sizeof(vpt_geometry_list) == 16, "geometry_list is 16 B");

/* ---- AtmosphereParameters: source/atmosphere/definitions.h:36-99 (464 B, align 16) ----------- */
typedef struct VPT_ALIGN(16) vpt_density_layer {   /* 32 B */
    float width, exp_term, exp_scale, linear_term, const_term;
    float _pad[3];
} vpt_density_layer;
typedef struct VPT_ALIGN(16) vpt_density_profile { vpt_density_layer layers[2]; } vpt_density_profile;
VPT_STATIC_ASSERT(sizeof(vpt_density_layer) == 32 && sizeof(vpt_density_profile) == 64, "DensityProfile layout");

typedef struct VPT_ALIGN(16) vpt_atmosphere {
    vpt_f3   sky_spectral_radiance_to_luminance;   /*   0 */
    vpt_f3   sun_spectral_radiance_to_luminance;   /*  12 */
    vpt_f3   solar_irradiance;                     /*  24 */
    float    angle;                                /*  36 */
    float    bottom_radius;                        /*  40 */
    float    top_radius;                           /*  44 */
    int32_t  use_luminance;                        /*  48 */
    uint8_t  _pad0[12];
    vpt_density_profile rayleigh_density;          /*  64 */
    vpt_f3   rayleigh_scattering;                  /* 128 */
    uint8_t  _pad1[4];
    vpt_density_profile mie_density;               /* 144 */
    vpt_f3   mie_scattering;                       /* 208 */
    vpt_f3   mie_extinction;                       /* 220 */
    float    mie_phase_function_g;                 /* 232 */
    uint8_t  _pad2[4];
    vpt_density_profile absorption_density;        /* 240 */
    vpt_f3   absorption_extinction;                /* 304 */
    vpt_f3   ground_albedo;                        /* 316 */
    float    sun_angular_radius;                   /* 328 */
    float    mu_s_min;                             /* 332 */
    float    exposure;                             /* 336 */
    vpt_f3   white_point;                          /* 340 */
    vpt_devptr_t scratch_buffers[9];               /* 352: precompute-only float4* buffers */
    vpt_tex_t transmittance_texture;               /* 424 */
    vpt_tex_t scattering_texture;                  /* 432 */
    vpt_tex_t irradiance_texture;                  /* 440 */
    vpt_tex_t single_mie_scattering_texture;       /* 448 */
    uint8_t  _pad3[8];
} vpt_atmosphere;
VPT_STATIC_ASSERT(sizeof(vpt_atmosphere) == 464, "AtmosphereParameters is 464 B");
VPT_STATIC_ASSERT(offsetof(vpt_atmosphere, rayleigh_density) == 64 && offsetof(vpt_atmosphere, rayleigh_scattering) == 128 &&
                  offsetof(vpt_atmosphere, mie_density) == 144 && offsetof(vpt_atmosphere, mie_scattering) == 208 &&
                  offsetof(vpt_atmosphere, mie_phase_function_g) == 232 && offsetof(vpt_atmosphere, absorption_density) == 240 &&
                  offsetof(vpt_atmosphere, absorption_extinction) == 304 && offsetof(vpt_atmosphere, ground_albedo) == 316 &&
                  offsetof(vpt_atmosphere, sun_angular_radius) == 328 && offsetof(vpt_atmosphere, white_point) == 340 &&
                  offsetof(vpt_atmosphere, scratch_buffers) == 352 && offsetof(vpt_atmosphere, transmittance_texture) == 424 &&
                  offsetof(vpt_atmosphere, single_mie_scattering_texture) == 448, "AtmosphereParameters offsets");

/* ---- Kernel_params: source/kernel_params.h:39-109 (312 B, align 8) -------------------------- */
typedef struct VPT_ALIGN(8) vpt_kernel_params {
    uint8_t      render;                    /*   0 bool */
    uint8_t      debug;                     /*   1 bool */
    uint8_t      _pad0[6];
    vpt_u2       resolution;                /*   8 */
    float        exposure_scale;            /*  16 */
    uint8_t      _pad1[4];
    vpt_devptr_t display_buffer;            /*  24 unsigned int*  */
    vpt_devptr_t raw_buffer;                /*  32 float4*        */
    vpt_devptr_t blue_noise_buffer;         /*  40 float3* (256x256) */
    vpt_devptr_t emission_texture;          /*  48 float3[256] blackbody LUT */
    float        emission_scale;            /*  56 */
    float        emission_pivot;            /*  60 */
    vpt_devptr_t density_color_texture;     /*  64 float3[256] */
    uint32_t     iteration;                 /*  72 */
    uint8_t      _pad2[4];
    vpt_devptr_t accum_buffer;              /*  80 float3* */
    vpt_devptr_t depth_buffer;              /*  88 float*  */
    uint32_t     max_interactions;          /*  96 */
    int32_t      ray_depth;                 /* 100 */
    int32_t      volume_depth;              /* 104 */
    float        min_extinction;            /* 108 */
    float        phase_g1;                  /* 112 */
    float        phase_g2;                  /* 116 */
    float        phase_f;                   /* 120 */
    vpt_f3       albedo;                    /* 124 */
    vpt_f3       extinction;                /* 136 */
    vpt_f3       transmittance;             /* 148 */
    float        tr_depth;                  /* 160 */
    float        density_mult;              /* 164 */
    uint32_t     environment_type;          /* 168 */
    float        azimuth;                   /* 172 */
    float        elevation;                 /* 176 */
    vpt_f3       sun_color;                 /* 180 */
    vpt_f3       sky_color;                 /* 192 */
    float        sun_mult;                  /* 204 */
    float        sky_mult;                  /* 208 */
    uint8_t      _pad3[4];
    double       energy_inject;             /* 216 */
    vpt_tex_t    env_tex;                   /* 224 */
    int32_t      env_sample_tex_res;        /* 232 */
    uint8_t      _pad4[4];
    vpt_tex_t    sky_tex;                   /* 240 */
    vpt_tex_t    env_func_tex;              /* 248 */
    vpt_tex_t    env_cdf_tex;               /* 256 */
    vpt_tex_t    env_marginal_func_tex;     /* 264 */
    vpt_tex_t    env_marginal_cdf_tex;      /* 272 */
    float        env_marginal_int;          /* 280 */
    uint8_t      _pad5[4];
    vpt_devptr_t debug_buffer;              /* 288 float3* */
    vpt_devptr_t cost_buffer;               /* 296 float3* */
    int32_t      integrator;                /* 304 */
    uint8_t      _pad6[4];
} vpt_kernel_params;
VPT_STATIC_ASSERT(sizeof(vpt_kernel_params) == 312, "Kernel_params is 312 B");
VPT_STATIC_ASSERT(offsetof(vpt_kernel_params, resolution) == 8 && offsetof(vpt_kernel_params, exposure_scale) == 16 &&
                  offsetof(vpt_kernel_params, display_buffer) == 24 && offsetof(vpt_kernel_params, blue_noise_buffer) == 40 &&
                  offsetof(vpt_kernel_params, emission_scale) == 56 && offsetof(vpt_kernel_params, density_color_texture) == 64 &&
                  offsetof(vpt_kernel_params, iteration) == 72 && offsetof(vpt_kernel_params, accum_buffer) == 80 &&
                  offsetof(vpt_kernel_params, max_interactions) == 96 && offsetof(vpt_kernel_params, volume_depth) == 104 &&
                  offsetof(vpt_kernel_params, phase_g1) == 112 && offsetof(vpt_kernel_params, albedo) == 124 &&
                  offsetof(vpt_kernel_params, extinction) == 136 && offsetof(vpt_kernel_params, tr_depth) == 160 &&
                  offsetof(vpt_kernel_params, environment_type) == 168 && offsetof(vpt_kernel_params, sun_color) == 180 &&
                  offsetof(vpt_kernel_params, sky_mult) == 208 && offsetof(vpt_kernel_params, energy_inject) == 216 &&
                  offsetof(vpt_kernel_params, env_tex) == 224 && offsetof(vpt_kernel_params, env_marginal_int) == 280 &&
                  offsetof(vpt_kernel_params, cost_buffer) == 296 && offsetof(vpt_kernel_params, integrator) == 304,
                  "Kernel_params field offsets");

/* Index of each entry in the `void* params[9]` array the reference passes to cuLaunchKernel
 * (source/main.cpp:1826): entry i points at the value of kernel parameter i. */
enum {
    VPT_ARG_CAMERA = 0,      /* -> vpt_camera            (by value)  */
    VPT_ARG_LIGHTS = 1,      /* -> vpt_light_list        (by value)  */
    VPT_ARG_VOLUMES = 2,     /* -> device ptr to vpt_gpu_vdb[N]      */
    VPT_ARG_SPHERE = 3,      /* -> device ptr to vpt_sphere          */
    VPT_ARG_GEO_LIST = 4,    /* -> device ptr to vpt_geometry_list (unused by live code) */
    VPT_ARG_BVH = 5,         /* -> device ptr to vpt_bvhnode[] (unused by live code)     */
    VPT_ARG_OCTREE = 6,      /* -> device ptr to the root vpt_octnode */
    VPT_ARG_ATMOSPHERE = 7,  /* -> vpt_atmosphere        (by value)  */
    VPT_ARG_KERNEL_PARAMS = 8,/* -> vpt_kernel_params     (by value)  */
    VPT_NUM_ARGS = 9
};

#endif /* VPT_ABI_H_ */
