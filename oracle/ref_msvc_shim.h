/* oracle/ref_msvc_shim.h -- TEST INFRASTRUCTURE (oracle side, never the product).
 * Force-included (-include) when the reference's atmosphere_kernels.cu is compiled with gcc as the host
 * compiler.  The reference is written against MSVC, which lets a temporary bind to the non-const
 * reference parameter of helper_math.h's unary minus (common/helper_math.h:273-295); gcc does not.
 * These const-reference overloads pick up exactly those rvalue uses and compute the same negation;
 * lvalue uses still resolve to the reference's own overloads.  No reference source is modified. */
#pragma once
#include <cuda_runtime.h>
inline __host__ __device__ float2 operator-(const float2& a) { return make_float2(-a.x, -a.y); }
inline __host__ __device__ float3 operator-(const float3& a) { return make_float3(-a.x, -a.y, -a.z); }
inline __host__ __device__ float4 operator-(const float4& a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }
