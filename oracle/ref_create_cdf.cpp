/*
 * ref_create_cdf.cpp -- TEST INFRASTRUCTURE ONLY (oracle side).
 *
 * Runs the reference's OWN create_cdf arithmetic: lines 681-751 of /root/reference/source/main.cpp are extracted at build time
 * (oracle/Makefile -> create_cdf_body.inc, never committed) and compiled here inside a function that supplies the variables the
 * block expects.  Two stand-ins, both outside the arithmetic under test:
 *   * `sample_atmosphere(...)` (the host-side analytic sky, a different function: vpt_env_sky_tabulate restates it) is replaced by a
 *     look-up into the caller's `func_in` table, returned as (v, 0, 0) so that `length()` reproduces v exactly;
 *   * the five arrays carry one guard element in front and one behind, because the block reads func[-1] and marginal_cdf[-1],
 *     writes cdf[-1] and writes marginal_cdf[res].  `guard` is the value found there.
 * Only tests/ may load this library.
 */
#include <cmath>
#include <cstring>
#include <vector>

struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
static inline float3 make_float3(float x, float y, float z) { float3 v = { x, y, z }; return v; }
static inline float length(float3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
struct KernelParamsStandIn { float3 sky_color; float env_marginal_int; int env_sample_tex_res; bool debug; };

static const float* g_func_in = nullptr;
#define sample_atmosphere(kp, p, d, col) make_float3(g_func_in[func_p - func], 0.0f, 0.0f)
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

extern "C" int vptref_create_cdf(const float* func_in, float* func_out, float* cdf_out, float* marginal_func_out, float* marginal_cdf_out,
                                 float* marginal_int_out, float guard)
{
    const unsigned res = 180;
    KernelParamsStandIn kernel_params; memset(&kernel_params, 0, sizeof(kernel_params));
    float3 pos = make_float3(0.0f, 0.0f, 0.0f);
    float az = 0, el = 0;
    std::vector<float3> val_store(res * res + 2);
    std::vector<float> func_store(res * res + 2, 0.0f), cdf_store(res * res + 2, 0.0f), mf_store(res + 2, 0.0f), mc_store(res + 2, 0.0f);
    memset(val_store.data(), 0, sizeof(float3) * val_store.size());
    func_store[0] = guard; cdf_store[0] = guard; mf_store[0] = guard; mc_store[0] = guard;
    float3* val = val_store.data() + 1, * val_p = val;
    float* func = func_store.data() + 1, * func_p = func;
    float* cdf = cdf_store.data() + 1, * cdf_p = cdf;
    float* marginal_func = mf_store.data() + 1, * marginal_func_p = marginal_func;
    float* marginal_cdf = mc_store.data() + 1, * marginal_cdf_p = marginal_cdf;
    g_func_in = func_in;

#include "create_cdf_body.inc"

    memcpy(func_out, func, sizeof(float) * res * res);
    memcpy(cdf_out, cdf, sizeof(float) * res * res);
    memcpy(marginal_func_out, marginal_func, sizeof(float) * res);
    memcpy(marginal_cdf_out, marginal_cdf, sizeof(float) * res);
    *marginal_int_out = kernel_params.env_marginal_int;
    return 0;
}
