// sky_table.cpp -- host-side tabulation of the sky's luminous power for the env sampling tables.
//
// The reference fills the tables its volumetric path integrator samples from (create_cdf, source/main.cpp:647-700) by
// evaluating, on the HOST, a small analytic sky: single-scattering Rayleigh + Mie along 16 view samples x 8 sun samples
// (`sample_atmosphere`, source/main.cpp:242-301, with `raySphereIntersect` :201-215 and `solveQuadratic` :181-198).
// This file restates that model in plain C++ so that `vpt_env_tables_create` can be fed without the reference application:
//     func[y * res + x] = | sky(origin, dir(az = x/(res-1) * 2pi, el = y/(res-1) * pi)) * sky_color |      (main.cpp:685-693)
// Pure CPU code, fp32 like the reference's.
#include <cfloat>
#include <cmath>
#include <cstddef>

namespace {

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline float clampf(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }
const float kPi = 3.14159265358979323846f;

bool solve_quadratic(float a, float b, float c, float& x1, float& x2) {        // main.cpp:181-198
    if (b == 0) {
        if (a == 0) return false;
        x1 = 0; x2 = sqrtf(-c / a);
        return true;
    }
    const float discr = b * b - 4 * a * c;
    if (discr < 0) return false;
    const float q = (b < 0.f) ? -0.5f * (b - sqrtf(discr)) : -0.5f * (b + sqrtf(discr));
    x1 = q / a; x2 = c / q;
    return true;
}

bool ray_sphere(V3 orig, V3 dir, float radius, float& t0, float& t1) {           // main.cpp:201-215
    const float A = dot(dir, dir), B = 2 * dot(dir, orig), C = dot(orig, orig) - radius * radius;
    if (!solve_quadratic(A, B, C, t0, t1)) return false;
    if (t0 > t1) { const float t = t1; t1 = t0; t0 = t; }
    return true;
}

V3 sun_direction(float azimuth, float elevation) {                              // host degree_to_cartesian, main.cpp:224-238 (elevation clamped to [0, 90])
    float az = clampf(azimuth, .0f, 360.0f), el = clampf(elevation, .0f, 90.0f);
    az = az * kPi / 180.0f; el = (90.0f - el) * kPi / 180.0f;
    V3 d{sinf(el) * cosf(az), cosf(el), sinf(el) * sinf(az)};
    return d * (1.0f / length(d));
}

V3 analytic_sky(V3 orig, V3 dir, V3 sun, V3 intensity) {                         // main.cpp:242-301
    const float atmosphere_radius = 6420e3f, earth_radius = 6360e3f, Hr = 7994.0f, Hm = 1200.0f;
    const V3 betaR{3.8e-6f, 13.5e-6f, 33.1e-6f}, betaM{21e-6f, 21e-6f, 21e-6f};
    float t0, t1, tmin, tmax = FLT_MAX;
    V3 pos = orig; pos.y += 1000 + 6360e3f;
    if (ray_sphere(pos, dir, earth_radius, t0, t1) && t1 > .0f) tmax = fmaxf(.0f, t0);
    tmin = .0f;
    if (!ray_sphere(pos, dir, atmosphere_radius, t0, t1) || t1 < 0) return V3{1.0f, .0f, .0f};
    if (t0 > tmin && t0 > 0) tmin = t0;
    if (t1 < tmax) tmax = t1;
    const unsigned n_view = 16, n_light = 8;
    const float segment = (tmax - tmin) / n_view;
    float t_cur = tmin;
    V3 sumR{0, 0, 0}, sumM{0, 0, 0};
    float depthR = 0, depthM = 0;
    const float mu = dot(dir, sun);
    const float phaseR = 3.f / (16.f * kPi) * (1 + mu * mu);
    const float g = 0.76f;
    const float phaseM = 3.f / (8.f * kPi) * ((1.f - g * g) * (1.f + mu * mu)) / ((2.f + g * g) * powf(1.f + g * g - 2.f * g * mu, 1.5f));
    for (unsigned i = 0; i < n_view; ++i) {
        const V3 p = pos + dir * (t_cur + segment * 0.5f);
        const float height = length(p) - earth_radius;
        const float hr = expf(-height / Hr) * segment, hm = expf(-height / Hm) * segment;
        depthR += hr; depthM += hm;
        float t0l, t1l;
        ray_sphere(p, sun, atmosphere_radius, t0l, t1l);
        const float seg_l = t1l / n_light; float t_l = 0, dlr = 0, dlm = 0;
        unsigned j;
        for (j = 0; j < n_light; ++j) {
            const V3 pl = p + sun * (t_l + seg_l * 0.5f);
            const float hl = length(pl) - earth_radius;
            if (hl < 0) break;
            dlr += expf(-hl / Hr) * seg_l; dlm += expf(-hl / Hm) * seg_l;
            t_l += seg_l;
        }
        if (j == n_light) {
            const V3 tau = betaR * (depthR + dlr) + betaM * 1.1f * (depthM + dlm);
            const V3 att{expf(-tau.x), expf(-tau.y), expf(-tau.z)};
            sumR = sumR + att * hr; sumM = sumM + att * hm;
        }
        t_cur += segment;
    }
    return (sumR * betaR * phaseR + sumM * betaM * phaseM) * intensity;
}

} // namespace

extern "C" int vpt_env_sky_tabulate(float azimuth_deg, float elevation_deg, const float sky_color[3], unsigned res, float* func_out) {
    if (!sky_color || !func_out || res < 2) return -1;                           // VPT_ERR_INVALID
    const V3 sun = sun_direction(azimuth_deg, elevation_deg);
    const V3 intensity{sky_color[0], sky_color[1], sky_color[2]};
    for (unsigned y = 0; y < res; ++y) {
        const float el = float(y) / float(res - 1) * kPi;                        // 0 .. 180 degrees
        for (unsigned x = 0; x < res; ++x) {
            const float az = float(x) / float(res - 1) * kPi * 2.0f;             // 0 .. 360 degrees
            const V3 dir{sinf(el) * cosf(az), cosf(el), sinf(el) * sinf(az)};
            func_out[(size_t)y * res + x] = length(analytic_sky(V3{0.f, 0.f, 0.f}, dir, sun, intensity));
        }
    }
    return 0;
}
