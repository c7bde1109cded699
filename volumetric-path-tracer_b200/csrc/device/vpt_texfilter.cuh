// vpt_texfilter.cuh -- the texture unit's trilinear filter in software, shared by the brick sampler (vpt_trace_brick.cuh) and the cell-table
// sampler (vpt_trace.cuh).  (included inside `namespace vpt`)
#pragma once

// The texture unit's trilinear filter, measured on the device and reproduced here (tools/tex_filter_probe.py, tools/tex_weight_dump.py,
// tools/tex_weight_fit.py; profiles/r02d_tex_filter_probe.txt, profiles/r02f_tex_weight_fit.txt):
//  1. per axis the normalised coordinate is TRUNCATED to 21 fractional bits, U = floor(u * 2^21); the texel coordinate x = U * N / 2^21 - 0.5
//     is then exact (integer arithmetic) and its fraction is rounded half-up to 8 bits: A = floor(f * 256 + 0.5).  (tools/tex_coord_fit.py:
//     this reproduces 200 000 probed coordinates on each of 14 texture sizes from 3 to 2047 without a single miss; an exact u * N - 0.5
//     misses 0.4 % of them at N = 96 and 11.5 % at N = 2047, always one weight step high.)  A == 256 moves to the next cell with A = 0;
//     coordinates clamped at either edge get A = 0.
//  2. the EIGHT corner weights are integers that sum to 256, split hierarchically z -> x -> y:
//        Z1 = Az, Z0 = 256 - Az;   per z half T:  X1 = round_half_up(T * Ax / 256), X0 = T - X1;
//        x = 1 branch:  Y1 = round_half_up(X1 * Ay / 256), Y0 = X1 - Y1;     x = 0 branch:  Y1 = round_half_DOWN(X0 * Ay / 256), Y0 = X0 - Y1.
//     (so the x marginal misses round(f * 256) only at ties, the y marginal on a third of the samples -- which is why a per-axis
//     8-bit-weight emulation agrees with tex3D on ~0.2 % of the fetches, and this rule on 99.997 % of 60 000 probed weight sets)
//  3. the blend is the exactly-rounded sum of weight * texel (evaluated in double here: 8-bit x 24-bit products are exact): bit-identical to
//     tex3D<float> on 99.8 % of random fetches of a random-valued texture, within 1 ulp on the rest.
// kWeightMode: 0 = this rule (production), 1 = per-axis weights truncated to 1/256, 2 = full-precision per-axis weights (fp32 blend) --
// vpt_debug_sampler_compare reports all three against tex3D.
#ifndef VPT_BRICK_WEIGHT_MODE
#define VPT_BRICK_WEIGHT_MODE 0
#endif
constexpr int kBrickWeightMode = VPT_BRICK_WEIGHT_MODE;

struct TexCell { int i, j, k; float a, b, c; int A, B, C; };   // cell, per-axis weights as floats (modes 1, 2) and as 8-bit integers (mode 0)

template <int kWeightMode>
VPT_DEV void filter_axis(float u, int n, int& cell, float& w, int& W8)
{
    // U = floor(u * 2^21) is exact in fp32 arithmetic for u in [0, 1]: scaling by a power of two, then a floor
    const long long X = (long long)floorf(u * 2097152.0f) * (long long)n - (1ll << 20);       // (u21 * N - 0.5) in units of 2^-21
    cell = (int)(X >> 21);                                                                     // floor, also for negative X
    const int frac = (int)(X & ((1ll << 21) - 1));
    W8 = (frac + (1 << 12)) >> 13;                                                             // round half up to 8 bits
    const float f = (float)frac * (1.0f / 2097152.0f);
    if (kWeightMode == 0) w = (float)W8 * (1.0f / 256.0f);
    else if (kWeightMode == 1) w = floorf(f * 256.0f) * (1.0f / 256.0f);
    else w = f;
    if (W8 >= 256 || w >= 1.0f) { W8 = 0; w = 0.0f; ++cell; }          // a fraction that rounded up to 1 is the next cell
    if (cell < 0) { cell = 0; W8 = 0; w = 0.0f; }                      // clamp addressing: the whole weight on the edge texel
    if (cell >= n - 1) { cell = n - 1; W8 = 0; w = 0.0f; }
}

// texel cell and weights of a normalised, linearly filtered, clamp-addressed fetch at uvw
template <int kWeightMode>
VPT_DEV TexCell tex_cell(float3 uvw, int nx, int ny, int nz)
{
    TexCell q;
    filter_axis<kWeightMode>(uvw.x, nx, q.i, q.a, q.A);
    filter_axis<kWeightMode>(uvw.y, ny, q.j, q.b, q.B);
    filter_axis<kWeightMode>(uvw.z, nz, q.k, q.c, q.C);
    return q;
}

// blend of a cell's eight corner texels (v[z][y][x] order: v000, v100 = x+1, v010 = y+1, ...)
template <int kWeightMode>
VPT_DEV float tex_blend(const TexCell& q, float v000, float v100, float v010, float v110, float v001, float v101, float v011, float v111)
{
    if (kWeightMode == 0) {
        double acc = 0.0;
        #pragma unroll
        for (int zb = 0; zb < 2; ++zb) {
            const int T = zb ? q.C : 256 - q.C;
            const int X1 = (T * q.A + 128) >> 8, X0 = T - X1;
            const int Y11 = (X1 * q.B + 128) >> 8, Y10 = X1 - Y11;          // x = 1 branch: ties up
            const int Y01 = (X0 * q.B + 127) >> 8, Y00 = X0 - Y01;          // x = 0 branch: ties down
            acc = fma((double)Y00, (double)(zb ? v001 : v000), acc);
            acc = fma((double)Y10, (double)(zb ? v101 : v100), acc);
            acc = fma((double)Y01, (double)(zb ? v011 : v010), acc);
            acc = fma((double)Y11, (double)(zb ? v111 : v110), acc);
        }
        return (float)(acc * (1.0 / 256.0));
    }
    const float a = q.a, b = q.b, c = q.c, na = 1.0f - a, nb = 1.0f - b, nc = 1.0f - c;
    return na * nb * nc * v000 + a * nb * nc * v100 + na * b * nc * v010 + a * b * nc * v110
         + na * nb * c * v001 + a * nb * c * v101 + na * b * c * v011 + a * b * c * v111;
}

