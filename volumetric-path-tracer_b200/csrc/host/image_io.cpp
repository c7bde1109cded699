// image_io.cpp -- Radiance HDR, BMP and (uncompressed) OpenEXR readers for the path's fixtures.
#include "image_io.h"

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <map>

namespace vpt {
namespace {

bool slurp(const std::string& path, std::vector<uint8_t>& d, std::string& err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open " + path; return false; }
    d.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

float half_to_float(uint16_t h) {
    const uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023;
    uint32_t out;
    if (e == 0) {
        if (m == 0) out = s << 31;
        else { int sh = 0; uint32_t mm = m; while (!(mm & 1024)) { mm <<= 1; ++sh; } out = (s << 31) | ((113 - sh) << 23) | ((mm & 1023) << 13); }
    } else if (e == 31) out = (s << 31) | 0x7f800000u | (m << 13);
    else out = (s << 31) | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &out, 4); return f;
}

} // namespace

bool load_hdr_float4(const std::string& path, std::vector<float>& rgba, unsigned& W, unsigned& H, std::string& err) {
    std::vector<uint8_t> d;
    if (!slurp(path, d, err)) return false;
    size_t p = 0;
    auto getline = [&](std::string& s) { s.clear(); while (p < d.size() && d[p] != '\n') s.push_back((char)d[p++]); if (p < d.size()) ++p; return p <= d.size(); };
    std::string line;
    getline(line);
    if (line.rfind("#?", 0) != 0) { err = "not a Radiance file"; return false; }
    W = H = 0;
    bool flipY = false;
    while (p < d.size()) {
        getline(line);
        if (line.empty() || line[0] == '#') continue;
        if (line.rfind("FORMAT=", 0) == 0) { if (line.find("32-bit_rle_rgbe") == std::string::npos) { err = "unsupported FORMAT"; return false; } continue; }
        if (line[0] == '-' || line[0] == '+') {
            char sy, ay, sx, ax; unsigned a, b;
            if (sscanf(line.c_str(), "%c%c %u %c%c %u", &sy, &ay, &a, &sx, &ax, &b) != 6 || ay != 'Y' || ax != 'X') { err = "unsupported resolution line"; return false; }
            H = a; W = b; flipY = (sy == '+');
            break;
        }
    }
    if (!W || !H) { err = "missing resolution"; return false; }
    if (W > 65536 || H > 65536) { err = "implausible resolution"; return false; }
    if (flipY) { err = "+Y orientation not supported"; return false; }
    rgba.assign(size_t(W) * H * 4, 0.0f);
    std::vector<uint8_t> row(size_t(W) * 4);
    const bool may_rle = !(W < 8 || W > 0x7fff);
    for (unsigned j = 0; j < H; ++j) {
        bool rle = false;
        if (may_rle && p + 4 <= d.size() && d[p] == 2 && d[p + 1] == 2 && !(d[p + 2] & 128)) {
            if (((unsigned)d[p + 2] << 8 | d[p + 3]) != W) { err = "bad RLE scanline length"; return false; }
            rle = true; p += 4;
        }
        if (!rle) {
            if (size_t(W) * 4 > d.size() - p) { err = "truncated pixels"; return false; }
            memcpy(row.data(), &d[p], size_t(W) * 4); p += size_t(W) * 4;
        } else {
            for (unsigned c = 0; c < 4; ++c) {
                for (unsigned pos = 0; pos < W;) {
                    if (p >= d.size()) { err = "truncated RLE"; return false; }
                    unsigned num = d[p++];
                    if (num > 128) {
                        num &= 127;
                        if (p >= d.size() || pos + num > W) { err = "bad RLE run"; return false; }
                        const uint8_t v = d[p++];
                        for (unsigned k = 0; k < num; ++k) row[(pos++) * 4 + c] = v;
                    } else {
                        if (num > d.size() - p || pos + num > W) { err = "bad RLE literal"; return false; }
                        for (unsigned k = 0; k < num; ++k) row[(pos++) * 4 + c] = d[p++];
                    }
                }
            }
        }
        for (unsigned i = 0; i < W; ++i) {
            const uint8_t* q = &row[size_t(i) * 4];
            float* o = &rgba[(size_t(j) * W + i) * 4];
            if (q[3] == 0) { o[0] = o[1] = o[2] = 0.0f; }
            else {
                uint32_t bits = (uint32_t(int(q[3]) - 9) << 23) & 0x7f800000u; float s; memcpy(&s, &bits, 4);
                o[0] = (float(q[0]) + 0.5f) * s; o[1] = (float(q[1]) + 0.5f) * s; o[2] = (float(q[2]) + 0.5f) * s;
            }
        }
    }
    return true;
}

bool load_bmp_float3_rbg(const std::string& path, std::vector<float>& xyz, int& W, int& H, std::string& err) {
    std::vector<uint8_t> d;
    if (!slurp(path, d, err)) return false;
    if (d.size() < 54 || d[0] != 'B' || d[1] != 'M') { err = "not a BMP"; return false; }
    uint32_t off; int32_t w, h; uint16_t bpp; uint32_t comp;
    memcpy(&off, &d[10], 4); memcpy(&w, &d[18], 4); memcpy(&h, &d[22], 4); memcpy(&bpp, &d[28], 2); memcpy(&comp, &d[30], 4);
    if (bpp != 24 || comp != 0 || w <= 0 || h == 0) { err = "only 24-bit uncompressed BMPs are supported"; return false; }
    const bool bottom_up = h > 0; if (h < 0) h = -h;
    const size_t stride = (size_t(w) * 3 + 3) & ~size_t(3);
    if (w > 65536 || h > 65536 || off > d.size() || stride * size_t(h) > d.size() - off) { err = "truncated BMP"; return false; }
    W = w; H = h; xyz.resize(size_t(w) * h * 3);
    for (int y = 0; y < h; ++y) {
        const uint8_t* src = &d[off + stride * size_t(bottom_up ? h - 1 - y : y)];
        for (int x = 0; x < w; ++x) {
            const uint8_t b = src[x * 3 + 0], g = src[x * 3 + 1], r = src[x * 3 + 2];
            float* o = &xyz[(size_t(y) * w + x) * 3];
            o[0] = float(r) / 255.0f; o[1] = float(b) / 255.0f; o[2] = float(g) / 255.0f;
        }
    }
    return true;
}

bool load_exr_float3(const std::string& path, std::vector<float>& rgb, int& W, int& H, std::string& err) {
    std::vector<uint8_t> d;
    if (!slurp(path, d, err)) return false;
    if (d.size() < 8 || d[0] != 0x76 || d[1] != 0x2f || d[2] != 0x31 || d[3] != 0x01) { err = "not an OpenEXR file"; return false; }
    uint32_t ver; memcpy(&ver, &d[4], 4);
    if (ver & 0x1E00) { err = "tiled / deep / multipart EXR not supported"; return false; }
    const size_t n = d.size();
    // every offset below comes from the file: nothing is dereferenced before it is checked against the file size
    auto fits = [&](size_t at, size_t k) { return at <= n && k <= n - at; };
    auto cstr = [&](size_t& at, std::string& out) {          // NUL-terminated string inside the file
        size_t e = at;
        while (e < n && d[e] != 0) ++e;
        if (e >= n) return false;
        out.assign((const char*)&d[at], e - at); at = e + 1;
        return true;
    };
    size_t p = 8;
    struct Ch { std::string name; int type; };
    std::vector<Ch> chans; int comp = -1; int32_t dw[4] = {0, 0, -1, -1};
    for (;;) {
        if (p >= n) { err = "truncated EXR header"; return false; }
        if (d[p] == 0) { ++p; break; }
        std::string name, type;
        if (!cstr(p, name) || !cstr(p, type) || !fits(p, 4)) { err = "truncated EXR attribute"; return false; }
        uint32_t sz; memcpy(&sz, &d[p], 4); p += 4;
        if (!fits(p, sz)) { err = "EXR attribute '" + name + "' runs past the end of the file"; return false; }
        const size_t end = p + sz;
        if (type == "chlist") {
            size_t q = p;
            for (;;) {
                if (q >= end) { err = "unterminated EXR channel list"; return false; }
                if (d[q] == 0) break;
                Ch c; size_t e = q;
                while (e < end && d[e] != 0) ++e;
                if (e >= end || end - (e + 1) < 16) { err = "truncated EXR channel entry"; return false; }
                c.name.assign((const char*)&d[q], e - q); q = e + 1;
                int32_t t; memcpy(&t, &d[q], 4); c.type = t; q += 16;
                if (c.type < 0 || c.type > 2) { err = "unknown EXR pixel type"; return false; }
                chans.push_back(c);
            }
        } else if (type == "compression") { if (sz < 1) { err = "bad EXR compression attribute"; return false; } comp = d[p]; }
        else if (name == "dataWindow") { if (sz < 16) { err = "bad EXR dataWindow attribute"; return false; } memcpy(dw, &d[p], 16); }
        p = end;
    }
    if (comp != 0) { err = "only uncompressed EXR is supported (compression=" + std::to_string(comp) + ")"; return false; }
    const int64_t w64 = (int64_t)dw[2] - dw[0] + 1, h64 = (int64_t)dw[3] - dw[1] + 1;
    if (w64 <= 0 || h64 <= 0 || w64 > 65536 || h64 > 65536) { err = "bad dataWindow"; return false; }
    W = (int)w64; H = (int)h64;
    if (!fits(p, size_t(H) * 8)) { err = "truncated EXR scanline offset table"; return false; }
    rgb.assign(size_t(W) * H * 3, 0.0f);
    std::vector<uint64_t> offs(H);
    memcpy(offs.data(), &d[p], size_t(H) * 8);
    for (int y = 0; y < H; ++y) {
        if (offs[y] > n || !fits((size_t)offs[y], 8)) { err = "EXR scanline offset outside the file"; return false; }
        size_t q = (size_t)offs[y];
        int32_t yy, bytes; memcpy(&yy, &d[q], 4); memcpy(&bytes, &d[q + 4], 4); q += 8;
        const int64_t row = (int64_t)yy - dw[1];
        if (row < 0 || row >= H) { err = "EXR scanline outside the dataWindow"; return false; }
        for (const Ch& c : chans) {                      // channels are stored in the chlist (alphabetical) order
            const int slot = c.name == "R" ? 0 : c.name == "G" ? 1 : c.name == "B" ? 2 : -1;
            const size_t esz = c.type == 1 ? 2 : 4;
            if (!fits(q, esz * W)) { err = "truncated EXR scanline"; return false; }
            if (slot >= 0) for (int x = 0; x < W; ++x) {
                float v;
                if (c.type == 1) { uint16_t hv; memcpy(&hv, &d[q + 2 * size_t(x)], 2); v = half_to_float(hv); }
                else if (c.type == 2) memcpy(&v, &d[q + 4 * size_t(x)], 4);
                else { uint32_t u; memcpy(&u, &d[q + 4 * size_t(x)], 4); v = float(u); }
                rgb[(size_t(row) * W + x) * 3 + slot] = v;
            }
            q += esz * W;
        }
    }
    return true;
}

} // namespace vpt
