// image_io.h -- the three small image decoders the render path's fixtures need.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace vpt {

// Radiance .hdr (RGBE, optional new-style RLE) -> float4 rows in file order, alpha = 0.
// Value rule follows source/hdr_loader.h:213-232: (mantissa + 0.5) * 2^(e - 136), e == 0 -> black.
bool load_hdr_float4(const std::string& path, std::vector<float>& rgba, unsigned& width, unsigned& height, std::string& err);

// 24-bit uncompressed BMP -> float3 rows top-down with the reference's channel order
// x = R/255, y = B/255, z = G/255 (source/util/fileIO.cpp:481-483, quirk Q16).
bool load_bmp_float3_rbg(const std::string& path, std::vector<float>& xyz, int& width, int& height, std::string& err);

// Uncompressed scan-line OpenEXR (HALF or FLOAT channels, must contain R, G, B) -> float3 per pixel,
// row-major (source/util/fileIO.cpp:356-390 reads the same files through OpenImageIO).
bool load_exr_float3(const std::string& path, std::vector<float>& rgb, int& width, int& height, std::string& err);

} // namespace vpt
