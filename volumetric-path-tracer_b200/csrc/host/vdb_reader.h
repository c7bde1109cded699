// vdb_reader.h -- minimal VDB-224 -> dense grid reader (see vdb_reader.cpp).
#pragma once
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>

namespace vpt {

struct VdbGridMeta {
    bool     has_file_bbox = false;
    int32_t  file_bbox_min[3] = {0, 0, 0}, file_bbox_max[3] = {0, 0, 0};
    int64_t  file_voxel_count = -1;
    bool     half_float = false;
};

struct VdbDenseGrid {
    int32_t  channels = 1;               // 1 (float) or 3 (vec3s)
    int32_t  bbox_min[3], bbox_max[3];   // inclusive active-voxel bounding box (index space)
    int32_t  dim[3];
    float    background[3];
    double   index_to_world[16];         // OpenVDB Mat4d, row-vector convention (translation in row 3)
    double   voxel_size[3];
    uint32_t leaf_count = 0, active_tiles = 0;
    uint64_t active_leaf_voxels = 0, active_tile_voxels = 0;
    VdbGridMeta meta;
    std::vector<float> values;           // x fastest: idx = (z*dim.y + y)*dim.x + x, `channels` floats each
};

std::vector<std::string> vdb_list_grids(const uint8_t* data, size_t n);
// Returns false when the file has no grid called `grid_name`; throws std::runtime_error on malformed input.
bool vdb_read_dense(const uint8_t* data, size_t n, const std::string& grid_name, VdbDenseGrid& out);

} // namespace vpt
