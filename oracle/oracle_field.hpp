// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the reference's dense distance field (GridMap) in plain
// C++17, in the reference's operation order. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build or call this.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
// this path and cannot be compiled here (needs Eigen, ROS,
// PCL — none on disk). The restatement is pinned by analytic known-answer tests
// (tests/test_oracle_field.py) instead.
//
// Reference files restated here (paths relative to /root/reference):
//   src/map/src/grid_map.cpp:33-54      geometry derivation      -> Field::init
//   src/map/src/grid_map.cpp:89-123     fillESDF                 -> fill_esdf
//   src/map/src/grid_map.cpp:125-521    GridMap::updateESDF      -> Field::update_esdf
//   src/map/src/grid_map.cpp:719-747    regenerateMap ingest     -> Field::clear / rasterize
//   src/map/src/grid_map.cpp:800-809    loadMap                  -> Field::load_map
//   src/map/include/map/grid_map.h:256-509  queries              -> distance2d/3d, dis_with_grad_2d/3d
//   src/map/include/map/grid_map.h:613-650  isWholeBodyCollision -> whole_body_collision (oracle_opt.hpp)
//   src/map/include/map/grid_map.h:727-885  index helpers
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "../include/topay_b200.h"

namespace oracle {

// grid_map.cpp:89-123. f_get/f_set are callables on the line index. The
// arithmetic (integer q*q promoted to double, DBL_MAX sentinels, the s <= z[k]
// pop rule) follows the reference line by line.
template <typename FGet, typename FSet>
inline void fill_esdf(FGet f_get, FSet f_set, int start, int end, int size) {
    std::vector<int> v(size);
    std::vector<double> z(size + 1);
    int k = start;
    v[start] = start;
    z[start] = -std::numeric_limits<double>::max();
    z[start + 1] = std::numeric_limits<double>::max();
    for (int q = start + 1; q <= end; q++) {
        k++;
        double s;
        do {
            k--;
            s = ((f_get(q) + q * q) - (f_get(v[k]) + v[k] * v[k])) / (2 * q - 2 * v[k]);
        } while (s <= z[k]);
        k++;
        v[k] = q;
        z[k] = s;
        z[k + 1] = std::numeric_limits<double>::max();
    }
    k = start;
    for (int q = start; q <= end; q++) {
        while (z[k + 1] < q) k++;
        double val = (q - v[k]) * (q - v[k]) + f_get(v[k]);
        f_set(q, val);
    }
}

struct Field {
    // params (grid_map.h:81-90)
    double resolution = 0, resolution_inv = 0;
    double map_origin[3], map_size[3], min_boundary[3], max_boundary[3];
    int voxel_num[3], min_idx[3], max_idx[3];
    int buffer_size_2d = 0, buffer_size_3d = 0;
    double chassis_colli_radius = 0.4, chassis_height = 0.155;
    bool ready = false;
    // data (grid_map.h:96-102)
    std::vector<char> occ_buffer_3d, occ_buffer_2d, occ_buffer_2d_critical;
    std::vector<double> esdf_buffer_3d, esdf_buffer_2d, esdf_buffer_2d_inflate, esdf_buffer_2d_critical;
    // integer squared distances (cells^2) behind each map, kept for bit-exact
    // parity of the integer grids: index by TOPAY_MAP2D_* / TOPAY_MAP3D.
    // DBL_MAX in the reference -> INT32_MAX here.
    std::vector<int32_t> sq_pos[4], sq_neg[4];

    // grid_map.cpp:33-65
    void init(const topay_grid_desc& d) {
        for (int i = 0; i < 3; i++) map_size[i] = d.map_size[i];
        resolution = d.resolution;
        chassis_colli_radius = d.chassis_colli_radius;
        chassis_height = d.chassis_height;
        for (int i = 0; i < 3; i++) {
            min_boundary[i] = -map_size[i] / 2.0;
            max_boundary[i] = map_size[i] / 2.0;
        }
        min_boundary[2] = 0.0;
        max_boundary[2] = map_size[2];
        for (int i = 0; i < 3; i++) map_origin[i] = min_boundary[i];
        resolution_inv = 1.0 / resolution;
        for (int i = 0; i < 3; i++) {
            voxel_num[i] = (int)std::ceil(map_size[i] / resolution);
            min_idx[i] = 0;
            max_idx[i] = voxel_num[i] - 1;
        }
        buffer_size_2d = voxel_num[0] * voxel_num[1];
        buffer_size_3d = buffer_size_2d * voxel_num[2];
        esdf_buffer_2d.assign(buffer_size_2d, 0.0);
        esdf_buffer_2d_inflate.assign(buffer_size_2d, 0.0);
        esdf_buffer_2d_critical.assign(buffer_size_2d, 0.0);
        occ_buffer_2d.assign(buffer_size_2d, 0);
        occ_buffer_2d_critical.assign(buffer_size_2d, 0);
        esdf_buffer_3d.assign(buffer_size_3d, 0.0);
        occ_buffer_3d.assign(buffer_size_3d, 0);
        ready = false;
    }

    // ---- index helpers, grid_map.h:727-885 (dense path only) ----
    inline int addr2(int x, int y) const { return x * voxel_num[1] + y; }
    inline int addr3(int x, int y, int z) const {
        return x * voxel_num[1] * voxel_num[2] + y * voxel_num[2] + z;
    }
    inline void pos_to_index2(const double* p, int* id) const {
        id[0] = (int)std::floor((p[0] - map_origin[0]) * resolution_inv);
        id[1] = (int)std::floor((p[1] - map_origin[1]) * resolution_inv);
    }
    inline void pos_to_index3(const double* p, int* id) const {
        id[0] = (int)std::floor((p[0] - map_origin[0]) * resolution_inv);
        id[1] = (int)std::floor((p[1] - map_origin[1]) * resolution_inv);
        id[2] = (int)std::floor((p[2] - map_origin[2]) * resolution_inv);
    }
    inline void index_to_pos2(const int* id, double* p) const {
        p[0] = (id[0] + 0.5) * resolution + map_origin[0];
        p[1] = (id[1] + 0.5) * resolution + map_origin[1];
    }
    inline void index_to_pos3(const int* id, double* p) const {
        p[0] = (id[0] + 0.5) * resolution + map_origin[0];
        p[1] = (id[1] + 0.5) * resolution + map_origin[1];
        p[2] = (id[2] + 0.5) * resolution + map_origin[2];
    }
    inline void bound2(int* id) const {
        id[0] = std::max(std::min(id[0], max_idx[0]), min_idx[0]);
        id[1] = std::max(std::min(id[1], max_idx[1]), min_idx[1]);
    }
    inline void bound3(int* id) const {
        id[0] = std::max(std::min(id[0], max_idx[0]), min_idx[0]);
        id[1] = std::max(std::min(id[1], max_idx[1]), min_idx[1]);
        id[2] = std::max(std::min(id[2], max_idx[2]), min_idx[2]);
    }
    inline bool in_map2_pos(const double* p) const {
        if (p[0] < min_boundary[0] + 1e-4 || p[1] < min_boundary[1] + 1e-4) return false;
        if (p[0] > max_boundary[0] - 1e-4 || p[1] > max_boundary[1] - 1e-4) return false;
        return true;
    }
    inline bool in_map3_pos(const double* p) const {
        if (p[0] < min_boundary[0] + 1e-4 || p[1] < min_boundary[1] + 1e-4 ||
            p[2] < min_boundary[2] + 1e-4)
            return false;
        if (p[0] > max_boundary[0] - 1e-4 || p[1] > max_boundary[1] - 1e-4 ||
            p[2] > max_boundary[2] - 1e-4)
            return false;
        return true;
    }
    inline bool in_map2_idx(const int* id) const {
        if (id[0] < 0 || id[1] < 0) return false;
        if (id[0] > voxel_num[0] - 1 || id[1] > voxel_num[1] - 1) return false;
        return true;
    }
    inline bool in_map3_idx(const int* id) const {
        if (id[0] < 0 || id[1] < 0 || id[2] < 0) return false;
        if (id[0] > voxel_num[0] - 1 || id[1] > voxel_num[1] - 1 || id[2] > voxel_num[2] - 1)
            return false;
        return true;
    }

    // grid_map.cpp:719-722 / 758-761: occ_2d and occ_3d reset, critical kept.
    void clear(bool clear_critical) {
        esdf_buffer_2d.assign(buffer_size_2d, 0.0);
        occ_buffer_2d.assign(buffer_size_2d, 0);
        esdf_buffer_3d.assign(buffer_size_3d, 0.0);
        occ_buffer_3d.assign(buffer_size_3d, 0);
        if (clear_critical) occ_buffer_2d_critical.assign(buffer_size_2d, 0);
        ready = false;
    }

    // grid_map.cpp:733-747. pcl::PointXYZ carries float32; the reference builds
    // Eigen double vectors from the floats and compares z as float-vs-double.
    void rasterize(const float* xyz, int64_t n) {
        for (int64_t i = 0; i < n; i++) {
            const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
            double p2[2] = {(double)px, (double)py};
            int id[2];
            pos_to_index2(p2, id);
            if (in_map2_idx(id)) {
                occ_buffer_2d_critical[addr2(id[0], id[1])] = 1;
                if (pz < chassis_height) occ_buffer_2d[addr2(id[0], id[1])] = 1;
            }
            double p3[3] = {(double)px, (double)py, (double)pz};
            int id3[3];
            pos_to_index3(p3, id3);
            if (in_map3_idx(id3)) occ_buffer_3d[addr3(id3[0], id3[1], id3[2])] = 1;
        }
    }

    // grid_map.cpp:800-809
    void load_map(const char* occ2d, const char* occ3d) {
        occ_buffer_2d.assign(occ2d, occ2d + buffer_size_2d);
        occ_buffer_3d.assign(occ3d, occ3d + buffer_size_3d);
        esdf_buffer_2d.assign(buffer_size_2d, 0.0);
        esdf_buffer_3d.assign(buffer_size_3d, 0.0);
    }

    static inline int32_t sq_to_i32(double v) {
        return v >= 2147483647.0 ? INT32_MAX : (int32_t)v;
    }

    // One signed 2-D map, the block repeated four times in grid_map.cpp:138-423.
    // `is_src(x,y)` says whether the cell is a source of the positive transform;
    // the negative transform uses the complement. Pass order: y for each x, then
    // x for each y; then `d = pos; if (neg > 0) d += -neg + res` (grid_map.cpp:201-207).
    template <typename IsSrc>
    void signed_edt_2d(IsSrc is_src, std::vector<double>& out, int which) {
        const int rows = voxel_num[0], cols = voxel_num[1];
        const double DMAX = std::numeric_limits<double>::max();
        std::vector<char> src(buffer_size_2d);
        for (int x = 0; x < rows; x++)
            for (int y = 0; y < cols; y++) src[addr2(x, y)] = is_src(x, y) ? 1 : 0;
        std::vector<double> tmp(buffer_size_2d), dist(buffer_size_2d), neg(buffer_size_2d);
        sq_pos[which].assign(buffer_size_2d, 0);
        sq_neg[which].assign(buffer_size_2d, 0);
        for (int pass = 0; pass < 2; pass++) {
            const char want = pass == 0 ? 1 : 0;
            std::vector<double>& dst = pass == 0 ? dist : neg;
            std::vector<int32_t>& sq = pass == 0 ? sq_pos[which] : sq_neg[which];
            for (int x = min_idx[0]; x <= max_idx[0]; x++) {
                fill_esdf([&](int y) { return src[addr2(x, y)] == want ? 0.0 : DMAX; },
                          [&](int y, double val) { tmp[addr2(x, y)] = val; }, min_idx[1],
                          max_idx[1], cols);
            }
            for (int y = min_idx[1]; y <= max_idx[1]; y++) {
                fill_esdf([&](int x) { return tmp[addr2(x, y)]; },
                          [&](int x, double val) {
                              dst[addr2(x, y)] = resolution * std::sqrt(val);
                              sq[addr2(x, y)] = sq_to_i32(val);
                          },
                          min_idx[0], max_idx[0], rows);
            }
        }
        for (int x = min_idx[0]; x <= max_idx[0]; ++x)
            for (int y = min_idx[1]; y <= max_idx[1]; ++y) {
                const int a = addr2(x, y);
                out[a] = dist[a];
                if (neg[a] > 0.0) out[a] += (-neg[a] + resolution);
            }
    }

    // grid_map.cpp:125-521
    void update_esdf() {
        // 2d (flat, z < chassis_height) — :138-207
        signed_edt_2d([&](int x, int y) { return occ_buffer_2d[addr2(x, y)] == 1; },
                      esdf_buffer_2d, TOPAY_MAP2D_FLAT);
        // 2d critical — :209-279
        signed_edt_2d([&](int x, int y) { return occ_buffer_2d_critical[addr2(x, y)] == 1; },
                      esdf_buffer_2d_critical, TOPAY_MAP2D_CRITICAL);
        // 2d critical inflate, overwrites esdf_buffer_2d_critical — :281-351.
        // Sources are taken from the first critical map before it is overwritten.
        {
            std::vector<double> prev = esdf_buffer_2d_critical;
            signed_edt_2d([&](int x, int y) { return prev[addr2(x, y)] < chassis_colli_radius; },
                          esdf_buffer_2d_critical, TOPAY_MAP2D_CRITICAL);
        }
        // 2d inflate — :353-423
        signed_edt_2d([&](int x, int y) { return esdf_buffer_2d[addr2(x, y)] < chassis_colli_radius; },
                      esdf_buffer_2d_inflate, TOPAY_MAP2D_INFLATE);

        // 3d — :425-518. z pass, then y, then x; positive then complement.
        const double DMAX = std::numeric_limits<double>::max();
        std::vector<double> distance_buffer(buffer_size_3d, 10000.0);
        std::vector<double> distance_buffer_neg(buffer_size_3d, 10000.0);
        std::vector<double> tmp1(buffer_size_3d, 0.0), tmp2(buffer_size_3d, 0.0);
        sq_pos[TOPAY_MAP3D].assign(buffer_size_3d, 0);
        sq_neg[TOPAY_MAP3D].assign(buffer_size_3d, 0);
        for (int pass = 0; pass < 2; pass++) {
            const char want = pass == 0 ? 1 : 0;
            std::vector<double>& dst = pass == 0 ? distance_buffer : distance_buffer_neg;
            std::vector<int32_t>& sq = pass == 0 ? sq_pos[TOPAY_MAP3D] : sq_neg[TOPAY_MAP3D];
            for (int x = min_idx[0]; x <= max_idx[0]; x++)
                for (int y = min_idx[1]; y <= max_idx[1]; y++)
                    fill_esdf([&](int z) { return occ_buffer_3d[addr3(x, y, z)] == want ? 0.0 : DMAX; },
                              [&](int z, double val) { tmp1[addr3(x, y, z)] = val; }, min_idx[2],
                              max_idx[2], max_idx[2] + 1);
            for (int x = min_idx[0]; x <= max_idx[0]; x++)
                for (int z = min_idx[2]; z <= max_idx[2]; z++)
                    fill_esdf([&](int y) { return tmp1[addr3(x, y, z)]; },
                              [&](int y, double val) { tmp2[addr3(x, y, z)] = val; }, min_idx[1],
                              max_idx[1], max_idx[1] + 1);
            for (int y = min_idx[1]; y <= max_idx[1]; y++)
                for (int z = min_idx[2]; z <= max_idx[2]; z++)
                    fill_esdf([&](int x) { return tmp2[addr3(x, y, z)]; },
                              [&](int x, double val) {
                                  dst[addr3(x, y, z)] = resolution * std::sqrt(val);
                                  sq[addr3(x, y, z)] = sq_to_i32(val);
                              },
                              min_idx[0], max_idx[0], max_idx[0] + 1);
        }
        for (int i = 0; i < buffer_size_3d; i++) {
            esdf_buffer_3d[i] = distance_buffer[i];
            if (distance_buffer_neg[i] > 0.0) esdf_buffer_3d[i] += (-distance_buffer_neg[i] + resolution);
        }
        ready = true;
    }

    // grid_map.h:511-536
    bool is_collision2d(const double* pos, double threshold) const {
        if (in_map2_pos(pos)) {
            double d;
            distance2d(pos, d);
            return d < threshold;
        }
        return true;
    }
    // grid_map.h:695-724
    bool is_collision3d(const double* pos, double threshold) const {
        if (in_map3_pos(pos)) {
            double d;
            distance3d(pos, d);
            return d < threshold;
        }
        return true;
    }
    // grid_map.h:538-556 (dense branch)
    bool is_collision_idx2d(int x, int y, double threshold) const { return esdf_buffer_2d[addr2(x, y)] < threshold; }
    // grid_map.h:565-611: Bresenham over the flat map, both end cells included
    bool is_line_collision_grid2d(const double* p1, const double* p2, double threshold) const {
        int s[2], e[2];
        pos_to_index2(p1, s);
        pos_to_index2(p2, e);
        const int dx = std::abs(e[0] - s[0]), dy = std::abs(e[1] - s[1]);
        const int sx = s[0] < e[0] ? 1 : -1, sy = s[1] < e[1] ? 1 : -1;
        int err = dx - dy, x0 = s[0], y0 = s[1];
        for (;;) {
            if (!in_map2_idx2(x0, y0)) return true;   // deviation: the reference reads outside the buffer here
            if (is_collision_idx2d(x0, y0, threshold)) return true;
            if (x0 == e[0] && y0 == e[1]) break;
            const int e2 = 2 * err;
            if (e2 > -dy) { err -= dy; x0 += sx; }
            if (e2 < dx) { err += dx; y0 += sy; }
        }
        return false;
    }
    bool in_map2_idx2(int x, int y) const { return x >= 0 && y >= 0 && x <= max_idx[0] && y <= max_idx[1]; }
    // grid_map.h:887-940 (dense branch): nearest cell, clamped, critical or inflated map
    double dist_coarse2i(const int* id_in, bool critical) const {
        int id[2] = {id_in[0], id_in[1]};
        bound2(id);
        return critical ? esdf_buffer_2d_critical[addr2(id[0], id[1])] : esdf_buffer_2d_inflate[addr2(id[0], id[1])];
    }
    double dist_coarse2d(const double* pos, bool critical) const {
        int id[2];
        pos_to_index2(pos, id);
        return dist_coarse2i(id, critical);
    }

    // grid_map.h:256-305
    void distance2d(const double* pos, double& distance) const {
        if (!in_map2_pos(pos)) {
            distance = 1e+10;
            return;
        }
        double pos_m[2] = {pos[0] - 0.5 * resolution, pos[1] - 0.5 * resolution};
        int idx[2];
        pos_to_index2(pos_m, idx);
        double idx_pos[2];
        index_to_pos2(idx, idx_pos);
        double diff[2] = {pos[0] - idx_pos[0], pos[1] - idx_pos[1]};
        diff[0] *= resolution_inv;
        diff[1] *= resolution_inv;
        double values[2][2];
        for (int x = 0; x < 2; x++)
            for (int y = 0; y < 2; y++) {
                int c[2] = {idx[0] + x, idx[1] + y};
                bound2(c);
                values[x][y] = esdf_buffer_2d[addr2(c[0], c[1])];
            }
        double v0 = values[0][0] * (1 - diff[0]) + values[1][0] * diff[0];
        double v1 = values[0][1] * (1 - diff[0]) + values[1][1] * diff[0];
        distance = v0 * (1 - diff[1]) + v1 * diff[1];
    }

    // grid_map.h:307-362
    void distance3d(const double* pos, double& distance) const {
        if (!in_map3_pos(pos)) {
            distance = 1e+10;
            return;
        }
        double values[2][2][2], diff[3];
        gather3(pos, values, diff);
        double v00 = values[0][0][0] * (1 - diff[0]) + values[1][0][0] * diff[0];
        double v01 = values[0][0][1] * (1 - diff[0]) + values[1][0][1] * diff[0];
        double v10 = values[0][1][0] * (1 - diff[0]) + values[1][1][0] * diff[0];
        double v11 = values[0][1][1] * (1 - diff[0]) + values[1][1][1] * diff[0];
        double v0 = v00 * (1 - diff[1]) + v10 * diff[1];
        double v1 = v01 * (1 - diff[1]) + v11 * diff[1];
        distance = v0 * (1.0 - diff[2]) + v1 * diff[2];
    }

    // grid_map.h:364-441 (dense path)
    void dis_with_grad_2d(const double* pos, double& distance, double* grad, bool inflate = false,
                          bool critical = false) const {
        if (!in_map2_pos(pos)) {
            distance = 0.0;
            grad[0] = grad[1] = 0.0;
            return;
        }
        double pos_m[2] = {pos[0] - 0.5 * resolution, pos[1] - 0.5 * resolution};
        int idx[2];
        pos_to_index2(pos_m, idx);
        double idx_pos[2];
        index_to_pos2(idx, idx_pos);
        double diff[2] = {pos[0] - idx_pos[0], pos[1] - idx_pos[1]};
        diff[0] *= resolution_inv;
        diff[1] *= resolution_inv;
        const std::vector<double>& buf =
            critical ? esdf_buffer_2d_critical : (inflate ? esdf_buffer_2d_inflate : esdf_buffer_2d);
        double values[2][2];
        for (int x = 0; x < 2; x++)
            for (int y = 0; y < 2; y++) {
                int c[2] = {idx[0] + x, idx[1] + y};
                bound2(c);
                values[x][y] = buf[addr2(c[0], c[1])];
            }
        double v0 = values[0][0] * (1 - diff[0]) + values[1][0] * diff[0];
        double v1 = values[0][1] * (1 - diff[0]) + values[1][1] * diff[0];
        distance = v0 * (1 - diff[1]) + v1 * diff[1];
        grad[1] = (v1 - v0) * resolution_inv;
        grad[0] = (1 - diff[1]) * (values[1][0] - values[0][0]);
        grad[0] += diff[1] * (values[1][1] - values[0][1]);
        grad[0] *= resolution_inv;
    }

    // grid_map.h:472-490 shared by the two 3-D queries
    inline void gather3(const double* pos, double values[2][2][2], double* diff) const {
        double pos_m[3] = {pos[0] - 0.5 * resolution, pos[1] - 0.5 * resolution,
                           pos[2] - 0.5 * resolution};
        int idx[3];
        pos_to_index3(pos_m, idx);
        double idx_pos[3];
        index_to_pos3(idx, idx_pos);
        for (int i = 0; i < 3; i++) diff[i] = (pos[i] - idx_pos[i]) * resolution_inv;
        for (int x = 0; x < 2; x++)
            for (int y = 0; y < 2; y++)
                for (int z = 0; z < 2; z++) {
                    int c[3] = {idx[0] + x, idx[1] + y, idx[2] + z};
                    bound3(c);
                    values[x][y][z] = esdf_buffer_3d[addr3(c[0], c[1], c[2])];
                }
    }

    // grid_map.h:443-509 (dense path)
    void dis_with_grad_3d(const double* pos, double& distance, double* grad) const {
        if (!in_map3_pos(pos)) {
            distance = 0.0;
            grad[0] = grad[1] = grad[2] = 0.0;
            return;
        }
        double values[2][2][2], diff[3];
        gather3(pos, values, diff);
        double v00 = values[0][0][0] * (1 - diff[0]) + values[1][0][0] * diff[0];
        double v01 = values[0][0][1] * (1 - diff[0]) + values[1][0][1] * diff[0];
        double v10 = values[0][1][0] * (1 - diff[0]) + values[1][1][0] * diff[0];
        double v11 = values[0][1][1] * (1 - diff[0]) + values[1][1][1] * diff[0];
        double v0 = v00 * (1 - diff[1]) + v10 * diff[1];
        double v1 = v01 * (1 - diff[1]) + v11 * diff[1];
        distance = v0 * (1.0 - diff[2]) + v1 * diff[2];
        grad[2] = (v1 - v0) * resolution_inv;
        grad[1] = ((v10 - v00) * (1.0 - diff[2]) + (v11 - v01) * diff[2]) * resolution_inv;
        grad[0] = (1.0 - diff[2]) * (1 - diff[1]) * (values[1][0][0] - values[0][0][0]);
        grad[0] += (1.0 - diff[2]) * diff[1] * (values[1][1][0] - values[0][1][0]);
        grad[0] += diff[2] * (1 - diff[1]) * (values[1][0][1] - values[0][0][1]);
        grad[0] += diff[2] * diff[1] * (values[1][1][1] - values[0][1][1]);
        grad[0] *= resolution_inv;
    }

    // grid_map.h:511-536, 699-725 (dense path)
    bool is_collision_2d(const double* pos, double threshold) const {
        if (in_map2_pos(pos)) {
            double d;
            distance2d(pos, d);
            return d < threshold;
        }
        return true;
    }
    bool is_collision_3d(const double* pos, double threshold) const {
        if (in_map3_pos(pos)) {
            double d;
            distance3d(pos, d);
            return d < threshold;
        }
        return true;
    }
};

}  // namespace oracle
