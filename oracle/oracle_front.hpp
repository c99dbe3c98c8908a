// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the front-end pieces on either side of the solve ("next" row N2 of SURVEY.md §8f):
//   GraphSearch::getDensePath      src/planner/src/graph_search.cpp:119-176 (+ normalizeAngle :6-13)
//   RayCaster::setInput / step     src/planner/src/utils/raycast.cpp:27-45, 253-346
//   TopologyPRM::lineVisib         src/planner/src/topo_prm.cpp:278-315
//   TopologyPRM::sameTopoPath      src/planner/src/topo_prm.cpp:424-448 (+ pathLength :462-470, discretizePath :472-506)
// in the reference's operation order. Pinned bit for bit against the reference's own code compiled unmodified
// (oracle/_ref, tests/test_ref_pin.py::test_dense_path_bit_exact / test_line_visib_bit_exact).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <vector>

#include "oracle_field.hpp"

namespace oracle {

// graph_search.cpp:6-13
inline void normalize_angle(double ref_angle, double& angle) {
    while (ref_angle - angle > M_PI) angle += 2 * M_PI;
    while (ref_angle - angle < -M_PI) angle -= 2 * M_PI;
}

// graph_search.cpp:119-176: raw 2-D waypoints -> (x, y, theta, dt) rows: every segment cut into ceil(len / step)
// pieces, a turn-in-place row before and after every move, rows shorter than 1e-3 s dropped.
inline std::vector<std::array<double, 4>> dense_path(const double* raw, int n, double step_size, double start_yaw,
                                                     double end_yaw, double v_max, double w_max) {
    std::vector<std::array<double, 2>> dense;
    dense.push_back({raw[0], raw[1]});
    for (int i = 1; i < n; i++) {
        const double ex = raw[2 * i] - raw[2 * i - 2], ey = raw[2 * i + 1] - raw[2 * i - 1];
        // Eigen: normalized() = v / sqrt(squaredNorm) when squaredNorm > 0, norm() = sqrt(squaredNorm)
        const double z = ex * ex + ey * ey;
        double dx = ex, dy = ey;
        if (z > 0.0) {
            const double nrm = std::sqrt(z);
            dx = ex / nrm;
            dy = ey / nrm;
        }
        const double len = std::sqrt(z);
        const int times = (int)std::max(std::ceil(len / step_size), 1.0);
        const double step = len / times;
        for (int j = 1; j <= times; j++)
            dense.push_back({raw[2 * i - 2] + (step * j) * dx, raw[2 * i - 1] + (step * j) * dy});
    }
    auto norm2 = [](double a, double b) { return std::sqrt(a * a + b * b); };
    std::vector<std::array<double, 4>> sp;
    sp.push_back({dense[0][0], dense[0][1], start_yaw, 0.0});
    double cur = std::atan2(dense[1][1] - dense[0][1], dense[1][0] - dense[0][0]);
    normalize_angle(start_yaw, cur);
    sp.back()[3] = std::fabs(cur - start_yaw) / w_max;
    sp.push_back({dense[0][0], dense[0][1], cur, 0.0});
    for (size_t i = 1; i + 1 < dense.size(); i++) {
        const double px = dense[i][0], py = dense[i][1];
        const double arc = norm2(px - sp.back()[0], py - sp.back()[1]);
        sp.back()[3] = arc / v_max;
        sp.push_back({px, py, sp.back()[2], 0.0});
        cur = std::atan2(dense[i + 1][1] - dense[i][1], dense[i + 1][0] - dense[i][0]);
        normalize_angle(sp.back()[2], cur);
        sp.back()[3] = std::fabs(cur - sp.back()[2]) / w_max;
        sp.push_back({px, py, cur, 0.0});
    }
    const double px = dense.back()[0], py = dense.back()[1];
    sp.back()[3] = norm2(px - sp.back()[0], py - sp.back()[1]) / v_max;
    sp.push_back({px, py, sp.back()[2], 0.0});
    cur = end_yaw;
    normalize_angle(sp.back()[2], cur);
    sp.back()[3] = std::fabs(cur - sp.back()[2]) / w_max;
    sp.push_back({px, py, cur, 0.0});
    std::vector<std::array<double, 4>> out;
    for (size_t i = 0; i + 1 < sp.size(); i++)
        if (sp[i][3] > 1.0e-3) out.push_back(sp[i]);
    out.push_back(sp.back());
    return out;
}

// raycast.cpp:27-45, 253-346 (Amanatides & Woo voxel walk in cell units)
struct PlannerRay {
    int x, y, z, ex, ey, ez, sx, sy, sz;
    double tmx, tmy, tmz, tdx, tdy, tdz;
    static int signum(int v) { return v == 0 ? 0 : v < 0 ? -1 : 1; }
    static double mod(double value, double modulus) { return std::fmod(std::fmod(value, modulus) + modulus, modulus); }
    static double intbound(double s, double ds) {
        if (ds < 0) return intbound(-s, -ds);
        s = mod(s, 1);
        return (1 - s) / ds;
    }
    bool set_input(const double* s, const double* e) {
        x = (int)std::floor(s[0]); y = (int)std::floor(s[1]); z = (int)std::floor(s[2]);
        ex = (int)std::floor(e[0]); ey = (int)std::floor(e[1]); ez = (int)std::floor(e[2]);
        const double dx = ex - x, dy = ey - y, dz = ez - z;
        sx = signum((int)dx); sy = signum((int)dy); sz = signum((int)dz);
        tmx = intbound(s[0], dx); tmy = intbound(s[1], dy); tmz = intbound(s[2], dz);
        tdx = ((double)sx) / dx; tdy = ((double)sy) / dy; tdz = ((double)sz) / dz;
        return !(sx == 0 && sy == 0 && sz == 0);
    }
    // the current cell; false when it is the end cell (which the callers therefore never test)
    bool step(int* cell) {
        cell[0] = x; cell[1] = y; cell[2] = z;
        if (x == ex && y == ey && z == ez) return false;
        if (tmx < tmy) {
            if (tmx < tmz) { x += sx; tmx += tdx; } else { z += sz; tmz += tdz; }
        } else {
            if (tmy < tmz) { y += sy; tmy += tdy; } else { z += sz; tmz += tdz; }
        }
        return true;
    }
};

// topo_prm.cpp:278-315. Returns visible; pc = midpoint of the blocking cell's centre and the previous cell's (z = 0).
// `max_steps` caps the walk (the reference loops for ever on a ray that steps past its end cell; stated deviation,
// never reached by rays whose ends lie in the cells they were built from).
inline bool line_visib(const Field& f, const double* p1, const double* p2, double thresh, bool use_critical, double* pc,
                       long max_steps = 1L << 20) {
    const double res = f.resolution;
    const double off[3] = {0.5 - f.map_origin[0] / res, 0.5 - f.map_origin[1] / res, 0.5 - f.map_origin[2] / res};
    const double s[3] = {p1[0] / res, p1[1] / res, p1[2] / res}, e[3] = {p2[0] / res, p2[1] / res, p2[2] / res};
    PlannerRay rc;
    const bool flag = rc.set_input(s, e);
    int prev[3] = {(int)std::floor(s[0]), (int)std::floor(s[1]), (int)std::floor(s[2])};
    int cell[3];
    long steps = 0;
    while (flag && rc.step(cell) && steps++ < max_steps) {
        int id[3] = {(int)(cell[0] + off[0]), (int)(cell[1] + off[1]), 0};   // double -> int truncation, as the reference
        const double dist = f.dist_coarse2i(id, use_critical);
        if (dist <= thresh) {
            double c1[3], c2[3];
            f.index_to_pos3(id, c1);
            f.index_to_pos3(prev, c2);
            pc[0] = 0.5 * (c1[0] + c2[0]);
            pc[1] = 0.5 * (c1[1] + c2[1]);
            pc[2] = 0.0;
            return false;
        }
        prev[0] = id[0]; prev[1] = id[1]; prev[2] = id[2];
    }
    return true;
}

// topo_prm.cpp:462-470
inline double path_length(const double* path, int n) {
    double length = 0.0;
    if (n < 2) return length;
    for (int i = 0; i < n - 1; ++i) {
        const double dx = path[3 * i + 3] - path[3 * i], dy = path[3 * i + 4] - path[3 * i + 1], dz = path[3 * i + 5] - path[3 * i + 2];
        length += std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    return length;
}
// topo_prm.cpp:472-506: pt_num points at equal arc-length spacing along the polyline
inline std::vector<std::array<double, 3>> discretize_path(const double* path, int n, int pt_num) {
    std::vector<double> len_list;
    len_list.push_back(0.0);
    for (int i = 0; i < n - 1; ++i) {
        const double dx = path[3 * i + 3] - path[3 * i], dy = path[3 * i + 4] - path[3 * i + 1], dz = path[3 * i + 5] - path[3 * i + 2];
        len_list.push_back(std::sqrt(dx * dx + dy * dy + dz * dz) + len_list[i]);
    }
    const double len_total = len_list.back();
    const double dl = len_total / double(pt_num - 1);
    std::vector<std::array<double, 3>> out;
    for (int i = 0; i < pt_num; ++i) {
        const double cur_l = double(i) * dl;
        int idx = -1;
        for (int j = 0; j < (int)len_list.size() - 1; ++j)
            if (cur_l >= len_list[j] - 1e-4 && cur_l <= len_list[j + 1] + 1e-4) {
                idx = j;
                break;
            }
        const double lambda = (cur_l - len_list[idx]) / (len_list[idx + 1] - len_list[idx]);
        out.push_back({(1 - lambda) * path[3 * idx] + lambda * path[3 * idx + 3],
                       (1 - lambda) * path[3 * idx + 1] + lambda * path[3 * idx + 4],
                       (1 - lambda) * path[3 * idx + 2] + lambda * path[3 * idx + 5]});
    }
    return out;
}
// topo_prm.cpp:424-448: two paths are the same topological class when every pair of equally spaced points sees
// each other. (Callers pass paths longer than one cell: pt_num >= 2.)
inline bool same_topo_path(const Field& f, const double* p1, int n1, const double* p2, int n2, double thresh,
                           bool use_critical) {
    const double max_len = std::max(path_length(p1, n1), path_length(p2, n2));
    const int pt_num = (int)std::ceil(max_len / f.resolution);
    const auto a = discretize_path(p1, n1, pt_num), b = discretize_path(p2, n2, pt_num);
    double pc[3];
    for (int i = 0; i < pt_num; ++i)
        if (!line_visib(f, a[i].data(), b[i].data(), thresh, use_critical, pc)) return false;
    return true;
}

}  // namespace oracle
