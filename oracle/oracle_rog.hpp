// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the reference's ROG-Map distance field (ESDFMap on top of
// CounterMap / SlidingMap): ring-buffer index arithmetic, the local-box signed
// EDT with its wrap rules, the two 2-D maps and the query functions. Plain
// C++17, the reference's operation order, ORIGIN_AT_CORNER discretisation
// (src/rog_map/CMakeLists.txt:14). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build or call this.
//
// PARITY UNPINNED: the reference has no tests or fixtures for this path and
// cannot be compiled here (Eigen, ROS, PCL absent). Pinned by brute-force and
// analytic known-answer tests in tests/test_oracle_rog.py.
//
// Reference files restated here (paths relative to /root/reference/src/rog_map):
//   src/rog_map/sliding_map.cpp:44-63    initSlidingMap           -> RogEsdf::init
//   src/rog_map/sliding_map.cpp:85-97    updateLocalMapOriginAndBound
//   src/rog_map/sliding_map.cpp:99-166   clearMemoryOutOfMap / mapSliding -> RogEsdf::slide
//   src/rog_map/sliding_map.cpp:168-301  index helpers            -> pos_to_global ... hash_from_pos
//   src/rog_map/counter_map.cpp:31-91    initCounterMap (sizes, sub_grid_num, unk_thresh)
//   src/rog_map/counter_map.cpp:94-151   updateGridCounter        -> RogEsdf::update_counter
//   src/rog_map/esdf_map.cpp:28-57       initESDFMap
//   src/rog_map/esdf_map.cpp:78-120      nearest-cell getters
//   src/rog_map/esdf_map.cpp:122-152     isLineFree2d  (+ include/utils/raycaster.cpp:66-192)
//   src/rog_map/esdf_map.cpp:154-500     updateESDF3D             -> RogEsdf::update_esdf
//   src/rog_map/esdf_map.cpp:842-900     fillESDF with ring wrap  -> RogEsdf::fill_line
//   src/rog_map/esdf_map.cpp:903-1100    evaluateEDT / evaluateFirstGrad / getValueGrad /
//                                        getCriticalValueGrad / getValueGrad2d
//
// Deviations, all stated: (1) the local-map bounds are initialised from the
// initial origin in init() — the reference leaves them unset until the first
// mapSliding(); (2) the 2-D combine loop, which walks y over the x range
// (esdf_map.cpp:391-398), is clipped to the buffer so that a non-square map
// cannot run off the end of the vector (undefined behaviour in the reference); (3) isLineFree2d's
// walk is capped (see is_line_free_2d); (4) update_esdf returns at once when the clipped update
// box is empty on any axis.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <array>
#include <cstdlib>
#include <limits>
#include <vector>

namespace oracle {

enum RogGridType : int { ROG_UNDEFINED = 0, ROG_UNKNOWN = 1, ROG_OUT_OF_MAP = 2, ROG_OCCUPIED = 3, ROG_KNOWN_FREE = 4 };

struct RogEsdf {
    // --- SlidingMap::sc_ / CounterMap::md_
    double res = 0, res_inv = 0;
    int half[3] = {0, 0, 0}, size[3] = {0, 0, 0};
    int64_t vox = 0;
    bool sliding_en = true;
    int origin_i[3] = {0, 0, 0}, bmin_i[3] = {0, 0, 0}, bmax_i[3] = {0, 0, 0};
    int sub_grid_num = 1, unk_thresh = 1;
    std::vector<int16_t> occupied_cnt, unknown_cnt;
    // --- ESDFMap
    std::vector<double> dist3, dist_crit, dist_flat, tmp1, tmp2;
    int half_box_i[3] = {0, 0, 0};
    int upd_min_i[3] = {0, 0, 0}, upd_max_i[3] = {0, 0, 0};

    static int ifloor(double v) { return (int)std::floor(v); }

    // sliding_map.cpp:175-184 (ORIGIN_AT_CORNER)
    void pos_to_global(const double p[3], int id[3]) const {
        for (int i = 0; i < 3; i++) id[i] = ifloor(p[i] * res_inv);
    }
    // sliding_map.cpp:195-203
    void global_to_pos(const int id[3], double p[3]) const {
        for (int i = 0; i < 3; i++) p[i] = ((double)id[i] + 0.5) * res;
    }
    // sliding_map.cpp:205-218
    void global_to_local(const int g[3], int l[3]) const {
        for (int i = 0; i < 3; i++) {
            int v = g[i] % size[i];
            if (v > half[i]) v -= size[i];
            else if (v < -half[i]) v += size[i];
            l[i] = v;
        }
    }
    // sliding_map.cpp:220-232
    void local_to_global(const int l[3], int g[3]) const {
        for (int i = 0; i < 3; i++) {
            const int min_g = -half[i] + origin_i[i];
            int min_l = min_g % size[i];
            if (min_l > half[i]) min_l -= size[i];
            if (min_l < -half[i]) min_l += size[i];
            int d = l[i] - min_l;
            if (d < 0) d += size[i];
            g[i] = d + min_g;
        }
    }
    // sliding_map.cpp:168-173
    int64_t local_hash(const int l[3]) const {
        return (int64_t)(l[0] + half[0]) * size[1] * size[2] + (int64_t)(l[1] + half[1]) * size[2] + (l[2] + half[2]);
    }
    int64_t hash_from_global(const int g[3]) const {
        int l[3];
        global_to_local(g, l);
        return local_hash(l);
    }
    int64_t hash_from_pos(const double p[3]) const {
        int g[3];
        pos_to_global(p, g);
        return hash_from_global(g);
    }
    // the (x, y) plane index the 2-D getters use (esdf_map.cpp:86-120)
    int64_t hash2_from_pos(const double p[3]) const {
        int g[3], l[3];
        pos_to_global(p, g);
        global_to_local(g, l);
        return (int64_t)(l[0] + half[0]) * size[1] + (l[1] + half[1]);
    }
    // sliding_map.cpp:71-76
    bool inside_local_map(const int g[3]) const {
        for (int i = 0; i < 3; i++)
            if (std::abs(g[i] - origin_i[i]) - half[i] > 0) return false;
        return true;
    }

    void set_origin(const int o[3]) {
        for (int i = 0; i < 3; i++) {
            origin_i[i] = o[i];
            bmax_i[i] = o[i] + half[i];
            bmin_i[i] = o[i] - half[i];
        }
    }

    // counter_map.cpp:31-91 + esdf_map.cpp:28-57. inflation_step = 0 for the ESDF counter map.
    void init(const int half_prob[3], double prob_res, double counter_res_in, const double local_update_box[3],
              bool map_sliding_en, const double fix_origin[3], double unk_thresh_ratio) {
        const int ratio = (int)std::round(counter_res_in / prob_res);
        const double cres = prob_res * ratio;
        for (int i = 0; i < 3; i++) {
            const double half_d = (double)half_prob[i] * prob_res;
            half[i] = (int)(half_d / cres) + 1;
            size[i] = 2 * half[i] + 1;
        }
        res = cres;
        res_inv = 1.0 / cres;
        sliding_en = map_sliding_en;
        vox = (int64_t)size[0] * size[1] * size[2];
        int o[3] = {0, 0, 0};
        if (!map_sliding_en) pos_to_global(fix_origin, o);
        set_origin(o);
        sub_grid_num = (int)std::pow(std::round(cres / prob_res), 3);
        unk_thresh = (int)std::ceil(unk_thresh_ratio * sub_grid_num);
        unk_thresh = std::min(std::max(1, unk_thresh), sub_grid_num);
        unknown_cnt.assign(vox, (int16_t)sub_grid_num);
        occupied_cnt.assign(vox, 0);
        dist3.assign(vox, 0.0);
        tmp1.assign(vox, 0.0);
        tmp2.assign(vox, 0.0);
        dist_crit.assign((size_t)size[0] * size[1], 0.0);
        dist_flat.assign((size_t)size[0] * size[1], 0.0);
        int b[3];
        pos_to_global(local_update_box, b);
        for (int i = 0; i < 3; i++) half_box_i[i] = b[i] / 2;
    }

    // esdf_map.cpp:72-76
    void reset_local_map() {
        std::fill(unknown_cnt.begin(), unknown_cnt.end(), (int16_t)sub_grid_num);
        std::fill(occupied_cnt.begin(), occupied_cnt.end(), (int16_t)0);
    }

    // sliding_map.cpp:113-166 (+ :99-111); counter_map.h:127-131 resets the two counters.
    void slide(const double odom[3]) {
        int no[3];
        pos_to_global(odom, no);
        int shift[3];
        for (int i = 0; i < 3; i++) shift[i] = no[i] - origin_i[i];
        for (int i = 0; i < 3; i++)
            if (std::fabs((double)shift[i]) > size[i]) {
                reset_local_map();
                set_origin(no);
                return;
            }
        auto normalize = [](int x, int a, int b) {
            const int range = b - a + 1;
            const int y = (x - a) % range;
            return (y < 0 ? y + range : y) + a;
        };
        for (int i = 0; i < 3; i++) {
            if (shift[i] == 0) continue;
            const int min_g = -half[i] + origin_i[i];
            const int min_l = min_g % size[i];
            std::vector<int> clear_id;
            if (shift[i] > 0)
                for (int k = 0; k < shift[i]; k++) clear_id.push_back(normalize(min_l + k, -half[i], half[i]));
            else
                for (int k = -1; k >= shift[i]; k--) clear_id.push_back(normalize(min_l + k, -half[i], half[i]));
            const int a1 = (i + 1) % 3, a2 = (i + 2) % 3;
            for (int idd : clear_id)
                for (int u = -half[a1]; u <= half[a1]; u++)
                    for (int w = -half[a2]; w <= half[a2]; w++) {
                        int l[3];
                        l[i] = idd;
                        l[a1] = u;
                        l[a2] = w;
                        const int64_t h = local_hash(l);
                        occupied_cnt[h] = 0;
                        unknown_cnt[h] = (int16_t)sub_grid_num;
                    }
        }
        set_origin(no);
    }

    // counter_map.cpp:94-151 (the jumping-edge hook is empty for ESDFMap, esdf_map.h:137-138)
    void update_counter(const double pos[3], int from_type, int to_type) {
        const int64_t a = hash_from_pos(pos);
        if (from_type == ROG_OCCUPIED) occupied_cnt[a] -= 1;
        if (to_type == ROG_OCCUPIED) occupied_cnt[a] += 1;
        if (from_type == ROG_UNKNOWN) unknown_cnt[a] -= 1;
        if (to_type == ROG_UNKNOWN) unknown_cnt[a] += 1;
    }

    bool occ_at(int64_t h) const { return occupied_cnt[h] > 0; }   // counter_map.h:117-119

    // esdf_map.cpp:842-900: lower envelope over box coordinates start..end; the callables get
    // ring-memory coordinates (box coordinate + id_l, wrapped past mem_end).
    template <typename FGet, typename FSet>
    void fill_line(FGet f_get, FSet f_set, int start, int end, int dim, int id_l) const {
        if (end < 0) return;
        const int n = size[dim];
        std::vector<int> v(n + 1);
        std::vector<double> z(n + 2);
        const int mem_end = (n - 1) - id_l;
        auto mem = [&](int q) { return q > mem_end ? q + id_l - n : q + id_l; };
        int k = start;
        v[start] = start;
        z[start] = -std::numeric_limits<double>::max();
        z[start + 1] = std::numeric_limits<double>::max();
        for (int q = start + 1; q <= end; q++) {
            k++;
            double s;
            do {
                k--;
                s = ((f_get(mem(q)) + q * q) - (f_get(mem(v[k])) + v[k] * v[k])) / (2 * q - 2 * v[k]);
            } while (s <= z[k]);
            k++;
            v[k] = q;
            z[k] = s;
            z[k + 1] = std::numeric_limits<double>::max();
        }
        k = start;
        for (int q = start; q <= end; q++) {
            while (z[k + 1] < q) k++;
            const double val = (q - v[k]) * (q - v[k]) + f_get(mem(v[k]));
            f_set(mem(q), val);
        }
    }

    // esdf_map.cpp:154-500
    void update_esdf(const double cur_odom[3]) {
        const double DMAX = std::numeric_limits<double>::max();
        int cur_i[3], idl[3], mem_end[3], lo[3], hi[3];
        pos_to_global(cur_odom, cur_i);
        global_to_local(bmin_i, idl);
        for (int i = 0; i < 3; i++) {
            idl[i] += half[i];
            mem_end[i] = size[i] - 1 - idl[i];
        }
        const int Sy = size[1], Sz = size[2];
        auto h3 = [&](int x, int y, int z) { return ((int64_t)x * Sy + y) * Sz + z; };
        auto h2 = [&](int x, int y) { return (int64_t)x * Sy + y; };
        auto wr = [&](int q, int a) { return q > mem_end[a] ? q + idl[a] - size[a] : q + idl[a]; };
        for (int i = 0; i < 3; i++) {
            upd_min_i[i] = std::max(cur_i[i] - half_box_i[i], bmin_i[i]);
            upd_max_i[i] = std::min(cur_i[i] + half_box_i[i], bmax_i[i]) - 1;
            lo[i] = upd_min_i[i] - bmin_i[i];
            hi[i] = upd_max_i[i] - bmin_i[i];
        }
        for (int i = 0; i < 3; i++)
            if (hi[i] < lo[i]) return;
        // ---- 3-D: positive transform z, y, x (:187-240)
        for (int sign = 0; sign < 2; sign++) {
            for (int x = lo[0]; x <= hi[0]; x++)
                for (int y = lo[1]; y <= hi[1]; y++) {
                    const int mx = wr(x, 0), my = wr(y, 1);
                    fill_line(
                        [&](int z) {
                            const bool o = occ_at(h3(mx, my, z));
                            return (sign == 0) == o ? 0.0 : DMAX;
                        },
                        [&](int z, double val) { tmp1[h3(mx, my, z)] = val; }, lo[2], hi[2], 2, idl[2]);
                }
            for (int x = lo[0]; x <= hi[0]; x++)
                for (int z = lo[2]; z <= hi[2]; z++) {
                    const int mx = wr(x, 0), mz = wr(z, 2);
                    fill_line([&](int y) { return tmp1[h3(mx, y, mz)]; },
                              [&](int y, double val) { tmp2[h3(mx, y, mz)] = val; }, lo[1], hi[1], 1, idl[1]);
                }
            std::vector<double>& dst = sign == 0 ? dist3 : tmp1;
            for (int y = lo[1]; y <= hi[1]; y++)
                for (int z = lo[2]; z <= hi[2]; z++) {
                    const int my = wr(y, 1), mz = wr(z, 2);
                    fill_line([&](int x) { return tmp2[h3(x, my, mz)]; },
                              [&](int x, double val) { dst[h3(x, my, mz)] = res * std::sqrt(val); }, lo[0], hi[0], 0,
                              idl[0]);
                }
        }
        // combine (:305-315): box coordinates used as memory coordinates, no wrap
        for (int x = lo[0]; x <= hi[0]; x++)
            for (int y = lo[1]; y <= hi[1]; y++)
                for (int z = lo[2]; z <= hi[2]; z++) {
                    const int64_t idx = h3(x, y, z);
                    if (tmp1[idx] > 0.0) dist3[idx] += (-tmp1[idx] + res);
                }
        // ---- 2-D maps (:320-400 critical: whole z range; :402-500 flat: up to the cell of z = 0.155)
        for (int which = 0; which < 2; which++) {
            std::vector<double> neg((size_t)size[0] * size[1], 0.0), tmp((size_t)size[0] * size[1], 0.0);
            std::vector<char> occ((size_t)size[0] * size[1], 0);
            std::vector<double>& out = which == 0 ? dist_crit : dist_flat;
            int z_hi = hi[2];
            if (which == 1) {
                const double p[3] = {0.0, 0.0, 0.155};
                int g[3], l[3];
                pos_to_global(p, g);
                global_to_local(g, l);
                const int lz = l[2] + half[2];
                z_hi = lz < hi[2] ? lz : hi[2];
            }
            for (int x = lo[0]; x <= hi[0]; x++)
                for (int y = lo[1]; y <= hi[1]; y++)
                    for (int z = lo[2]; z <= z_hi; z++) {
                        const int mx = wr(x, 0), my = wr(y, 1), mz = wr(z, 2);
                        const bool o = occ_at(h3(mx, my, mz));
                        occ[h2(mx, my)] = o;
                        if (o) break;
                    }
            for (int sign = 0; sign < 2; sign++) {
                for (int x = lo[0]; x <= hi[0]; x++) {
                    const int mx = wr(x, 0);
                    fill_line([&](int y) { return (sign == 0) == (bool)occ[h2(mx, y)] ? 0.0 : DMAX; },
                              [&](int y, double val) { tmp[h2(mx, y)] = val; }, lo[1], hi[1], 1, idl[1]);
                }
                std::vector<double>& dst = sign == 0 ? out : neg;
                for (int y = lo[1]; y <= hi[1]; y++) {
                    const int my = wr(y, 1);
                    fill_line([&](int x) { return tmp[h2(x, my)]; },
                              [&](int x, double val) { dst[h2(x, my)] = res * std::sqrt(val); }, lo[0], hi[0], 0,
                              idl[0]);
                }
            }
            // combine: y walks the x range (reference quirk), unwrapped; clipped to the row (deviation 2)
            for (int x = lo[0]; x <= hi[0]; x++)
                for (int y = lo[0]; y <= hi[0]; y++) {
                    if (y >= size[1]) continue;
                    const int64_t idx = h2(x, y);
                    out[idx] = neg[idx] > 0.0 ? out[idx] - neg[idx] + res : out[idx];
                }
        }
    }

    // ---- nearest-cell getters (esdf_map.cpp:78-120)
    double get_distance(const double p[3]) const { return dist3[hash_from_pos(p)]; }
    double get_critical_distance(const double p[3]) const { return dist_crit[hash2_from_pos(p)]; }
    double get_distance2d(const double p[3]) const { return dist_flat[hash2_from_pos(p)]; }

    // esdf_map.cpp:903-925
    void surround(const double pos[3], double pts[2][2][2][3], double diff[3]) const {
        double pm[3], ip[3];
        int idx[3];
        for (int i = 0; i < 3; i++) pm[i] = pos[i] - 0.5 * res * 1.0;
        pos_to_global(pm, idx);
        global_to_pos(idx, ip);
        for (int i = 0; i < 3; i++) diff[i] = (pos[i] - ip[i]) / res;
        for (int x = 0; x < 2; x++)
            for (int y = 0; y < 2; y++)
                for (int z = 0; z < 2; z++) {
                    const int c[3] = {idx[0] + x, idx[1] + y, idx[2] + z};
                    global_to_pos(c, pts[x][y][z]);
                }
    }
    void surround_dist(const double pos[3], double d[2][2][2], double diff[3]) const {
        double pts[2][2][2][3];
        surround(pos, pts, diff);
        for (int x = 0; x < 2; x++)
            for (int y = 0; y < 2; y++)
                for (int z = 0; z < 2; z++) d[x][y][z] = get_distance(pts[x][y][z]);
    }
    // esdf_map.cpp:1113-1125 (evaluateEDT :951-961)
    double evaluate_edt(const double pos[3]) const {
        double d[2][2][2], f[3];
        surround_dist(pos, d, f);
        const double v00 = (1 - f[0]) * d[0][0][0] + f[0] * d[1][0][0];
        const double v01 = (1 - f[0]) * d[0][0][1] + f[0] * d[1][0][1];
        const double v10 = (1 - f[0]) * d[0][1][0] + f[0] * d[1][1][0];
        const double v11 = (1 - f[0]) * d[0][1][1] + f[0] * d[1][1][1];
        const double v0 = (1 - f[1]) * v00 + f[1] * v10;
        const double v1 = (1 - f[1]) * v01 + f[1] * v11;
        return (1 - f[2]) * v0 + f[2] * v1;
    }
    // esdf_map.cpp:1127-1146 (evaluateFirstGrad :963-974) and getValueGrad :976-1003
    void value_grad(const double pos[3], double& dist, double grad[3]) const {
        double d[2][2][2], f[3];
        surround_dist(pos, d, f);
        const double v00 = (1 - f[0]) * d[0][0][0] + f[0] * d[1][0][0];
        const double v01 = (1 - f[0]) * d[0][0][1] + f[0] * d[1][0][1];
        const double v10 = (1 - f[0]) * d[0][1][0] + f[0] * d[1][1][0];
        const double v11 = (1 - f[0]) * d[0][1][1] + f[0] * d[1][1][1];
        const double v0 = (1 - f[1]) * v00 + f[1] * v10;
        const double v1 = (1 - f[1]) * v01 + f[1] * v11;
        dist = (1 - f[2]) * v0 + f[2] * v1;
        grad[2] = (v1 - v0) * res_inv;
        grad[1] = ((1 - f[2]) * (v10 - v00) + f[2] * (v11 - v01)) * res_inv;
        grad[0] = (1 - f[2]) * (1 - f[1]) * (d[1][0][0] - d[0][0][0]);
        grad[0] += (1 - f[2]) * f[1] * (d[1][1][0] - d[0][1][0]);
        grad[0] += f[2] * (1 - f[1]) * (d[1][0][1] - d[0][0][1]);
        grad[0] += f[2] * f[1] * (d[1][1][1] - d[0][1][1]);
        grad[0] *= res_inv;
    }
    // esdf_map.cpp:1005-1051 (critical) and :1053-1097 (flat): bilinear in x, y; z of the taps = z of pos
    void value_grad_2d(const double pos[3], bool critical, double& dist, double grad[3]) const {
        double pm[3] = {pos[0] - 0.5 * res * 1.0, pos[1] - 0.5 * res * 1.0, pos[2] - 0.5 * res * 0.0};
        int idx[3];
        double ip[3], f[3];
        pos_to_global(pm, idx);
        global_to_pos(idx, ip);
        for (int i = 0; i < 3; i++) f[i] = (pos[i] - ip[i]) / res;
        double d[2][2];
        for (int x = 0; x < 2; x++)
            for (int y = 0; y < 2; y++) {
                const int c[3] = {idx[0] + x, idx[1] + y, idx[2]};
                double cp[3];
                global_to_pos(c, cp);
                d[x][y] = critical ? get_critical_distance(cp) : get_distance2d(cp);
            }
        const double fxy1 = f[0] * d[1][0] + (1 - f[0]) * d[0][0];
        const double fxy2 = f[0] * d[1][1] + (1 - f[0]) * d[0][1];
        dist = (1 - f[1]) * fxy1 + f[1] * fxy2;
        const double g0 = (1 - f[1]) * (d[1][0] - d[0][0]) + f[1] * (d[1][1] - d[0][1]);
        const double g1 = -fxy1 + fxy2;
        grad[0] = g0 * res_inv;
        grad[1] = g1 * res_inv;
        grad[2] = 0.0 * res_inv;
    }

    // esdf_map.cpp:122-152 with the stepping rule of include/utils/raycaster.cpp:66-192. The cell
    // that contains `end` is never tested; start and end in the same cell test nothing.
    bool is_line_free_2d(const double s2[2], const double e2[2], double threshold) const {
        const double DMAX = std::numeric_limits<double>::max();
        const double s[3] = {s2[0], s2[1], 0.0}, e[3] = {e2[0], e2[1], 0.0};
        int si[3], ei[3], cur[3], dir[3];
        for (int i = 0; i < 3; i++) {
            si[i] = ifloor(s[i] / res);
            ei[i] = ifloor(e[i] / res);
            cur[i] = si[i];
            const int dlt = ei[i] - si[i];
            dir[i] = (0 < dlt) - (dlt < 0);
        }
        double t_step[3] = {DMAX, DMAX, DMAX}, t_bound[3] = {DMAX, DMAX, DMAX};
        if (!(dir[0] == 0 && dir[1] == 0 && dir[2] == 0)) {
            double dd[3];
            for (int i = 0; i < 3; i++) dd[i] = std::fabs(e[i] - s[i]);
            const double tmax = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
            for (int i = 0; i < 3; i++) dd[i] /= tmax;
            for (int i = 0; i < 3; i++) {
                t_step[i] = dir[i] == 0 ? DMAX : std::fabs(res / dd[i]);
                const double centre = ((double)si[i] + 0.5) * res;
                const double nb = centre + dir[i] * res * 0.5;
                t_bound[i] = dir[i] == 0 ? DMAX : std::fabs(nb - s[i]) / dd[i];
            }
        }
        // a ray that steps past its end cell through rounding never terminates in the reference
        // (deviation 3): the walk is capped at the Manhattan cell distance, after which it reports free
        const long cap = (long)std::abs(ei[0] - si[0]) + std::abs(ei[1] - si[1]) + std::abs(ei[2] - si[2]) + 1;
        for (long it = 0; it < cap; it++) {
            double pt[3];
            for (int i = 0; i < 3; i++) pt[i] = ((double)cur[i] + 0.5) * res;
            if (cur[0] == ei[0] && cur[1] == ei[1] && cur[2] == ei[2]) return true;
            if (t_bound[0] < t_bound[1]) {
                if (t_bound[0] < t_bound[2]) { cur[0] += dir[0]; t_bound[0] += t_step[0]; }
                else { cur[2] += dir[2]; t_bound[2] += t_step[2]; }
            } else {
                if (t_bound[1] < t_bound[2]) { cur[1] += dir[1]; t_bound[1] += t_step[1]; }
                else { cur[2] += dir[2]; t_bound[2] += t_step[2]; }
            }
            if (dist_flat[hash2_from_pos(pt)] < threshold) return false;
        }
        return true;
    }
};


// ---------------------------------------------------------------------------------------------------------------
// rog_map::ProbMap — the probabilistic occupancy layer that drives the ESDF counter map from point clouds
// (SURVEY §8 row N3). Restates, in the reference's operation order:
//   src/rog_map/src/rog_map/prob_map.cpp:25-88    initProbMap (geometry, ceil / ground snapped to the grid)
//   src/rog_map/src/rog_map/prob_map.cpp:291-300  slideAllMap
//   src/rog_map/src/rog_map/prob_map.cpp:302-373  updateProbMap (incl. the first-frame sphere clearing)
//   src/rog_map/src/rog_map/prob_map.cpp:512-541  resetCell (cells that leave the map when it slides)
//   src/rog_map/src/rog_map/prob_map.cpp:543-569  probabilisticMapFromCache
//   src/rog_map/src/rog_map/prob_map.cpp:571-664  hitPointUpdate / missPointUpdate (log-odds, type transitions)
//   src/rog_map/src/rog_map/prob_map.cpp:666-778  raycastProcess (point filters, clipping, 3-D ray walk)
//   src/rog_map/src/rog_map/prob_map.cpp:780-789  insertUpdateCandidate
//   src/rog_map/src/rog_map/prob_map.cpp:791-820  updateLocalBox
//   src/rog_map/include/utils/common_lib.hpp:148-170 lineBoxIntersectPoint
//   src/rog_map/include/utils/raycaster.cpp:66-192   RayCaster::setInput / step
// The inflation map (InfMap) and the frontier counter map receive the same notifications in the reference; they
// do not feed the ESDF and are not restated (frontier_extraction_en = false).
struct RogProb {
    // SlidingMap of the probability grid
    double res = 0, res_inv = 0;
    int half[3] = {0, 0, 0}, size[3] = {0, 0, 0};
    int64_t vox = 0;
    bool sliding_en = true;
    double sliding_thresh = 0;
    int origin_i[3] = {0, 0, 0};
    double origin_d[3] = {0, 0, 0}, bound_min_d[3] = {0, 0, 0}, bound_max_d[3] = {0, 0, 0};
    // config (config.hpp:160-262, 336-382)
    float l_hit = 0, l_miss = 0, l_min = 0, l_max = 0, l_occ = 0, l_free = 0;
    double range_min = 0, range_max = 0, sqr_range_max = 0;
    double ceil_h = 0, ground_h = 0;
    int half_update_box_i[3] = {0, 0, 0};
    int point_filt_num = 1, batch_update_size = 1, intensity_thresh = -1;
    bool raycasting_en = true;
    // state
    std::vector<float> occupancy;
    std::vector<uint16_t> op_cnt, hit_cnt;
    std::vector<std::array<int, 3>> cache;      // update_cache_id_g (a queue: pushed and drained in order)
    int batch_counter = 0;
    bool inited = false, map_empty = true, first_frame = true;
    double box_min[3] = {0, 0, 0}, box_max[3] = {0, 0, 0};   // local_update_box_min / max
    RogEsdf* esdf = nullptr;

    static int ifloor(double v) { return (int)std::floor(v); }
    void pos_to_global(const double p[3], int id[3]) const {
        for (int i = 0; i < 3; i++) id[i] = ifloor(p[i] * res_inv);
    }
    void global_to_pos(const int id[3], double p[3]) const {
        for (int i = 0; i < 3; i++) p[i] = ((double)id[i] + 0.5) * res;
    }
    void global_to_local(const int g[3], int l[3]) const {
        for (int i = 0; i < 3; i++) {
            int v = g[i] % size[i];
            if (v > half[i]) v -= size[i];
            else if (v < -half[i]) v += size[i];
            l[i] = v;
        }
    }
    int64_t local_hash(const int l[3]) const {
        return (int64_t)(l[0] + half[0]) * size[1] * size[2] + (int64_t)(l[1] + half[1]) * size[2] + (l[2] + half[2]);
    }
    int64_t hash_from_global(const int g[3]) const {
        int l[3];
        global_to_local(g, l);
        return local_hash(l);
    }
    int64_t hash_from_pos(const double p[3]) const {
        int g[3];
        pos_to_global(p, g);
        return hash_from_global(g);
    }
    bool inside_local_map_i(const int g[3]) const {
        for (int i = 0; i < 3; i++)
            if (std::abs(g[i] - origin_i[i]) - half[i] > 0) return false;
        return true;
    }
    bool inside_local_map(const double p[3]) const {
        int g[3];
        pos_to_global(p, g);
        return inside_local_map_i(g);
    }
    // sliding_map.cpp:262-268 + :234-259 (ORIGIN_AT_CORNER): position of a ring cell under the CURRENT origin
    void hash_to_pos(int64_t h, double p[3]) const {
        int l[3];
        l[0] = (int)(h / ((int64_t)size[1] * size[2]));
        l[1] = (int)((h - (int64_t)l[0] * size[1] * size[2]) / size[2]);
        l[2] = (int)(h - (int64_t)l[0] * size[1] * size[2] - (int64_t)l[1] * size[2]);
        for (int i = 0; i < 3; i++) {
            l[i] -= half[i];
            const int min_g = -half[i] + origin_i[i];
            int min_l = min_g % size[i];
            min_l -= min_l > half[i] ? size[i] : 0;
            min_l += min_l < -half[i] ? size[i] : 0;
            int d = l[i] - min_l;
            d = d < 0 ? size[i] + d : d;
            p[i] = ((double)(d + min_g) + 0.5) * res;
        }
    }
    // sliding_map.cpp:85-97
    void set_origin(const double od[3], const int oi[3]) {
        int bmin[3], bmax[3];
        for (int i = 0; i < 3; i++) {
            origin_i[i] = oi[i];
            origin_d[i] = od[i];
            bmax[i] = oi[i] + half[i];
            bmin[i] = oi[i] - half[i];
        }
        global_to_pos(bmin, bound_min_d);
        global_to_pos(bmax, bound_max_d);
    }
    bool is_occupied(float v) const { return (double)v >= (double)l_occ; }      // prob_map.h:125-135 (double compare)
    bool is_known_free(float v) const { return (double)v < (double)l_free; }
    int type_of(float v) const { return is_occupied(v) ? ROG_OCCUPIED : (is_known_free(v) ? ROG_KNOWN_FREE : ROG_UNKNOWN); }

    // prob_map.cpp:25-88 with config.hpp's derived quantities; `logit` in float as config.hpp:229-235
    void init(RogEsdf* esdf_, const int half_map_size_i[3], double resolution, bool map_sliding_en,
              double map_sliding_thresh, const double fix_origin[3], const float p[6] /*hit miss min max occ free*/,
              double ray_min, double ray_max, double virtual_ceil, double virtual_ground, double inflation_resolution,
              int inflation_step, const double local_update_box_d[3], int filt_num, int batch_size, int intensity,
              bool raycasting) {
        esdf = esdf_;
        res = resolution;
        res_inv = 1.0 / resolution;
        sliding_en = map_sliding_en;
        sliding_thresh = map_sliding_thresh;
        for (int i = 0; i < 3; i++) {
            half[i] = half_map_size_i[i];
            size[i] = 2 * half[i] + 1;
        }
        vox = (int64_t)size[0] * size[1] * size[2];
        auto logit = [](float x) -> float { return std::log(x / (1 - x)); };
        l_hit = logit(p[0]); l_miss = logit(p[1]); l_min = logit(p[2]);
        l_max = logit(p[3]); l_occ = logit(p[4]); l_free = logit(p[5]);
        range_min = ray_min;
        range_max = ray_max;
        sqr_range_max = ray_max * ray_max;
        point_filt_num = filt_num <= 0 ? 1 : filt_num;
        batch_update_size = batch_size <= 0 ? 1 : batch_size;
        intensity_thresh = intensity;
        raycasting_en = raycasting;
        for (int i = 0; i < 3; i++) half_update_box_i[i] = (int)((local_update_box_d[i] / 2) / resolution);   // config.hpp:377-379
        // InfMap's constructor takes the Config by reference and pulls ceil / ground in by inflation_step cells of the
        // inflation grid (inf_map.cpp:81-89; config.hpp:338-349 rounds that grid up to a multiple of the resolution) ...
        const int inf_ratio = (int)std::ceil(inflation_resolution / resolution);
        const double inf_res = resolution * inf_ratio;
        const int ceil_id = (int)(virtual_ceil / inf_res + 0.5) - inflation_step;
        const int ground_id = (int)(virtual_ground / inf_res + 0.5) + inflation_step;
        virtual_ceil = ceil_id * inf_res;
        virtual_ground = ground_id * inf_res;
        // ... then prob_map.cpp:66-71 snaps them to the probability grid
        ceil_h = (double)ifloor(virtual_ceil * res_inv) * resolution;
        ground_h = (double)ifloor(virtual_ground * res_inv) * resolution;
        occupancy.assign(vox, 0.f);
        op_cnt.assign(vox, 0);
        hit_cnt.assign(vox, 0);
        cache.clear();
        batch_counter = 0;
        inited = false;
        map_empty = true;
        first_frame = true;
        if (!map_sliding_en) {
            // sliding_map.cpp:58-61 + prob_map.cpp:73-77
            int oi[3];
            pos_to_global(fix_origin, oi);
            for (int i = 0; i < 3; i++) {
                origin_d[i] = fix_origin[i];
                origin_i[i] = oi[i];
            }
            slide_all(fix_origin);
        }
    }

    // prob_map.cpp:512-541 (inf / frontier notifications dropped)
    void reset_cell(int64_t h) {
        float& ret = occupancy[h];
        if (is_occupied(ret)) {
            double p[3];
            hash_to_pos(h, p);
            if (esdf) esdf->update_counter(p, ROG_OCCUPIED, ROG_UNKNOWN);
        } else if (is_known_free(ret)) {
            double p[3];
            hash_to_pos(h, p);
            if (esdf) esdf->update_counter(p, ROG_KNOWN_FREE, ROG_UNKNOWN);
        }
        ret = 0;
    }
    // prob_map.cpp:822-833
    void reset_local_map() {
        std::fill(occupancy.begin(), occupancy.end(), 0.f);
        cache.clear();
        batch_counter = 0;
        std::fill(op_cnt.begin(), op_cnt.end(), (uint16_t)0);
        std::fill(hit_cnt.begin(), hit_cnt.end(), (uint16_t)0);
    }
    // sliding_map.cpp:113-166 with ProbMap's resetCell / resetLocalMap
    void map_sliding(const double odom[3]) {
        int no[3];
        pos_to_global(odom, no);
        double nd[3];
        for (int i = 0; i < 3; i++) nd[i] = (double)no[i] * res;
        int shift[3];
        for (int i = 0; i < 3; i++) shift[i] = no[i] - origin_i[i];
        for (int i = 0; i < 3; i++)
            if (std::fabs((double)shift[i]) > size[i]) {
                reset_local_map();
                set_origin(nd, no);
                return;
            }
        auto normalize = [](int x, int a, int b) {
            const int range = b - a + 1;
            const int y = (x - a) % range;
            return (y < 0 ? y + range : y) + a;
        };
        for (int i = 0; i < 3; i++) {
            if (shift[i] == 0) continue;
            const int min_g = -half[i] + origin_i[i];
            const int min_l = min_g % size[i];
            std::vector<int> clear_id;
            if (shift[i] > 0)
                for (int k = 0; k < shift[i]; k++) clear_id.push_back(normalize(min_l + k, -half[i], half[i]));
            else
                for (int k = -1; k >= shift[i]; k--) clear_id.push_back(normalize(min_l + k, -half[i], half[i]));
            const int a1 = (i + 1) % 3, a2 = (i + 2) % 3;
            for (int idd : clear_id)
                for (int u = -half[a1]; u <= half[a1]; u++)
                    for (int w = -half[a2]; w <= half[a2]; w++) {
                        int l[3];
                        l[i] = idd;
                        l[a1] = u;
                        l[a2] = w;
                        reset_cell(local_hash(l));
                    }
        }
        set_origin(nd, no);
    }
    // prob_map.cpp:291-300
    void slide_all(const double pos[3]) {
        map_sliding(pos);
        if (esdf) esdf->slide(pos);
    }

    // prob_map.cpp:791-820
    void update_local_box(const double odom[3]) {
        int oi[3], lo[3], hi[3];
        pos_to_global(odom, oi);
        for (int i = 0; i < 3; i++) {
            hi[i] = raycasting_en ? oi[i] + half_update_box_i[i] : 0;     // (uninitialised in the reference without raycasting)
            lo[i] = raycasting_en ? oi[i] - half_update_box_i[i] : 0;
        }
        global_to_pos(lo, box_min);
        global_to_pos(hi, box_max);
        for (int i = 0; i < 3; i++) {
            box_max[i] = std::min(box_max[i], bound_max_d[i]);
            box_min[i] = std::max(box_min[i], bound_min_d[i]);
        }
    }
    // prob_map.cpp:780-789
    void insert_candidate(const int g[3], bool is_hit) {
        const int64_t h = hash_from_global(g);
        op_cnt[h]++;
        if (op_cnt[h] == 1) cache.push_back({g[0], g[1], g[2]});
        if (is_hit) hit_cnt[h]++;
    }
    // prob_map.cpp:571-664
    void hit_miss_update(const double pos[3], int64_t h, int num, bool hit) {
        float& ret = occupancy[h];
        const int from = type_of(ret);
        if (hit) {
            ret += l_hit * num;
            if (ret > l_max) ret = l_max;
        } else {
            ret += l_miss * num;
            if (ret < l_min) ret = l_min;
        }
        const int to = type_of(ret);
        if (from != to) {
            int g[3];
            double c[3];
            pos_to_global(pos, g);
            global_to_pos(g, c);
            if (esdf) esdf->update_counter(c, from, to);
        }
    }
    // prob_map.cpp:543-569
    void from_cache() {
        for (size_t k = 0; k < cache.size(); k++) {
            const int g[3] = {cache[k][0], cache[k][1], cache[k][2]};
            const int64_t h = hash_from_global(g);
            double pos[3];
            global_to_pos(g, pos);
            if (hit_cnt[h] > 0) hit_miss_update(pos, h, hit_cnt[h], true);
            else hit_miss_update(pos, h, (int)op_cnt[h] - (int)hit_cnt[h], false);
            hit_cnt[h] = 0;
            op_cnt[h] = 0;
        }
        cache.clear();
    }
    // common_lib.hpp:148-170
    static void line_box_intersect(const double pt[3], const double pos[3], const double bmin[3], const double bmax[3],
                                   double out[3]) {
        double diff[3], max_tc[3], min_tc[3];
        for (int i = 0; i < 3; i++) {
            diff[i] = pt[i] - pos[i];
            max_tc[i] = bmax[i] - pos[i];
            min_tc[i] = bmin[i] - pos[i];
        }
        double min_t = 1000000;
        for (int i = 0; i < 3; i++)
            if (std::fabs(diff[i]) > 0) {
                const double t1 = max_tc[i] / diff[i];
                if (t1 > 0 && t1 < min_t) min_t = t1;
                const double t2 = min_tc[i] / diff[i];
                if (t2 > 0 && t2 < min_t) min_t = t2;
            }
        for (int i = 0; i < 3; i++) out[i] = pos[i] + (min_t - 1e-3) * diff[i];
    }
    // (p - o).normalized() * k + o   with Eigen's evaluation order: squaredNorm = (x^2 + y^2) + z^2, v / sqrt(.)
    static void along(const double p[3], const double o[3], double k, double out[3]) {
        const double d[3] = {p[0] - o[0], p[1] - o[1], p[2] - o[2]};
        const double n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const double n = std::sqrt(n2);
        for (int i = 0; i < 3; i++) out[i] = (n2 > 0.0 ? d[i] / n : d[i]) * k + o[i];
    }
    // prob_map.cpp:666-778. cloud: n x (x, y, z, intensity) float32
    void raycast_process(const float* cloud, int64_t n, const double odom[3]) {
        std::vector<std::array<double, 3>> rays;
        int temporal = 0;
        for (int64_t i = 0; i < n; i++) {
            const float* c = cloud + 4 * i;
            if (intensity_thresh > 0 && c[3] < intensity_thresh) continue;
            if (temporal++ % point_filt_num) continue;
            double p[3] = {(double)c[0], (double)c[1], (double)c[2]};
            int g[3];
            if (!raycasting_en) {
                if (inside_local_map(p)) {
                    pos_to_global(p, g);
                    insert_candidate(g, true);
                }
                continue;
            }
            bool update_hit = true;
            if (p[2] > ceil_h) {
                update_hit = false;
                const double dz = p[2] - odom[2], pc = ceil_h - odom[2];
                double q[3];
                along(p, odom, 1.0, q);     // (p - o).normalized() + o, then scaled below
                const double d[3] = {p[0] - odom[0], p[1] - odom[1], p[2] - odom[2]};
                const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                for (int k = 0; k < 3; k++) p[k] = odom[k] + (d[k] / nn) * pc / dz;
            } else if (p[2] < ground_h) {
                update_hit = false;
                const double dz = p[2] - odom[2], pc = ground_h - odom[2];
                const double d[3] = {p[0] - odom[0], p[1] - odom[1], p[2] - odom[2]};
                const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                for (int k = 0; k < 3; k++) p[k] = odom[k] + (d[k] / nn) * pc / dz;
            }
            {
                const double d[3] = {p[0] - odom[0], p[1] - odom[1], p[2] - odom[2]};
                const double sqr = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (sqr > sqr_range_max) {
                    const double k = range_max / std::sqrt(sqr);
                    for (int a = 0; a < 3; a++) p[a] = k * d[a] + odom[a];
                    update_hit = false;
                }
            }
            {
                double lo = p[0] - box_min[0], hi = p[0] - box_max[0];
                for (int a = 1; a < 3; a++) {
                    lo = std::min(lo, p[a] - box_min[a]);
                    hi = std::max(hi, p[a] - box_max[a]);
                }
                if (lo < 0 || hi > 0) {
                    double q[3];
                    line_box_intersect(p, odom, box_min, box_max, q);
                    for (int a = 0; a < 3; a++) p[a] = q[a];
                    update_hit = false;
                }
            }
            rays.push_back({p[0], p[1], p[2]});
            if (update_hit) {
                pos_to_global(p, g);
                insert_candidate(g, true);
            }
        }
        if (!raycasting_en) return;
        const double DMAX = std::numeric_limits<double>::max();
        // the reference keeps ONE RayCaster: a ray whose end points share a cell leaves the stepping tables of the
        // previous ray in place (setInput returns early, raycaster.cpp:101-103) — harmless, its first step ends it
        for (const auto& pr : rays) {
            const double p[3] = {pr[0], pr[1], pr[2]};
            double s[3];
            {
                const double d[3] = {p[0] - odom[0], p[1] - odom[1], p[2] - odom[2]};
                const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                for (int a = 0; a < 3; a++) s[a] = (d[a] / nn) * range_min + odom[a];
            }
            int si[3], ei[3], cur[3], dir[3];
            for (int a = 0; a < 3; a++) {
                si[a] = ifloor(s[a] / res);
                ei[a] = ifloor(p[a] / res);
                cur[a] = si[a];
                const int dlt = ei[a] - si[a];
                dir[a] = (0 < dlt) - (dlt < 0);
            }
            if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) continue;      // first step() returns false
            double t_step[3], t_bound[3], dd[3];
            for (int a = 0; a < 3; a++) dd[a] = std::fabs(p[a] - s[a]);
            const double tmax = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
            for (int a = 0; a < 3; a++) dd[a] /= tmax;
            for (int a = 0; a < 3; a++) {
                t_step[a] = dir[a] == 0 ? DMAX : std::fabs(res / dd[a]);
                const double centre = ((double)si[a] + 0.5) * res;
                const double nb = centre + dir[a] * res * 0.5;
                t_bound[a] = dir[a] == 0 ? DMAX : std::fabs(nb - s[a]) / dd[a];
            }
            while (true) {
                double pt[3];
                for (int a = 0; a < 3; a++) pt[a] = ((double)cur[a] + 0.5) * res;
                if (cur[0] == ei[0] && cur[1] == ei[1] && cur[2] == ei[2]) break;
                if (t_bound[0] < t_bound[1]) {
                    if (t_bound[0] < t_bound[2]) { cur[0] += dir[0]; t_bound[0] += t_step[0]; }
                    else { cur[2] += dir[2]; t_bound[2] += t_step[2]; }
                } else {
                    if (t_bound[1] < t_bound[2]) { cur[1] += dir[1]; t_bound[1] += t_step[1]; }
                    else { cur[2] += dir[2]; t_bound[2] += t_step[2]; }
                }
                int g[3];
                pos_to_global(pt, g);
                if (!inside_local_map_i(g)) break;
                insert_candidate(g, false);
            }
        }
    }
    // prob_map.cpp:302-373
    void update(const float* cloud, int64_t n, const double pos[3]) {
        if (sliding_en && !inside_local_map(pos) && batch_counter == 0) {
            slide_all(pos);
            return;
        }
        if (pos[2] > ceil_h) return;
        else if (pos[2] < ground_h) return;
        {
            const double d[3] = {pos[0] - origin_d[0], pos[1] - origin_d[1], pos[2] - origin_d[2]};
            const double nrm = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (batch_counter == 0 && (map_empty || (sliding_en && nrm > sliding_thresh))) slide_all(pos);
        }
        if (!inited) {
            inited = true;
            slide_all(pos);
        }
        update_local_box(pos);
        raycast_process(cloud, n, pos);
        batch_counter++;
        if (batch_counter >= batch_update_size) {
            batch_counter = 0;
            from_cache();
            map_empty = false;
        }
        if (esdf) esdf->update_esdf(pos);
        if (first_frame) {
            // prob_map.cpp:357-372 (a function-level static in the reference: once per process)
            first_frame = false;
            for (double dx = -range_min; dx <= range_min; dx += res)
                for (double dy = -range_min; dy <= range_min; dy += res)
                    for (double dz = -range_min; dz <= range_min; dz += res) {
                        const double nrm = std::sqrt(dx * dx + dy * dy + dz * dz);
                        if (nrm <= range_min) {
                            const double pp[3] = {pos[0] + dx, pos[1] + dy, pos[2] + dz};
                            hit_miss_update(pp, hash_from_pos(pp), 999, false);
                        }
                    }
        }
    }
};

}  // namespace oracle
