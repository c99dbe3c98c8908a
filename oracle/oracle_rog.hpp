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

}  // namespace oracle
