// Reference-facing C++ shims over the C ABI: the same class names, member names and argument
// meaning as the reference's GridMap (src/map/include/map/grid_map.h:77-219) and MomaTrajOpt
// (src/planner/include/planner/moma_traj_opt.h:613-675), so that Planner's worker
// (src/planner/src/planner.cpp:847-918) compiles against them unchanged apart from the include.
//
// The reference passes Eigen types; Eigen is not available in this build environment, so the shims
// are templates over "anything with size() and operator[] / operator()(r, c)". With Eigen present,
// Eigen::VectorXd / Eigen::MatrixXd / Eigen::Vector3d satisfy these directly.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/topay_b200.h"

namespace nmoma_planner {

inline void topay_check(int rc, const char* what) {
    if (rc != TOPAY_OK) throw std::runtime_error(std::string(what) + ": " + topay_last_error());
}

// MomaTraj (moma_traj_opt.h:26-247): durations + MINCO coefficients of one optimised trajectory.
struct MomaTraj {
    bool is_init = false;
    std::vector<double> durations;          // N
    std::vector<double> coeff;              // 6N x 9, row 6i+k = coefficient of t^k of piece i
    double start_se2[3] = {0, 0, 0};
    double getTotalDuration() const {
        double s = 0;
        for (double d : durations) s += d;
        return s;
    }
    int getPieceNum() const { return (int)durations.size(); }
};

class GridMap {
public:
    typedef std::shared_ptr<GridMap> Ptr;
    double resolution = 0, resolution_inv = 0;
    int voxel_num[3] = {0, 0, 0};
    bool use_rog = false;

    // GridMap::init (grid_map.cpp:6-87) with the rosparams passed explicitly.
    void init(double map_size_x, double map_size_y, double map_size_z, double res, int device = 0) {
        topay_grid_desc d;
        d.map_size[0] = map_size_x; d.map_size[1] = map_size_y; d.map_size[2] = map_size_z;
        d.resolution = res; d.chassis_colli_radius = 0.4; d.chassis_height = 0.155;
        topay_check(topay_field_create(&d, device, &f_), "topay_field_create");
        int32_t dims[3];
        topay_field_dims(f_, dims);
        for (int i = 0; i < 3; i++) voxel_num[i] = dims[i];
        resolution = res; resolution_inv = 1.0 / res;
    }
    ~GridMap() { topay_field_destroy(f_); }

    void loadMap(const std::vector<char>& occ_2d, const std::vector<char>& occ_3d) {      // grid_map.cpp:800
        topay_check(topay_field_set_occupancy(f_, (const int8_t*)occ_3d.data(), (const int8_t*)occ_2d.data(), nullptr),
                    "topay_field_set_occupancy");
        updateESDF();
    }
    // ingest half of regenerateMap / cloudCallback (grid_map.cpp:716-750, 543-578)
    void regenerateMap(const float* xyz, int64_t n_points) {
        topay_check(topay_field_clear(f_, 0), "topay_field_clear");
        topay_check(topay_field_rasterize_points(f_, xyz, n_points), "topay_field_rasterize_points");
        updateESDF();
    }
    void updateESDF() { topay_check(topay_field_rebuild(f_), "topay_field_rebuild"); map_ready_ = true; }
    bool mapReady() const { return map_ready_; }

    template <class V3> void getDisWithGradI3d(const V3& pos, double& distance, V3& grad) {   // grid_map.h:443
        double p[3] = {pos[0], pos[1], pos[2]}, g[3];
        topay_check(topay_field_query3d(f_, p, 1, &distance, g), "topay_field_query3d");
        grad[0] = g[0]; grad[1] = g[1]; grad[2] = g[2];
    }
    template <class V2> void getDisWithGradI2d(const V2& pos, double& distance, V2& grad, bool inflate = false,
                                               bool critical = false) {                      // grid_map.h:364
        double p[2] = {pos[0], pos[1]}, g[2];
        const int which = critical ? TOPAY_MAP2D_CRITICAL : (inflate ? TOPAY_MAP2D_INFLATE : TOPAY_MAP2D_FLAT);
        topay_check(topay_field_query2d(f_, p, 1, which, &distance, g), "topay_field_query2d");
        grad[0] = g[0]; grad[1] = g[1];
    }
    template <class V3> void getDistance3d(const V3& pos, double& distance) {                   // grid_map.h:307
        double p[3] = {pos[0], pos[1], pos[2]};
        topay_check(topay_field_distance3d(f_, p, 1, &distance), "topay_field_distance3d");
    }
    template <class V2> void getDistance2d(const V2& pos, double& distance) {                   // grid_map.h:256
        double p[2] = {pos[0], pos[1]};
        topay_check(topay_field_distance2d(f_, p, 1, &distance), "topay_field_distance2d");
    }
    template <class V> bool isWholeBodyCollision(const V& state) {                               // grid_map.h:613
        double s[10];
        for (int i = 0; i < 10; i++) s[i] = state[i];
        topay_robot_params rp;
        topay_robot_params_default(&rp);
        int8_t out = 1;
        topay_check(topay_field_whole_body_collision(f_, &rp, s, 1, &out), "topay_field_whole_body_collision");
        return out != 0;
    }
    // batched forms for the front-end (one launch for many samples)
    void getDisWithGradI3dBatch(const double* pos, int64_t n, double* dist, double* grad) {
        topay_check(topay_field_query3d(f_, pos, n, dist, grad), "topay_field_query3d");
    }
    std::vector<double> getESDFBuffer3d() {                                                      // grid_map.h:217
        std::vector<double> b((size_t)voxel_num[0] * voxel_num[1] * voxel_num[2]);
        topay_check(topay_field_download(f_, TOPAY_MAP3D, b.data()), "topay_field_download");
        return b;
    }
    std::vector<double> getESDFBuffer2d() {                                                      // grid_map.h:216
        std::vector<double> b((size_t)voxel_num[0] * voxel_num[1]);
        topay_check(topay_field_download(f_, TOPAY_MAP2D_FLAT, b.data()), "topay_field_download");
        return b;
    }
    topay_field* handle() const { return f_; }

private:
    topay_field* f_ = nullptr;
    bool map_ready_ = false;
};

class MomaTrajOpt {
public:
    typedef std::shared_ptr<MomaTrajOpt> Ptr;
    topay_opt_params opt_param;          // MomaTrajOptParam (moma_traj_opt.h:553-564)
    double traj_cost = 0.0;

    // max_cand > 1 turns the instance into the batched optimizer that replaces the <= 8
    // per-thread instances of planner.cpp:59-66.
    explicit MomaTrajOpt(GridMap::Ptr grid_map_, int max_cand = 1, int max_pieces = 32)
        : grid_map(grid_map_), max_cand_(max_cand), max_pieces_(max_pieces) {
        topay_opt_params_default(&opt_param);
        topay_robot_params_default(&moma_param);
    }
    ~MomaTrajOpt() { topay_solver_destroy(s_); }

    // void init(ros::NodeHandle&) reads the rosparams into opt_param (moma_traj_opt.h:845-941); here the
    // caller fills opt_param (defaults = params/optimizer.yaml) and then calls init().
    void init() { topay_check(topay_solver_create(&opt_param, &moma_param, grid_map->handle(), max_cand_, max_pieces_, &s_),
                              "topay_solver_create"); }

    // bool optimizeTraj(std::vector<Eigen::VectorXd> init_path, const Eigen::MatrixXd& boundary_vel,
    //                   const Eigen::MatrixXd& boundary_acc)          (moma_traj_opt.cpp:142)
    template <class Path, class Mat>
    bool optimizeTraj(const Path& init_path, const Mat& boundary_vel, const Mat& boundary_acc) {
        std::vector<std::vector<double>> flat(1);
        for (const auto& wp : init_path)
            for (int d = 0; d < 10; d++) flat[0].push_back(wp[d]);
        double bv[20], ba[20];
        for (int r = 0; r < 10; r++)
            for (int c = 0; c < 2; c++) { bv[r * 2 + c] = boundary_vel(r, c); ba[r * 2 + c] = boundary_acc(r, c); }
        std::vector<int> ok = optimizeTrajBatch(flat, bv, ba);
        return ok[0] != 0;
    }
    // All candidates of a plan at once: paths[c] is a flattened (len x 10) waypoint list; bvel / bacc are
    // n x 10 x 2 row-major. Returns the per-candidate success flags; best index by the reference's
    // shortest-duration rule (planner.cpp:999-1010) in best_by_duration.
    std::vector<int> optimizeTrajBatch(const std::vector<std::vector<double>>& paths, const double* bvel,
                                       const double* bacc) {
        const int n = (int)paths.size();
        std::vector<int32_t> len(n);
        std::vector<double> all;
        for (int c = 0; c < n; c++) {
            len[c] = (int32_t)(paths[c].size() / 10);
            all.insert(all.end(), paths[c].begin(), paths[c].end());
        }
        status_.assign(n, 0); pieces_.assign(n, 0); cost_.assign(n, 0.0);
        T_.assign((size_t)n * max_pieces_, 0.0);
        coeff_.assign((size_t)n * 6 * max_pieces_ * 9, 0.0);
        starts_.assign((size_t)n * 3, 0.0);
        for (int c = 0; c < n; c++)
            for (int d = 0; d < 3; d++) starts_[c * 3 + d] = paths[c][d];
        topay_result_batch out = {};
        out.status = status_.data(); out.piece_num = pieces_.data(); out.cost = cost_.data();
        out.T = T_.data(); out.coeff = coeff_.data();
        topay_check(topay_solver_solve_batch(s_, n, len.data(), all.data(), bvel, bacc, &out, &best_by_duration,
                                             &best_by_cost), "topay_solver_solve_batch");
        traj_cost = cost_[0];
        return std::vector<int>(status_.begin(), status_.end());
    }
    // MomaTraj getTraj() const (moma_traj_opt.h:943-946)
    MomaTraj getTraj(int idx = 0) const {
        MomaTraj t;
        const int N = pieces_[idx];
        t.durations.assign(T_.begin() + (size_t)idx * max_pieces_, T_.begin() + (size_t)idx * max_pieces_ + N);
        const size_t o = (size_t)idx * 6 * max_pieces_ * 9;
        t.coeff.assign(coeff_.begin() + o, coeff_.begin() + o + (size_t)6 * N * 9);
        for (int d = 0; d < 3; d++) t.start_se2[d] = starts_[idx * 3 + d];
        t.is_init = true;
        return t;
    }
    int32_t best_by_duration = -1, best_by_cost = -1;

private:
    topay_robot_params moma_param;
    GridMap::Ptr grid_map;
    topay_solver* s_ = nullptr;
    int max_cand_, max_pieces_;
    std::vector<int32_t> status_, pieces_;
    std::vector<double> cost_, T_, coeff_, starts_;
};

}  // namespace nmoma_planner
