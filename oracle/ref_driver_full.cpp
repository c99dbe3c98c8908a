// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C entry points over the REFERENCE'S OWN OPTIMIZER AND MAP, compiled unmodified from where they lie
// under /root/reference (oracle/Makefile target `_ref`; sources are named by path, nothing is copied):
//   src/planner/src/moma_traj_opt.cpp + include/planner/moma_traj_opt.h   MomaTrajOpt (optimizeTraj, both cost
//                                                                         callbacks, both penalty loops, MomaTraj,
//                                                                         checkFeasible, printConstraintsSituations)
//   src/map/src/grid_map.cpp + include/map/grid_map.h                     GridMap (cloudCallback rasterisation,
//                                                                         updateESDF / fillESDF, all lookups)
//   src/simulator/random_map_generator/src/random_map_generator.cpp      (GridMap owns one; never called here)
//   src/planner/src/graph_search.cpp, src/planner/src/topo_prm.cpp,       front-end pieces next to the solve (row N2):
//   src/planner/src/utils/raycast.cpp                                     GraphSearch::getDensePath,
//                                                                         TopologyPRM::lineVisib over RayCaster
// against the Eigen / ROS / PCL / boost stand-ins of oracle/ref_stubs. rog_map is not compiled: GridMap runs
// with `use_rog: false` (params/grid_map.yaml:3). `#define private public` below only opens the classes to
// this driver (problem set-up for single evaluations, buffer downloads); the reference's translation units
// are compiled as they are.
//
// Two knobs live in the stand-in ROS layer, because the reference reads them from ROS: the clock
// (ros::Time::now — frozen, so the 1.0 s ALM cap of moma_traj_opt.cpp:403 never fires, or real) and
// ros::ok() (a per-thread budget = the deterministic ALM round cap the oracle and the device use).
#include <array>
#include <atomic>
#include <cmath>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <fstream>
#include <queue>
#include <set>
#include <Eigen/Eigen>

#define private public
#define protected public
#include "planner/moma_traj_opt.h"
#include <list>
#include "planner/graph_search.h"
#include "planner/topo_prm.h"
#undef private
#undef protected

#include "../include/topay_b200.h"

using nmoma_planner::GridMap;
using nmoma_planner::MomaTraj;
using nmoma_planner::MomaTrajOpt;

namespace {

struct Quiet {   // the reference prints progress to std::cout; nestable and shared by the worker threads
    static std::mutex& mu() { static std::mutex m; return m; }
    static int& depth() { static int d = 0; return d; }
    static std::streambuf*& saved() { static std::streambuf* s = nullptr; return s; }
    Quiet() {
        std::lock_guard<std::mutex> l(mu());
        if (depth()++ == 0) saved() = std::cout.rdbuf(nullptr);
    }
    ~Quiet() {
        std::lock_guard<std::mutex> l(mu());
        if (--depth() == 0) {
            std::cout.rdbuf(saved());
            std::cout.clear();
        }
    }
};

struct RefGrid {
    GridMap::Ptr gm;
};

struct RefOpt {
    std::shared_ptr<MomaTrajOpt> opt;
};

void set_lbfgs(lbfgs::lbfgs_parameter_t& r, const topay_lbfgs_params& p) {
    r.mem_size = p.mem_size;
    r.g_epsilon = p.g_epsilon;
    r.past = p.past;
    r.delta = p.delta;
    r.max_iterations = p.max_iterations;
    r.max_linesearch = p.max_linesearch;
    r.min_step = p.min_step;
    r.max_step = p.max_step;
    r.f_dec_coeff = p.f_dec_coeff;
    r.s_curv_coeff = p.s_curv_coeff;
    r.cautious_factor = p.cautious_factor;
    r.machine_prec = p.machine_prec;
}

Eigen::VectorXd vec2(const double* v) {
    Eigen::VectorXd o(2);
    o[0] = v[0];
    o[1] = v[1];
    return o;
}

void set_params(MomaTrajOpt& o, const topay_opt_params& p) {
    auto& q = o.opt_param;
    q.int_K = p.int_K;
    q.min_piece_num = p.min_piece_num;
    q.relu_mu = p.relu_mu;
    q.sample_interval = p.sample_interval;
    q.mean_time_lowb = 0.5;   // shadowed by locals in the reference (moma_traj_opt.cpp:1752-1753)
    q.mean_time_uppb = 2.0;
    q.energy_weights.resize(9);
    for (int d = 0; d < 9; d++) q.energy_weights[d] = p.energy_weights[d];
    auto& a = q.first_stage;
    a.time_weight = p.s1_time_weight;
    a.moment_weight = p.s1_moment_weight;
    a.acc_weight = p.s1_acc_weight;
    a.domega_weight = p.s1_domega_weight;
    a.mean_time_weight = p.s1_mean_time_weight;
    a.path_pos_weight = p.s1_path_pos_weight;
    a.lbgfs_normal_past = p.s1_lbfgs_normal_past;
    a.lbgfs_shot_path_past = p.s1_lbfgs_shot_path_past;
    a.shot_path_horizon = p.s1_shot_path_horizon;
    set_lbfgs(a.lbfgs_param, p.s1_lbfgs);
    auto& b = q.second_stage;
    b.time_weight = p.s2_time_weight;
    b.moment_weight = p.s2_moment_weight;
    b.acc_weight = p.s2_acc_weight;
    b.domega_weight = p.s2_domega_weight;
    b.collision_weight = p.s2_collision_weight;
    b.mani_colli_weight = p.s2_mani_colli_weight;
    b.self_colli_weight = p.s2_self_colli_weight;
    b.mani_pos_weight = p.s2_mani_pos_weight;
    b.mani_vel_weight = p.s2_mani_vel_weight;
    b.mani_acc_weight = p.s2_mani_acc_weight;
    b.mean_time_weight = p.s2_mean_time_weight;
    set_lbfgs(b.lbfgs_param, p.s2_lbfgs);
    b.alm_param.init_lambda = vec2(p.alm_init_lambda);
    b.alm_param.init_rho = vec2(p.alm_init_rho);
    b.alm_param.rho_max = vec2(p.alm_rho_max);
    b.alm_param.gamma = vec2(p.alm_gamma);
    b.alm_param.tolerance = Eigen::VectorXd::Constant(2, p.alm_tolerance);
}

const char* const kTermNames[TOPAY_NTERMS] = {"jerk", "time", "chassis_colli", "moment", "acc", "domega", "mani_colli",
                                              "self_colli", "mani_pos", "mani_vel", "mani_acc", "mean_time", "endp"};

}  // namespace

extern "C" {

// ------------------------------------------------------------------ GridMap
void* ref_grid_create(const topay_grid_desc* d) {
    Quiet q;
    auto& P = ros::stub_params();
    P["grid_map/map_size_x"] = {d->map_size[0]};
    P["grid_map/map_size_y"] = {d->map_size[1]};
    P["grid_map/map_size_z"] = {d->map_size[2]};
    P["grid_map/resolution"] = {d->resolution};
    P["agent/fixed_sequence"] = {1.0};
    P["grid_map/use_rog"] = {0.0};
    ros::stub_string_params()["agent/mode"] = "planner";   // no scene generation inside init (grid_map.cpp:67-75)
    ros::NodeHandle nh;
    RefGrid* g = new RefGrid();
    g->gm = std::make_shared<GridMap>();
    g->gm->init(nh);
    return g;
}
void ref_grid_dims(void* h, int32_t* dims) {
    GridMap& m = *((RefGrid*)h)->gm;
    for (int i = 0; i < 3; i++) dims[i] = m.voxel_num(i);
}
// GridMap::cloudCallback (grid_map.cpp:543-578): rasterise the cloud, updateESDF. `clear` first applies the reset
// of regenerateMap (grid_map.cpp:719-722: occ_2d and occ_3d; critical only when clear == 2).
void ref_grid_set_cloud(void* h, const float* xyz, int64_t n, int clear) {
    Quiet q;
    GridMap& m = *((RefGrid*)h)->gm;
    if (clear) {
        std::fill(m.occ_buffer_2d.begin(), m.occ_buffer_2d.end(), 0);
        std::fill(m.occ_buffer_3d.begin(), m.occ_buffer_3d.end(), 0);
        if (clear == 2) std::fill(m.occ_buffer_2d_critical.begin(), m.occ_buffer_2d_critical.end(), 0);
    }
    sensor_msgs::PointCloud2 msg;
    msg.xyzi.resize((size_t)n * 4);
    for (int64_t i = 0; i < n; i++) {
        msg.xyzi[4 * i] = xyz[3 * i];
        msg.xyzi[4 * i + 1] = xyz[3 * i + 1];
        msg.xyzi[4 * i + 2] = xyz[3 * i + 2];
        msg.xyzi[4 * i + 3] = 0.f;
    }
    m.map_ready = false;
    m.cloudCallback(msg);
}
// GridMap::loadMap (grid_map.cpp:800-809)
void ref_grid_load_map(void* h, const int8_t* occ2d, const int8_t* occ3d) {
    Quiet q;
    GridMap& m = *((RefGrid*)h)->gm;
    std::vector<char> a(occ2d, occ2d + m.buffer_size_2d), b(occ3d, occ3d + m.buffer_size_3d);
    m.loadMap(a, b);
}
void ref_grid_download(void* h, int which, double* out) {
    GridMap& m = *((RefGrid*)h)->gm;
    const std::vector<double>* v = &m.esdf_buffer_3d;
    if (which == TOPAY_MAP2D_FLAT) v = &m.esdf_buffer_2d;
    if (which == TOPAY_MAP2D_INFLATE) v = &m.esdf_buffer_2d_inflate;
    if (which == TOPAY_MAP2D_CRITICAL) v = &m.esdf_buffer_2d_critical;
    std::memcpy(out, v->data(), v->size() * sizeof(double));
}
void ref_grid_download_occupancy(void* h, int which, int8_t* out) {
    GridMap& m = *((RefGrid*)h)->gm;
    const std::vector<char>* v = &m.occ_buffer_3d;
    if (which == TOPAY_MAP2D_FLAT) v = &m.occ_buffer_2d;
    if (which == TOPAY_MAP2D_CRITICAL) v = &m.occ_buffer_2d_critical;
    std::memcpy(out, v->data(), v->size());
}
void ref_grid_query3d(void* h, const double* pos, int64_t n, double* dist, double* grad) {
    GridMap& m = *((RefGrid*)h)->gm;
    for (int64_t i = 0; i < n; i++) {
        Eigen::Vector3d p(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), g;
        m.getDisWithGradI3d(p, dist[i], g);
        if (grad)
            for (int k = 0; k < 3; k++) grad[3 * i + k] = g[k];
    }
}
void ref_grid_query2d(void* h, const double* pos, int64_t n, int which, double* dist, double* grad) {
    GridMap& m = *((RefGrid*)h)->gm;
    for (int64_t i = 0; i < n; i++) {
        Eigen::Vector2d p(pos[2 * i], pos[2 * i + 1]), g;
        m.getDisWithGradI2d(p, dist[i], g, which == TOPAY_MAP2D_INFLATE, which == TOPAY_MAP2D_CRITICAL);
        if (grad) {
            grad[2 * i] = g[0];
            grad[2 * i + 1] = g[1];
        }
    }
}
void ref_grid_distance3d(void* h, const double* pos, int64_t n, double* dist) {
    GridMap& m = *((RefGrid*)h)->gm;
    for (int64_t i = 0; i < n; i++) m.getDistance3d(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), dist[i]);
}
void ref_grid_distance2d(void* h, const double* pos, int64_t n, double* dist) {
    GridMap& m = *((RefGrid*)h)->gm;
    for (int64_t i = 0; i < n; i++) m.getDistance2d(Eigen::Vector2d(pos[2 * i], pos[2 * i + 1]), dist[i]);
}
void ref_grid_whole_body_collision(void* h, const double* states, int64_t n, int8_t* out) {
    GridMap& m = *((RefGrid*)h)->gm;
    for (int64_t i = 0; i < n; i++) {
        Eigen::VectorXd s(10);
        for (int k = 0; k < 10; k++) s[k] = states[10 * i + k];
        out[i] = m.isWholeBodyCollision(s) ? 1 : 0;
    }
}

// ------------------------------------------------------------------ MomaTrajOpt
void* ref_opt_create(void* grid, const topay_opt_params* p) {
    Quiet q;
    ros::NodeHandle nh;
    {   // MomaTrajOpt::init indexes the ALM vectors it reads (moma_traj_opt.h:921): they must be in the table
        static std::mutex mu;
        std::lock_guard<std::mutex> l(mu);
        auto& P = ros::stub_params();
        P["moma_traj_opt/energy_weights"] = std::vector<double>(p->energy_weights, p->energy_weights + 9);
        P["moma_traj_opt/second_stage/alm_param/init_lambda"] = {p->alm_init_lambda[0], p->alm_init_lambda[1]};
        P["moma_traj_opt/second_stage/alm_param/init_rho"] = {p->alm_init_rho[0], p->alm_init_rho[1]};
        P["moma_traj_opt/second_stage/alm_param/rho_max"] = {p->alm_rho_max[0], p->alm_rho_max[1]};
        P["moma_traj_opt/second_stage/alm_param/gamma"] = {p->alm_gamma[0], p->alm_gamma[1]};
        P["moma_traj_opt/second_stage/alm_param/tolerance"] = {p->alm_tolerance, p->alm_tolerance};
    }
    RefOpt* o = new RefOpt();
    o->opt = std::make_shared<MomaTrajOpt>(((RefGrid*)grid)->gm);
    o->opt->init(nh);        // publishers + the 13 debug_manager terms; the parameters follow
    set_params(*o->opt, *p);
    return o;
}
void ref_opt_destroy(void* h) { delete (RefOpt*)h; }

// One evaluation at x of one candidate whose problem data are given in the packed form the device ABI
// uses (head / tail 9 x 3 row-major, inner_xy N x 2). Sets the members optimizeTraj would have set
// (moma_traj_opt.cpp:146-147, 248-321) and calls the reference's own cost callback.
void ref_opt_eval(void* h, int stage, int N, const double* head, const double* tail, const double* sxy,
                  const double* exy, const double* inner_xy, const double* lambda, const double* rho, const double* x,
                  double* cost, double* grad, double* terms, double* coeff_out, double* final_xy) {
    Quiet q;
    MomaTrajOpt& o = *((RefOpt*)h)->opt;
    o.piece_num = N;
    o.start_state = Eigen::VectorXd::Zero(10);
    o.end_state = Eigen::VectorXd::Zero(10);
    o.start_state[0] = sxy[0];
    o.start_state[1] = sxy[1];
    o.start_state[2] = head[0];
    o.end_state[0] = exy[0];
    o.end_state[1] = exy[1];
    o.minco_start_state = Eigen::MatrixXd::Zero(9, 3);
    o.minco_end_state = Eigen::MatrixXd::Zero(9, 3);
    for (int d = 0; d < 9; d++)
        for (int k = 0; k < 3; k++) {
            o.minco_start_state(d, k) = head[d * 3 + k];
            o.minco_end_state(d, k) = tail[d * 3 + k];
        }
    o.init_inner_xy.clear();
    for (int i = 0; i < N; i++) o.init_inner_xy.push_back(Eigen::Vector2d(inner_xy[2 * i], inner_xy[2 * i + 1]));
    o.times = Eigen::VectorXd::Constant(N, 1.0);
    o.inner_pts = Eigen::MatrixXd::Zero(9, N - 1);
    o.minco_opt.reset(N, o.opt_param.energy_weights);
    o.alm_lambda = vec2(lambda);
    o.alm_rho = vec2(rho);
    const int n = topay_num_vars(N);
    Eigen::VectorXd xv(n), g(n);
    for (int i = 0; i < n; i++) xv[i] = x[i];
    g.setZero();
    *cost = stage == 1 ? MomaTrajOpt::firstStageCostCallback(&o, xv, g) : MomaTrajOpt::secondStageCostCallback(&o, xv, g);
    for (int i = 0; i < n; i++) grad[i] = g[i];
    if (terms)
        for (int t = 0; t < TOPAY_NTERMS; t++) terms[t] = o.debug_manager[kTermNames[t]];
    if (coeff_out) {
        const Eigen::MatrixXd& c = o.minco_opt.getCoeffs();
        for (int i = 0; i < 6 * N; i++)
            for (int d = 0; d < 9; d++) coeff_out[(size_t)i * 9 + d] = c(i, d);
    }
    if (final_xy) {
        final_xy[0] = o.final_xy_error[0];
        final_xy[1] = o.final_xy_error[1];
    }
}

struct RefSolveOut {
    int32_t status, piece_num;
    double cost, duration;
    double final_xy_err[2];
};

// MomaTrajOpt::optimizeTraj (moma_traj_opt.cpp:142-498) on one candidate. init_path: len x 10; bvel / bacc:
// 10 x 2 row-major. alm_max_rounds > 0 caps the ALM loop through ros::ok() (the deterministic cap of the
// oracle / device); wall_clock != 0 lets the reference's own 1.0 s cap run on the real clock.
int ref_opt_solve(void* h, const double* init_path, int len, const double* bvel, const double* bacc,
                  int alm_max_rounds, int wall_clock, RefSolveOut* out, double* T_out, double* coeff_out) {
    Quiet q;
    MomaTrajOpt& o = *((RefOpt*)h)->opt;
    std::vector<Eigen::VectorXd> path;
    for (int i = 0; i < len; i++) {
        Eigen::VectorXd s(10);
        for (int k = 0; k < 10; k++) s[k] = init_path[(size_t)i * 10 + k];
        path.push_back(s);
    }
    Eigen::MatrixXd bv(10, 2), ba(10, 2);
    for (int r = 0; r < 10; r++)
        for (int c = 0; c < 2; c++) {
            bv(r, c) = bvel[r * 2 + c];
            ba(r, c) = bacc[r * 2 + c];
        }
    ros::stub_clock_frozen() = wall_clock == 0;
    ros::stub_ok_budget() = alm_max_rounds > 0 ? alm_max_rounds : -1;
    const bool ok = o.optimizeTraj(path, bv, ba);
    ros::stub_ok_budget() = -1;
    const int N = o.piece_num;
    out->status = ok ? 1 : 0;
    out->piece_num = N;
    out->cost = o.traj_cost;
    out->duration = o.minco_opt.T1.sum();
    out->final_xy_err[0] = o.final_xy_error[0];
    out->final_xy_err[1] = o.final_xy_error[1];
    if (T_out)
        for (int i = 0; i < N; i++) T_out[i] = o.minco_opt.T1[i];
    if (coeff_out) {
        const Eigen::MatrixXd& c = o.minco_opt.getCoeffs();
        for (int i = 0; i < 6 * N; i++)
            for (int d = 0; d < 9; d++) coeff_out[(size_t)i * 9 + d] = c(i, d);
    }
    return ok ? 1 : 0;
}

// The success gate of the worker (planner.cpp:877-880) on the trajectory of the last solve.
void ref_opt_gate(void* h, int32_t* feasible_check, int32_t* feasible_print, double* total_duration) {
    Quiet q;
    MomaTrajOpt& o = *((RefOpt*)h)->opt;
    MomaTraj traj = o.getTraj();
    if (feasible_check) *feasible_check = o.checkFeasible(traj) ? 1 : 0;
    if (feasible_print) *feasible_print = o.printConstraintsSituations(traj) ? 1 : 0;
    if (total_duration) *total_duration = traj.getTotalDuration();
}

// Thread per candidate (planner.cpp:921-925): each worker owns a MomaTrajOpt, the GridMap is shared.
int ref_solve_batch(void* grid, const topay_opt_params* p, int n_cand, const int32_t* path_len, const double* init_paths,
                    const double* bvel, const double* bacc, int alm_max_rounds, int wall_clock, int n_threads,
                    RefSolveOut* out) {
    Quiet q;
    std::vector<size_t> off(n_cand + 1, 0);
    for (int i = 0; i < n_cand; i++) off[i + 1] = off[i] + (size_t)path_len[i] * 10;
    std::atomic<int> next(0);
    auto worker = [&]() {
        void* o = ref_opt_create(grid, p);
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n_cand) break;
            ref_opt_solve(o, init_paths + off[i], path_len[i], bvel + (size_t)i * 20, bacc + (size_t)i * 20,
                          alm_max_rounds, wall_clock, &out[i], nullptr, nullptr);
        }
        ref_opt_destroy(o);
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, n_threads); t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    return 0;
}


// ------------------------------------------------------------------ front-end pieces (row N2)
// GraphSearch::getDensePath (graph_search.cpp:119-176); out rows (x, y, theta, dt); returns the row count
int ref_dense_path(void* grid, const double* raw_xy, int n, double step_size, double start_yaw, double end_yaw,
                   double v_max, double w_max, double* out, int cap) {
    Quiet q;
    JPS::GraphSearch gs(((RefGrid*)grid)->gm, 0.0);
    std::vector<Eigen::Vector2d> raw;
    for (int i = 0; i < n; i++) raw.push_back(Eigen::Vector2d(raw_xy[2 * i], raw_xy[2 * i + 1]));
    const std::vector<Eigen::Vector4d> r = gs.getDensePath(raw, step_size, start_yaw, end_yaw, v_max, w_max);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++)
        for (int k = 0; k < 4; k++) out[4 * i + k] = r[i](k);
    return (int)r.size();
}
// TopologyPRM::lineVisib (topo_prm.cpp:278-315) on n segments; visible[i], pc[i] (untouched when visible)
void ref_line_visib(void* grid, const double* p1, const double* p2, int64_t n, double thresh, int use_critical,
                    int8_t* visible, double* pc) {
    Quiet q;
    auto& P = ros::stub_params();
    P["topo_prm/max_raw_path"] = {1.0};
    ros::NodeHandle nh;
    nmoma_planner::TopologyPRM prm;
    prm.grid_map_ptr = ((RefGrid*)grid)->gm;
    prm.init(nh);
    prm.use_critical = use_critical != 0;
    for (int64_t i = 0; i < n; i++) {
        Eigen::Vector3d a(p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]), b(p2[3 * i], p2[3 * i + 1], p2[3 * i + 2]), c;
        c << pc[3 * i], pc[3 * i + 1], pc[3 * i + 2];
        visible[i] = prm.lineVisib(a, b, thresh, c, 0) ? 1 : 0;
        for (int k = 0; k < 3; k++) pc[3 * i + k] = c(k);
    }
}

// TopologyPRM::sameTopoPath / discretizePath / pathLength (topo_prm.cpp:424-506)
static void make_prm(nmoma_planner::TopologyPRM& prm, void* grid, int use_critical) {
    auto& P = ros::stub_params();
    P["topo_prm/max_raw_path"] = {1.0};
    ros::NodeHandle nh;
    prm.grid_map_ptr = ((RefGrid*)grid)->gm;
    prm.init(nh);
    prm.use_critical = use_critical != 0;
}
static std::vector<Eigen::Vector3d> to_path(const double* p, int n) {
    std::vector<Eigen::Vector3d> v;
    for (int i = 0; i < n; i++) v.push_back(Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
    return v;
}
int ref_same_topo_path(void* grid, const double* p1, int n1, const double* p2, int n2, double thresh, int use_critical) {
    Quiet q;
    nmoma_planner::TopologyPRM prm;
    make_prm(prm, grid, use_critical);
    return prm.sameTopoPath(to_path(p1, n1), to_path(p2, n2), thresh) ? 1 : 0;
}
int ref_discretize_path(void* grid, const double* path, int n, int pt_num, double* out) {
    Quiet q;
    nmoma_planner::TopologyPRM prm;
    make_prm(prm, grid, 0);
    const std::vector<Eigen::Vector3d> r = prm.discretizePath(to_path(path, n), pt_num);
    for (size_t i = 0; i < r.size(); i++)
        for (int k = 0; k < 3; k++) out[3 * i + k] = r[i](k);
    return (int)r.size();
}
double ref_path_length(void* grid, const double* path, int n) {
    Quiet q;
    nmoma_planner::TopologyPRM prm;
    make_prm(prm, grid, 0);
    return prm.pathLength(to_path(path, n));
}

}  // extern "C"
