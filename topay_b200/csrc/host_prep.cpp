// Host-side pre-processing of one candidate: waypoints -> NLP problem data + x0.
//
// Mirrors the first half of MomaTrajOpt::optimizeTraj
// (src/planner/src/moma_traj_opt.cpp:146-344) with the helpers of
// src/planner/include/planner/moma_traj_opt.h:676-807. SURVEY.md §8 row a1: negligible
// work, stays on the host. Also the default parameter fillers of the C ABI.
#include <cmath>
#include <cstring>
#include <vector>

#include "hd.cuh"

namespace {

struct Waypoint {          // one row of the reference's `sampled_path` (:152)
    double x, y, yaw, dyaw, darc;
    double q[TOPAY_DOF];
};

// getDurationTrapezoid, moma_traj_opt.h:676-697
double trapezoid_duration(double len, double v0, double v1, double vmax, double amax) {
    double v0s = v0 * v0, v1s = v1 * v1;
    const double vms = vmax * vmax;
    if (v0 > vmax) v0s = vms;
    if (v1 > vmax) v1s = vms;
    const double critical = (vms - v0s) / (2 * amax) + (vms - v1s) / (2 * amax);
    if (len >= critical) return (vmax - v0) / amax + (vmax - v1) / amax + (len - critical) / vmax;
    const double vp = std::sqrt(0.5 * (v0s + v1s + 2 * amax * len));
    return (vp - v0) / amax + (vp - v1) / amax;
}

// getArcTrapezoid, moma_traj_opt.h:699-733
double trapezoid_arc(double t, double len, double v0, double v1, double vmax, double amax) {
    double v0s = v0 * v0, v1s = v1 * v1;
    const double vms = vmax * vmax;
    if (v0 > vmax) v0s = vms;
    if (v1 > vmax) v1s = vms;
    const double critical = (vms - v0s) / (2 * amax) + (vms - v1s) / (2 * amax);
    if (len >= critical) {
        const double t1 = (vmax - v0) / amax;
        const double t2 = t1 + (len - critical) / vmax;
        if (t <= t1) return v0 * t + 0.5 * amax * (t * t);
        if (t <= t2) return v0 * t1 + 0.5 * amax * (t1 * t1) + (t - t1) * vmax;
        return v0 * t1 + 0.5 * amax * (t1 * t1) + (t2 - t1) * vmax + vmax * (t - t2) -
               0.5 * amax * (t - t2) * (t - t2);
    }
    const double vp = std::sqrt(0.5 * (v0s + v1s + 2 * amax * len));
    const double tp = (vp - v0) / amax;
    if (t <= tp) return v0 * t + 0.5 * amax * (t * t);
    return v0 * tp + 0.5 * amax * (tp * tp) + vp * (t - tp) - 0.5 * amax * (t - tp) * (t - tp);
}

// normalizeAngle, moma_traj_opt.h:735-742
void unwrap_towards(double ref, double& a) {
    while (ref - a > M_PI) a += 2 * M_PI;
    while (ref - a < -M_PI) a -= 2 * M_PI;
}

double inv_sigmoidC2(double q, double qmax) {   // moma_traj_opt.h:796-800
    const double b = 0.5 * (qmax + q) / qmax;
    return tp_logC2(b / (1 - b));
}

}  // namespace

extern "C" int topay_prepare_candidate(const topay_opt_params* opt, const topay_robot_params* robot,
                                       const double* init_path, int path_len, const double* bvel,
                                       const double* bacc, int max_pieces, int32_t* piece_num, double* head_pva,
                                       double* tail_pva, double* start_xy, double* end_xy, double* init_inner_xy,
                                       double* x0, int32_t* s1_past) {
    if (!opt || !robot || !init_path || path_len < 1 || !bvel || !bacc || !piece_num) return TOPAY_ERR_INVALID_ARG;
    auto wp = [&](int i) { return init_path + (size_t)i * 10; };

    // 1. split every hop into rotate / translate / rotate states (:149-213)
    std::vector<Waypoint> sp;
    {
        Waypoint w{};
        w.x = wp(0)[0];
        w.y = wp(0)[1];
        w.yaw = wp(0)[2];
        std::memcpy(w.q, wp(0) + 3, sizeof(w.q));
        sp.push_back(w);
    }
    for (int i = 1; i < path_len; i++) {
        const double* cur = wp(i);
        const double* prv = wp(i - 1);
        const double ddx = cur[0] - prv[0], ddy = cur[1] - prv[1];
        const double arc = std::sqrt(ddx * ddx + ddy * ddy);
        double yaw = cur[2];
        unwrap_towards(sp.back().yaw, yaw);
        double dyaw = yaw - sp.back().yaw;
        Waypoint w{};
        if (std::fabs(dyaw) > 1e-2) {
            if (arc < 1e-2) {
                w.x = cur[0]; w.y = cur[1]; w.yaw = yaw; w.dyaw = dyaw; w.darc = 0.0;
                std::memcpy(w.q, cur + 3, sizeof(w.q));
                sp.push_back(w);
            } else {
                w = sp.back();
                double heading = std::atan2(cur[1] - sp.back().y, cur[0] - sp.back().x);
                unwrap_towards(sp.back().yaw, heading);
                dyaw = heading - sp.back().yaw;
                w.yaw = heading; w.dyaw = dyaw; w.darc = 0.0;
                sp.push_back(w);

                w.x = cur[0]; w.y = cur[1]; w.yaw = heading; w.dyaw = 0.0; w.darc = arc;
                std::memcpy(w.q, cur + 3, sizeof(w.q));
                sp.push_back(w);

                unwrap_towards(sp.back().yaw, yaw);
                dyaw = yaw - sp.back().yaw;
                w.yaw = yaw; w.dyaw = dyaw; w.darc = 0.0;
                sp.push_back(w);
            }
        } else if (arc > 1e-2) {
            w.x = cur[0]; w.y = cur[1]; w.yaw = yaw; w.dyaw = 0.0; w.darc = arc;
            std::memcpy(w.q, cur + 3, sizeof(w.q));
            sp.push_back(w);
        }
    }

    // 2. cumulative arc lengths, plain and weighted (:224-239)
    const size_t np = sp.size();
    std::vector<double> arcs(np, 0.0), warcs(np, 0.0);
    {
        double a = 0.0, wa = 0.0;
        for (size_t i = 1; i < np; i++) {
            a += sp[i].darc;
            arcs[i] = a;
            wa += 0.2 * std::fabs(sp[i].dyaw) + 1.4 * std::fabs(sp[i].darc);
            warcs[i] = wa;
        }
    }
    const double wlen = warcs.back();
    const double v_start = bvel[0 * 2 + 0];
    const double total_time = trapezoid_duration(wlen, v_start, 0.0, robot->max_v, robot->max_a);
    const int want = std::max(int(total_time / opt->sample_interval + 0.5), (int)opt->min_piece_num);
    const double dt = total_time / want;

    // 3. resample at dt along the trapezoid profile (:244-277)
    std::vector<double> inner;      // 9 per inner point
    std::vector<double> inner_xy;   // 2 per piece
    size_t at = 1;
    for (double t = dt; t < total_time - 1e-3; t += dt) {
        const double arc = trapezoid_arc(t, wlen, v_start, 0.0, robot->max_v, robot->max_a);
        for (size_t k = at; k < np; k++) {
            if (warcs[k] >= arc) {
                at = k;
                const Waypoint &b = sp[k], &a = sp[k - 1];
                const double l1 = warcs[k] - arc;
                const double l = warcs[k] - warcs[k - 1];
                inner.push_back(a.yaw + (l - l1) / l * (b.dyaw));
                inner.push_back(arcs[k - 1] + (l - l1) / l * (b.darc));
                for (int q = 0; q < TOPAY_DOF; q++) inner.push_back(a.q[q] + (l - l1) / l * (b.q[q] - a.q[q]));
                inner_xy.push_back(l1 / l * a.x + (l - l1) / l * (b.x));
                inner_xy.push_back(l1 / l * a.y + (l - l1) / l * (b.y));
                break;
            }
        }
    }
    const double* last = wp(path_len - 1);
    inner_xy.push_back(last[0]);
    inner_xy.push_back(last[1]);

    const int N = (int)(inner.size() / 9) + 1;
    *piece_num = N;
    if (N > max_pieces) return TOPAY_ERR_TOO_LARGE;

    // 4. boundary PVA (:281-297); bvel / bacc are 10 x 2 row-major, row 0 = v, row 1 = w, rows 3.. = joints
    std::memset(head_pva, 0, 27 * sizeof(double));
    std::memset(tail_pva, 0, 27 * sizeof(double));
    head_pva[0 * 3 + 0] = sp.front().yaw;
    head_pva[0 * 3 + 1] = bvel[1 * 2 + 0];
    head_pva[0 * 3 + 2] = bacc[1 * 2 + 0];
    head_pva[1 * 3 + 1] = bvel[0 * 2 + 0];
    head_pva[1 * 3 + 2] = bacc[0 * 2 + 0];
    tail_pva[0 * 3 + 0] = sp.back().yaw;
    tail_pva[1 * 3 + 0] = arcs.back();
    for (int q = 0; q < TOPAY_DOF; q++) {
        head_pva[(2 + q) * 3 + 0] = sp.front().q[q];
        head_pva[(2 + q) * 3 + 1] = bvel[(3 + q) * 2 + 0];
        head_pva[(2 + q) * 3 + 2] = bacc[(3 + q) * 2 + 0];
        tail_pva[(2 + q) * 3 + 0] = sp.back().q[q];
        tail_pva[(2 + q) * 3 + 1] = bvel[(3 + q) * 2 + 1];
        tail_pva[(2 + q) * 3 + 2] = bacc[(3 + q) * 2 + 1];
    }
    start_xy[0] = wp(0)[0];
    start_xy[1] = wp(0)[1];
    end_xy[0] = last[0];
    end_xy[1] = last[1];
    std::memset(init_inner_xy, 0, (size_t)max_pieces * 2 * sizeof(double));
    std::memcpy(init_inner_xy, inner_xy.data(), inner_xy.size() * sizeof(double));

    // 5. pack x (:324-344)
    double* tau = x0;
    double* theta = tau + N;
    double* arcv = theta + (N - 1);
    double* vq = arcv + N;
    const double tau0 = tp_logC2(dt);
    for (int i = 0; i < N - 1; i++) {
        tau[i] = tau0;
        theta[i] = inner[(size_t)i * 9 + 0];
        arcv[i] = inner[(size_t)i * 9 + 1];
        for (int q = 0; q < TOPAY_DOF; q++)
            vq[(size_t)i * TOPAY_DOF + q] = inv_sigmoidC2(inner[(size_t)i * 9 + 2 + q], robot->joint_pos_limit_max[q]);
    }
    tau[N - 1] = tau0;
    arcv[N - 1] = tail_pva[1 * 3 + 0];
    if (s1_past)   // :354-357
        *s1_past = std::fabs(tail_pva[1 * 3 + 0]) < opt->s1_shot_path_horizon ? opt->s1_lbfgs_shot_path_past
                                                                                : opt->s1_lbfgs_normal_past;
    return TOPAY_OK;
}

// MomaParam::MomaParam(), moma_param.h:72-144
extern "C" void topay_robot_params_default(topay_robot_params* rp) {
    std::memset(rp, 0, sizeof(*rp));
    rp->chassis_height = 0.155;
    rp->chassis_colli_radius = 0.4;
    rp->max_v = 1.0;
    rp->max_a = 0.8;
    rp->max_w = 1.25;
    rp->max_dw = 1.0;
    const double link[TOPAY_DOF + 1] = {0.139, 0.1015, 0.1525, 0.1035, 0.1285, 0.0815, 0.144, 0.05};
    const double off[2 * (TOPAY_DOF + 1)] = {0.139 - 0.09, 0.139, 0.0, 0.1015, 0.1525 - 0.08, 0.1525, 0.0, 0.1035,
                                             0.1285 - 0.07, 0.1285, 0.0, 0.0815, 0.144 - 0.07, 0.144, 0.0, 0.1};
    const double rad[2 * (TOPAY_DOF + 1)] = {0.06, 0.06, 0.0, 0.08, 0.04, 0.04, 0.0, 0.07,
                                             0.035, 0.035, 0.0, 0.06, 0.035, 0.035, 0.0, 0.08};
    const double cylinder_radius = 0.055;
    for (int i = 0; i < TOPAY_DOF + 1; i++) rp->colli_length[i] = link[i];
    for (int i = 0; i < 2 * (TOPAY_DOF + 1); i++) {
        rp->colli_points[i] = off[i];
        double r = rad[i];
        if (r > 1e-4 && r < cylinder_radius) r = cylinder_radius;
        rp->colli_point_radius[i] = r;
    }
    const double qmax[TOPAY_DOF] = {3.1, 2.26, 3.1, 2.355, 3.1, 2.23, 6.28};
    for (int i = 0; i < TOPAY_DOF; i++) {
        rp->joint_pos_limit_max[i] = qmax[i];
        rp->joint_vel_limit[i] = 2.35;
        rp->joint_acc_limit[i] = 6.28;
    }
    const double c = 0.7071068;
    const double R[9] = {c, c, 0.0, -c, c, 0.0, 0.0, 0.0, 1.0};
    std::memcpy(rp->relative_R, R, sizeof(R));
    rp->relative_t[0] = 0.0;
    rp->relative_t[1] = 0.115;
    rp->relative_t[2] = 0.016;
    // zero-pose sphere layout: every frame axis is +z, centres on the line (0, 0.115, z)
    double zc[TOPAY_NSPHERE], rc[TOPAY_NSPHERE];
    int n = 0;
    double zbase = rp->chassis_height + rp->relative_t[2];
    for (int i = 0; i < TOPAY_DOF + 1; i++) {
        for (int j = 0; j < 2; j++) {
            if (rp->colli_points[2 * i + j] == 0.0) continue;
            zc[n] = zbase + rp->colli_points[2 * i + j];
            rc[n] = rp->colli_point_radius[2 * i + j];
            n++;
        }
        zbase += rp->colli_length[i];
    }
    for (int i = 0; i < TOPAY_NSPHERE * TOPAY_NSPHERE; i++) rp->collision_matrix[i] = -1;
    for (int i = 0; i < n; i++)
        for (int j = i; j < n; j++) {
            if (i == j) rp->collision_matrix[i * TOPAY_NSPHERE + j] = 1;
            if (std::fabs(zc[i] - zc[j]) < rc[i] + rc[j])
                rp->collision_matrix[i * TOPAY_NSPHERE + j] = rp->collision_matrix[j * TOPAY_NSPHERE + i] = 1;
        }
}

// src/planner/params/optimizer.yaml + lbfgs.hpp:13-129 defaults
extern "C" void topay_opt_params_default(topay_opt_params* o) {
    std::memset(o, 0, sizeof(*o));
    o->int_K = 12;
    o->min_piece_num = 3;
    o->relu_mu = 1.0e-3;
    o->sample_interval = 1.5;
    o->energy_weights[0] = 0.33;
    for (int i = 1; i < TOPAY_DIM; i++) o->energy_weights[i] = 1.0;
    topay_lbfgs_params lb;
    lb.mem_size = 256;
    lb.g_epsilon = 0.0;
    lb.past = 3;
    lb.delta = 1.0e-4;
    lb.max_iterations = 8000;
    lb.max_linesearch = 64;
    lb.min_step = 1.0e-32;
    lb.max_step = 1.0e+20;
    lb.f_dec_coeff = 1.0e-4;
    lb.s_curv_coeff = 0.9;
    lb.cautious_factor = 1.0e-6;
    lb.machine_prec = 1.0e-16;
    o->s1_time_weight = 20.0;
    o->s1_moment_weight = o->s1_acc_weight = o->s1_domega_weight = o->s1_mean_time_weight = 1000.0;
    o->s1_path_pos_weight = 200000.0;
    o->s1_lbfgs_normal_past = 2;
    o->s1_lbfgs_shot_path_past = 8;
    o->s1_shot_path_horizon = 0.5;
    o->s1_lbfgs = lb;
    o->s1_lbfgs.past = 2;
    o->s1_lbfgs.min_step = 0.0;
    o->s1_lbfgs.delta = 1.0e-2;
    o->s2_time_weight = 50.0;
    o->s2_moment_weight = 300.0;
    o->s2_acc_weight = o->s2_domega_weight = 3000.0;
    o->s2_collision_weight = o->s2_mani_colli_weight = o->s2_self_colli_weight = 500000.0;
    o->s2_mani_pos_weight = o->s2_mani_vel_weight = o->s2_mani_acc_weight = 500.0;
    o->s2_mean_time_weight = 5000.0;
    o->s2_lbfgs = lb;
    for (int i = 0; i < 2; i++) {
        o->alm_init_lambda[i] = 0.0;
        o->alm_init_rho[i] = 10000.0;
        o->alm_rho_max[i] = 1.0e+10;
        o->alm_gamma[i] = 9.0;
    }
    o->alm_tolerance = 0.01;
    o->alm_max_rounds = 20;
}

// ------------------------------------------------------------------------------------------------------------
// GraphSearch::getDensePath (src/planner/src/graph_search.cpp:119-176, normalizeAngle :6-13). Two passes over the
// waypoints instead of the reference's three vectors: rows are produced in order and filtered on the fly (a row's
// dt is final once the next row has been started). Same operations on the same operands, so the rows are
// bit-identical.
namespace {
struct DenseRows {
    double* out;
    int cap, count;
    double cur[4];
    bool has;
    void flush(bool last) {
        if (!has) return;
        if (last || cur[3] > 1.0e-3) {
            if (count < cap) memcpy(out + 4 * (size_t)count, cur, sizeof(cur));
            count++;
        }
    }
    void start(double x, double y, double th) {
        flush(false);
        cur[0] = x; cur[1] = y; cur[2] = th; cur[3] = 0.0;
        has = true;
    }
};
inline void unwrap(double ref, double& a) {
    while (ref - a > M_PI) a += 2 * M_PI;
    while (ref - a < -M_PI) a -= 2 * M_PI;
}
}  // namespace

extern "C" int topay_dense_path(const double* raw, int n, double step_size, double start_yaw, double end_yaw,
                                double v_max, double w_max, double* out, int cap) {
    if (!raw || n < 2 || !(step_size > 0.0) || !(v_max > 0.0) || !(w_max > 0.0) || cap < 0 || (cap > 0 && !out))
        return TOPAY_ERR_INVALID_ARG;
    // the dense polyline, point by point
    struct Walker {
        const double* raw;
        int n, i, j, times;
        double step, dx, dy, step_size;
        void segment() {
            const double ex = raw[2 * i] - raw[2 * i - 2], ey = raw[2 * i + 1] - raw[2 * i - 1];
            const double z = ex * ex + ey * ey, len = sqrt(z);
            dx = ex;
            dy = ey;
            if (z > 0.0) {
                dx = ex / len;
                dy = ey / len;
            }
            times = (int)std::max(ceil(len / step_size), 1.0);
            step = len / times;
            j = 1;
        }
        bool next(double* p) {          // false after the last point
            if (i >= n) return false;
            p[0] = raw[2 * i - 2] + (step * j) * dx;
            p[1] = raw[2 * i - 1] + (step * j) * dy;
            if (++j > times && ++i < n) segment();
            return true;
        }
    } w{raw, n, 1, 1, 1, 0.0, 0.0, 0.0, step_size};
    w.segment();
    DenseRows rows{out, cap, 0, {0, 0, 0, 0}, false};
    double p0[2] = {raw[0], raw[1]}, p1[2], p2[2];
    w.next(p1);                                         // dense[1] exists: every segment has at least one step
    rows.start(p0[0], p0[1], start_yaw);
    double th = atan2(p1[1] - p0[1], p1[0] - p0[0]);
    unwrap(start_yaw, th);
    rows.cur[3] = fabs(th - start_yaw) / w_max;
    rows.start(p0[0], p0[1], th);
    // interior points: a move row ending at the point, then a turn row towards the next point
    while (w.next(p2)) {
        const double ax = p1[0] - rows.cur[0], ay = p1[1] - rows.cur[1];
        rows.cur[3] = sqrt(ax * ax + ay * ay) / v_max;
        const double keep = rows.cur[2];
        rows.start(p1[0], p1[1], keep);
        th = atan2(p2[1] - p1[1], p2[0] - p1[0]);
        unwrap(keep, th);
        rows.cur[3] = fabs(th - keep) / w_max;
        rows.start(p1[0], p1[1], th);
        p1[0] = p2[0];
        p1[1] = p2[1];
    }
    // the last point: move, then turn to the goal yaw
    {
        const double ax = p1[0] - rows.cur[0], ay = p1[1] - rows.cur[1];
        rows.cur[3] = sqrt(ax * ax + ay * ay) / v_max;
        const double keep = rows.cur[2];
        rows.start(p1[0], p1[1], keep);
        th = end_yaw;
        unwrap(keep, th);
        rows.cur[3] = fabs(th - keep) / w_max;
        rows.start(p1[0], p1[1], th);
        rows.flush(true);
    }
    return rows.count;
}

// ------------------------------------------------------------------------------------------------------------
// TopologyPRM::pathLength / discretizePath (src/planner/src/topo_prm.cpp:462-506): pt_num points at equal arc-length
// spacing along a 3-D polyline. The reference looks the segment of every point up from the start of the path with
// a 1e-4 tolerance window and takes the first hit; arc length only grows along the points, so a moving index
// finds the same segment. Same arithmetic per point, bit-identical.
extern "C" double topay_path_length(const double* path, int n) {
    double length = 0.0;
    if (!path || n < 2) return length;
    for (int i = 0; i + 1 < n; i++) {
        const double dx = path[3 * i + 3] - path[3 * i], dy = path[3 * i + 4] - path[3 * i + 1],
                     dz = path[3 * i + 5] - path[3 * i + 2];
        length += sqrt(dx * dx + dy * dy + dz * dz);
    }
    return length;
}

extern "C" int topay_discretize_path(const double* path, int n, int pt_num, double* out) {
    if (!path || !out || n < 2 || pt_num < 2) return TOPAY_ERR_INVALID_ARG;
    std::vector<double> cum(n);
    cum[0] = 0.0;
    for (int i = 0; i + 1 < n; i++) {
        const double dx = path[3 * i + 3] - path[3 * i], dy = path[3 * i + 4] - path[3 * i + 1],
                     dz = path[3 * i + 5] - path[3 * i + 2];
        cum[i + 1] = sqrt(dx * dx + dy * dy + dz * dz) + cum[i];
    }
    const double dl = cum[n - 1] / double(pt_num - 1);
    int j = 0;
    for (int i = 0; i < pt_num; i++) {
        const double cur_l = double(i) * dl;
        while (j < n - 2 && !(cur_l >= cum[j] - 1e-4 && cur_l <= cum[j + 1] + 1e-4)) j++;
        const double lambda = (cur_l - cum[j]) / (cum[j + 1] - cum[j]);
        for (int k = 0; k < 3; k++) out[3 * i + k] = (1 - lambda) * path[3 * j + k] + lambda * path[3 * j + 3 + k];
    }
    return TOPAY_OK;
}
