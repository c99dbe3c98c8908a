// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the reference's result post-processing and success gate (SURVEY.md §8
// row a16 and "next" row N1): PolyTrajectory<9,5> evaluation, the MomaTraj Simpson pose table
// and getState/getDState, MomaTrajOpt::checkFeasible / printConstraintsSituations and the
// planner's shortest-duration selection. Plain C++17, the reference's operation order.
// PARITY UNPINNED (no reference tests for this path; pinned by analytic known answers in
// tests/test_oracle_traj.py).
//
// Reference files restated here (paths relative to /root/reference/src/planner):
//   include/utils/minco.hpp:112-147, 304-313, 356-392   Piece::getPos/getVel/getAcc, locatePieceIdx
//   include/utils/minco.hpp:908-921                     MinJerkOpt::getTraj (descending-power re-ordering)
//   include/planner/moma_traj_opt.h:26-68               MomaTraj constructor (car_seq)
//   include/planner/moma_traj_opt.h:121-160             getState / getDState
//   include/planner/moma_traj_opt.h:948-1045            checkFeasible
//   include/planner/moma_traj_opt.h:1047-1210           printConstraintsSituations (the verdict only)
//   src/planner.cpp:877-880, 999-1010                   success gate, shortest-duration selection
//
// Deviation, stated: getState clamps the car_seq index to the table (the reference indexes
// car_seq[floor(t / 0.1)] unchecked; the two floors can disagree by one at t = total duration).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "oracle_field.hpp"
#include "oracle_robot.hpp"
#include "oracle_rog.hpp"

namespace oracle {

struct PolyTraj {
    int N = 0;
    const double* T = nullptr;   // [N]
    const double* c = nullptr;   // row 6i+k = coefficient of t^k of piece i, 9 columns (MinJerkOpt layout);
                                 // the reference stores the same numbers in descending order (minco.hpp:917)
    double total() const {       // minco.hpp:304-313
        double s = 0.0;
        for (int i = 0; i < N; i++) s += T[i];
        return s;
    }
    int locate(double& t) const {   // minco.hpp:356-374
        int idx;
        double dur;
        for (idx = 0; idx < N && t > (dur = T[idx]); idx++) t -= dur;
        if (idx == N) {
            idx--;
            t += T[idx];
        }
        return idx;
    }
    void pos(double t, double out[9]) const {   // minco.hpp:112-124
        const int p = locate(t);
        for (int d = 0; d < 9; d++) out[d] = 0.0;
        double tn = 1.0;
        for (int k = 0; k <= 5; k++) {
            for (int d = 0; d < 9; d++) out[d] += tn * c[(6 * p + k) * 9 + d];
            tn *= t;
        }
    }
    void vel(double t, double out[9]) const {   // minco.hpp:126-139
        const int p = locate(t);
        for (int d = 0; d < 9; d++) out[d] = 0.0;
        double tn = 1.0;
        int n = 1;
        for (int k = 1; k <= 5; k++) {
            for (int d = 0; d < 9; d++) out[d] += n * tn * c[(6 * p + k) * 9 + d];
            tn *= t;
            n++;
        }
    }
    void acc(double t, double out[9]) const {   // minco.hpp:141-156
        const int p = locate(t);
        for (int d = 0; d < 9; d++) out[d] = 0.0;
        double tn = 1.0;
        int m = 1, n = 2;
        for (int k = 2; k <= 5; k++) {
            for (int d = 0; d < 9; d++) out[d] += m * n * tn * c[(6 * p + k) * 9 + d];
            tn *= t;
            m++;
            n++;
        }
    }
};

struct MomaTrajO {
    double seq_res = 0.1;
    int approx_res = 4;
    PolyTraj poly;
    double start[3] = {0, 0, 0};
    std::vector<double> car_seq;   // x, y, yaw, t per entry

    // moma_traj_opt.h:38-68
    void init(const PolyTraj& p, const double start_se2[3]) {
        poly = p;
        for (int i = 0; i < 3; i++) start[i] = start_se2[i];
        const double r = seq_res / approx_res, half = r / 2.0, r16 = r / 6.0;
        double cx = start[0], cy = start[1];
        car_seq.clear();
        car_seq.insert(car_seq.end(), {cx, cy, start[2], 0.0});
        const int seq_num = (int)std::floor(poly.total() / r);
        double P[9], V[9];
        poly.pos(0.0, P);
        poly.vel(0.0, V);
        double p3[2] = {P[0], P[1]}, v3[2] = {V[0], V[1]};
        for (int i = 0; i < seq_num; i++) {
            const double p1[2] = {p3[0], p3[1]}, v1[2] = {v3[0], v3[1]};
            poly.pos(i * r + half, P);
            poly.vel(i * r + half, V);
            const double p2[2] = {P[0], P[1]}, v2[2] = {V[0], V[1]};
            poly.pos(i * r + r, P);
            poly.vel(i * r + r, V);
            p3[0] = P[0]; p3[1] = P[1];
            v3[0] = V[0]; v3[1] = V[1];
            cx += r16 * (v1[1] * std::cos(p1[0]) + 4.0 * v2[1] * std::cos(p2[0]) + v3[1] * std::cos(p3[0]));
            cy += r16 * (v1[1] * std::sin(p1[0]) + 4.0 * v2[1] * std::sin(p2[0]) + v3[1] * std::sin(p3[0]));
            if (i % approx_res == approx_res - 1) car_seq.insert(car_seq.end(), {cx, cy, p3[0], (i + 1) * r});
        }
    }
    double total() const { return poly.total(); }

    // moma_traj_opt.h:121-149
    void get_state(double t, double state[10]) const {
        t = std::min(std::max(t, 0.0), total());
        int index = (int)std::floor(t / seq_res);
        const int last = (int)(car_seq.size() / 4) - 1;
        if (index > last) index = last;   // deviation, see the header
        const double floor_t = index * seq_res;
        const double diff_t = t - floor_t;
        double cx = car_seq[4 * index], cy = car_seq[4 * index + 1];
        double P1[9], V1[9], P2[9], V2[9], P3[9], V3[9];
        poly.pos(floor_t, P1);
        poly.vel(floor_t, V1);
        poly.pos(floor_t + diff_t / 2.0, P2);
        poly.vel(floor_t + diff_t / 2.0, V2);
        poly.pos(t, P3);
        poly.vel(t, V3);
        cx += diff_t / 6.0 * (V1[1] * std::cos(P1[0]) + 4.0 * V2[1] * std::cos(P2[0]) + V3[1] * std::cos(P3[0]));
        cy += diff_t / 6.0 * (V1[1] * std::sin(P1[0]) + 4.0 * V2[1] * std::sin(P2[0]) + V3[1] * std::sin(P3[0]));
        state[0] = cx;
        state[1] = cy;
        state[2] = P3[0];
        for (int i = 0; i < 7; i++) state[3 + i] = P3[2 + i];
    }
    // moma_traj_opt.h:151-160
    void get_dstate(double t, double ds[10]) const {
        t = std::min(std::max(t, 0.0), total());
        double V[9];
        poly.vel(t, V);
        ds[0] = V[1];
        ds[1] = V[0];
        ds[2] = 0.0;
        for (int i = 0; i < 7; i++) ds[3 + i] = V[2 + i];
    }
};

struct Feasibility {
    int feasible = 1, feasible_print = 1, n_samples = 0;
    double max_vel = 0, max_acc = 0, max_domega = 0, max_d2omega = 0;
    double max_q[TOPAY_DOF] = {}, max_dq[TOPAY_DOF] = {}, max_d2q[TOPAY_DOF] = {};
    double min_dist = 1.0e+10;
    double min_dist_mani[TOPAY_NSPHERE];
};

// moma_traj_opt.h:948-1045 (checkFeasible) and :1047-1210 (printConstraintsSituations: the same
// scan; the manipulator-clearance test does not clear its verdict, :1199).
inline Feasibility check_feasible(const MomaTrajO& traj, const topay_robot_params& rp, const Field* grid,
                                  const RogEsdf* rog = nullptr) {
    Feasibility F;
    const double res = 0.01;
    double zero[10] = {0};
    double rpts[TOPAY_NSPHERE + 4][4];
    const int nsph = get_colli_pts(rp, zero, rpts);
    for (int i = 0; i < TOPAY_NSPHERE; i++) F.min_dist_mani[i] = 1.0e+10;
    const double total = traj.total();
    for (double t = 0.0; t < total; t += res) {
        F.n_samples++;
        double state[10], vel[9], acc[9];
        traj.get_state(t, state);
        traj.poly.vel(t, vel);
        traj.poly.acc(t, acc);
        if (std::fabs(vel[1]) > std::fabs(F.max_vel)) F.max_vel = vel[1];
        if (std::fabs(acc[1]) > std::fabs(F.max_acc)) F.max_acc = acc[1];
        if (std::fabs(vel[0]) > std::fabs(F.max_domega)) F.max_domega = vel[0];
        if (std::fabs(acc[0]) > std::fabs(F.max_d2omega)) F.max_d2omega = acc[0];
        for (int i = 0; i < TOPAY_DOF; i++) {
            if (std::fabs(state[i + 3]) > std::fabs(F.max_q[i])) F.max_q[i] = state[i + 3];
            if (std::fabs(vel[i + 2]) > std::fabs(F.max_dq[i])) F.max_dq[i] = vel[i + 2];
            if (std::fabs(acc[i + 2]) > std::fabs(F.max_d2q[i])) F.max_d2q[i] = acc[i + 2];
        }
        double d = 0.0;
        if (rog) {   // GridMap::getDistance2d, use_rog branch (grid_map.h:256-267)
            const double p3[3] = {state[0], state[1], 0.0};
            double g3[3];
            rog->value_grad_2d(p3, false, d, g3);
        } else {
            grid->distance2d(state, d);
        }
        if (d < F.min_dist) F.min_dist = d;
        double pts[TOPAY_NSPHERE + 4][4];
        const int n = get_colli_pts(rp, state, pts);
        for (int i = 0; i < n; i++) {
            double dd = 0.0;
            if (rog) dd = rog->evaluate_edt(pts[i]);   // getDistance3d, use_rog branch (grid_map.h:307-322)
            else grid->distance3d(pts[i], dd);
            if (dd < F.min_dist_mani[i]) F.min_dist_mani[i] = dd;
        }
    }
    bool ok = true;
    if (std::fabs(F.max_vel) > 1.01 * rp.max_v) ok = false;
    if (std::fabs(F.max_acc) > 1.01 * rp.max_a) ok = false;
    if (std::fabs(F.max_domega) > 1.01 * rp.max_w) ok = false;
    if (std::fabs(F.max_d2omega) > 1.01 * rp.max_dw) ok = false;
    for (int i = 0; i < TOPAY_DOF; i++) {
        if (std::fabs(F.max_q[i]) > 1.01 * rp.joint_pos_limit_max[i]) ok = false;
        if (std::fabs(F.max_dq[i]) > 1.01 * rp.joint_vel_limit[i]) ok = false;
        if (std::fabs(F.max_d2q[i]) > 1.01 * rp.joint_acc_limit[i]) ok = false;
    }
    if (F.min_dist < 0.99 * rp.chassis_colli_radius) ok = false;
    F.feasible_print = ok ? 1 : 0;
    for (int i = 0; i < nsph; i++)
        if (F.min_dist_mani[i] < 0.99 * rpts[i][3]) ok = false;
    F.feasible = ok ? 1 : 0;
    return F;
}

// planner.cpp:999-1010: first success, replaced by any later strictly shorter one
inline int select_shortest(const int* succ, const double* duration, int n) {
    int best = -1;
    for (int i = 0; i < n; i++) {
        if (!succ[i]) continue;
        if (best == -1) best = i;
        if (duration[i] < duration[best]) best = i;
    }
    return best;
}

}  // namespace oracle
