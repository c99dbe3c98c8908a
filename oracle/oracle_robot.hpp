// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the robot model used inside the NLP solve:
//   src/simulator/fake_moma/include/fake_moma/moma_param.h:72-144   constants -> robot_defaults
//   src/simulator/fake_moma/include/fake_moma/moma_param.h:203-247  getColliPts
//   src/simulator/fake_moma/include/fake_moma/moma_param.h:249-337  getColliGrads
//   src/map/include/map/grid_map.h:613-650                          isWholeBodyCollision
// PARITY UNPINNED (see oracle_field.hpp); pinned by finite differences of the
// FK in tests/test_oracle_robot.py.
#pragma once
#include <cmath>
#include <cstring>

#include "../include/topay_b200.h"
#include "oracle_field.hpp"

namespace oracle {

struct M3 {
    double a[9];  // row-major
    double& operator()(int r, int c) { return a[3 * r + c]; }
    double operator()(int r, int c) const { return a[3 * r + c]; }
    static M3 identity() {
        M3 m;
        std::memset(m.a, 0, sizeof(m.a));
        m.a[0] = m.a[4] = m.a[8] = 1.0;
        return m;
    }
};
// Eigen's fixed-size product: res(i,j) = sum_k lhs(i,k) * rhs(k,j), k ascending.
inline M3 mul(const M3& A, const M3& B) {
    M3 R;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = A(i, 0) * B(0, j);
            s += A(i, 1) * B(1, j);
            s += A(i, 2) * B(2, j);
            R(i, j) = s;
        }
    return R;
}
inline void mulv(const M3& A, const double* v, double* out) {
    for (int i = 0; i < 3; i++) {
        double s = A(i, 0) * v[0];
        s += A(i, 1) * v[1];
        s += A(i, 2) * v[2];
        out[i] = s;
    }
}
inline double dot3(const double* a, const double* b) {
    double s = a[0] * b[0];
    s += a[1] * b[1];
    s += a[2] * b[2];
    return s;
}

inline M3 rot_z(double q) {
    M3 R;
    R(0, 0) = std::cos(q); R(0, 1) = -std::sin(q); R(0, 2) = 0.0;
    R(1, 0) = std::sin(q); R(1, 1) = std::cos(q);  R(1, 2) = 0.0;
    R(2, 0) = 0.0;         R(2, 1) = 0.0;          R(2, 2) = 1.0;
    return R;
}
inline M3 drot_z(double q) {
    M3 R;
    R(0, 0) = -std::sin(q); R(0, 1) = -std::cos(q); R(0, 2) = 0.0;
    R(1, 0) = std::cos(q);  R(1, 1) = -std::sin(q); R(1, 2) = 0.0;
    R(2, 0) = 0.0;          R(2, 1) = 0.0;          R(2, 2) = 0.0;
    return R;
}
inline M3 rot_y(double q) {
    M3 R;
    R(0, 0) = std::cos(q);  R(0, 1) = 0.0; R(0, 2) = std::sin(q);
    R(1, 0) = 0.0;          R(1, 1) = 1.0; R(1, 2) = 0.0;
    R(2, 0) = -std::sin(q); R(2, 1) = 0.0; R(2, 2) = std::cos(q);
    return R;
}
inline M3 drot_y(double q) {
    M3 R;
    R(0, 0) = -std::sin(q); R(0, 1) = 0.0; R(0, 2) = std::cos(q);
    R(1, 0) = 0.0;          R(1, 1) = 0.0; R(1, 2) = 0.0;
    R(2, 0) = -std::cos(q); R(2, 1) = 0.0; R(2, 2) = -std::sin(q);
    return R;
}
inline M3 as_m3(const double* r) {
    M3 m;
    for (int i = 0; i < 9; i++) m.a[i] = r[i];
    return m;
}

// moma_param.h:203-247. moma_pos = (x, y, yaw, q1..q7). Returns the number of
// spheres written to pts (x,y,z,r each).
inline int get_colli_pts(const topay_robot_params& rp, const double* moma_pos, double pts[][4]) {
    int n = 0;
    double now_p[3] = {moma_pos[0], moma_pos[1], rp.chassis_height};
    M3 now_R = rot_z(moma_pos[2]);
    double t[3];
    mulv(now_R, rp.relative_t, t);
    for (int i = 0; i < 3; i++) now_p[i] += t[i];
    now_R = mul(now_R, as_m3(rp.relative_R));
    for (int i = 0; i < TOPAY_DOF + 1; i++) {
        for (int j = 0; j < 2; j++) {
            if (rp.colli_points[i * 2 + j] == 0.0) continue;
            for (int d = 0; d < 3; d++) pts[n][d] = now_p[d] + now_R(d, 2) * rp.colli_points[i * 2 + j];
            pts[n][3] = rp.colli_point_radius[i * 2 + j];
            n++;
        }
        for (int d = 0; d < 3; d++) now_p[d] += now_R(d, 2) * rp.colli_length[i];
        if (i == TOPAY_DOF) break;
        M3 dof_R = (i % 2 == 0) ? rot_z(1.0 * moma_pos[3 + i]) : rot_y(1.0 * moma_pos[3 + i]);
        now_R = mul(now_R, dof_R);
    }
    return n;
}

// moma_param.h:249-337, including the O(dof^3) explicit chain products.
inline void get_colli_grads(const topay_robot_params& rp, const double* moma_pos,
                            const double pos_grads[][3], double* colli_grads /*10*/) {
    const int dof = TOPAY_DOF;
    int gidx = 0;
    for (int i = 0; i < 3 + dof; i++) colli_grads[i] = 0.0;
    double grad_p_list[TOPAY_DOF + 1][3] = {}, grad_R_list[TOPAY_DOF + 1][3] = {};
    M3 now_R_list[TOPAY_DOF + 1], R_list[TOPAY_DOF + 1], dR_list[TOPAY_DOF + 1];
    double now_p[3] = {moma_pos[0], moma_pos[1], rp.chassis_height};
    M3 now_R = rot_z(moma_pos[2]);
    double t[3];
    mulv(now_R, rp.relative_t, t);
    for (int i = 0; i < 3; i++) now_p[i] += t[i];
    now_R = mul(now_R, as_m3(rp.relative_R));
    int nR = 0;
    R_list[nR] = now_R;
    M3 dR = drot_z(moma_pos[2]);
    dR = mul(dR, as_m3(rp.relative_R));
    dR_list[nR] = dR;
    nR++;
    for (int i = 0; i < dof + 1; i++) {
        now_R_list[i] = now_R;
        for (int j = 0; j < 2; j++) {
            if (rp.colli_points[i * 2 + j] == 0.0) continue;
            for (int d = 0; d < 3; d++) {
                grad_p_list[i][d] += pos_grads[gidx][d];
                grad_R_list[i][d] += rp.colli_points[i * 2 + j] * pos_grads[gidx][d];
            }
            gidx++;
        }
        if (i == dof) break;
        for (int d = 0; d < 3; d++) now_p[d] += now_R(d, 2) * rp.colli_length[i];
        M3 dof_R, ddR;
        if (i % 2 == 0) {
            dof_R = rot_z(moma_pos[3 + i]);
            ddR = drot_z(moma_pos[3 + i]);
        } else {
            dof_R = rot_y(moma_pos[3 + i]);
            ddR = drot_y(moma_pos[3 + i]);
        }
        now_R = mul(now_R, dof_R);
        R_list[nR] = dof_R;
        dR_list[nR] = ddR;
        nR++;
    }
    for (int i = dof; i > 0; i--) {
        {
            double col[3] = {dR_list[i](0, 2), dR_list[i](1, 2), dR_list[i](2, 2)};
            double v[3];
            mulv(now_R_list[i - 1], col, v);
            colli_grads[2 + i] += dot3(v, grad_R_list[i]);
        }
        for (int j = 0; j < i; j++) {
            M3 dRj = dR_list[j];
            M3 R1 = M3::identity();
            for (int k = 0; k < j; k++) R1 = mul(R1, R_list[k]);
            R1 = mul(R1, dRj);
            for (int k = j + 1; k <= i; k++) R1 = mul(R1, R_list[k]);
            double col[3] = {R1(0, 2), R1(1, 2), R1(2, 2)};
            colli_grads[2 + j] += dot3(col, grad_R_list[i]);
        }
        for (int d = 0; d < 3; d++) {
            grad_R_list[i - 1][d] += grad_p_list[i][d] * rp.colli_length[i - 1];
            grad_p_list[i - 1][d] += grad_p_list[i][d];
        }
    }
    {
        double col[3] = {dR_list[0](0, 2), dR_list[0](1, 2), dR_list[0](2, 2)};
        colli_grads[2] += dot3(col, grad_R_list[0]);
    }
    dR = drot_z(moma_pos[2]);
    double v[3];
    mulv(dR, rp.relative_t, v);
    colli_grads[2] += dot3(v, grad_p_list[0]);
    colli_grads[0] = grad_p_list[0][0];
    colli_grads[1] = grad_p_list[0][1];
}

// moma_param.h:72-144
inline void robot_defaults(topay_robot_params* rp) {
    rp->chassis_height = 0.155;
    rp->chassis_colli_radius = 0.4;
    rp->max_v = 1.0;
    rp->max_a = 0.8;
    rp->max_w = 1.25;
    rp->max_dw = 1.0;
    const double cylinder_radius = 0.055;
    const double cl[8] = {0.139, 0.1015, 0.1525, 0.1035, 0.1285, 0.0815, 0.144, 0.05};
    const double cp[16] = {0.139 - 0.09, 0.139, 0.0, 0.1015, 0.1525 - 0.08, 0.1525, 0.0, 0.1035,
                           0.1285 - 0.07, 0.1285, 0.0, 0.0815, 0.144 - 0.07, 0.144, 0.0, 0.1};
    const double cr[16] = {0.06, 0.06, 0.0, 0.08, 0.04, 0.04, 0.0, 0.07,
                           0.035, 0.035, 0.0, 0.06, 0.035, 0.035, 0.0, 0.08};
    for (int i = 0; i < 8; i++) rp->colli_length[i] = cl[i];
    for (int i = 0; i < 16; i++) {
        rp->colli_points[i] = cp[i];
        rp->colli_point_radius[i] = cr[i];
        if (rp->colli_point_radius[i] > 1e-4 && rp->colli_point_radius[i] < cylinder_radius)
            rp->colli_point_radius[i] = cylinder_radius;
    }
    const double qmax[7] = {3.1, 2.26, 3.1, 2.355, 3.1, 2.23, 6.28};
    for (int i = 0; i < 7; i++) {
        rp->joint_pos_limit_max[i] = qmax[i];
        rp->joint_vel_limit[i] = 2.35;
        rp->joint_acc_limit[i] = 6.28;
    }
    const double rr[9] = {0.7071068, 0.7071068, 0.0, -0.7071068, 0.7071068, 0.0, 0.0, 0.0, 1.0};
    for (int i = 0; i < 9; i++) rp->relative_R[i] = rr[i];
    rp->relative_t[0] = 0.0;
    rp->relative_t[1] = 0.115;
    rp->relative_t[2] = 0.016;
    // collision_matrix from the zero pose (moma_param.h:128-143)
    double zero[10] = {0};
    double cpts[TOPAY_NSPHERE][4];
    const int n = get_colli_pts(*rp, zero, cpts);
    for (int i = 0; i < TOPAY_NSPHERE * TOPAY_NSPHERE; i++) rp->collision_matrix[i] = -1;
    for (int i = 0; i < n; i++)
        for (int j = i; j < n; j++) {
            if (i == j) rp->collision_matrix[i * TOPAY_NSPHERE + j] = 1;
            double d[3] = {cpts[i][0] - cpts[j][0], cpts[i][1] - cpts[j][1], cpts[i][2] - cpts[j][2]};
            double dist = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (dist < cpts[i][3] + cpts[j][3])
                rp->collision_matrix[i * TOPAY_NSPHERE + j] = rp->collision_matrix[j * TOPAY_NSPHERE + i] = 1;
        }
}

// grid_map.h:613-650 (dense path)
inline bool whole_body_collision(const Field& f, const topay_robot_params& rp, const double* state) {
    for (int i = 0; i < TOPAY_DOF; i++)
        if (state[i + 3] > rp.joint_pos_limit_max[i] || state[i + 3] < -rp.joint_pos_limit_max[i])
            return true;
    if (f.is_collision_2d(state, rp.chassis_colli_radius)) return true;
    double pts[TOPAY_NSPHERE][4];
    const int n = get_colli_pts(rp, state, pts);
    for (int i = 0; i < n; i++) {
        if (f.is_collision_3d(pts[i], pts[i][3])) return true;
        if (i > 2 && pts[i][2] < rp.chassis_height + pts[i][3]) {
            double dx = pts[i][0] - state[0], dy = pts[i][1] - state[1];
            if (std::sqrt(dx * dx + dy * dy) < rp.chassis_colli_radius + pts[i][3]) return true;
        }
        for (int j = i + 1; j < n; j++) {
            double d[3] = {pts[i][0] - pts[j][0], pts[i][1] - pts[j][1], pts[i][2] - pts[j][2]};
            double dist = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (dist < pts[i][3] + pts[j][3] && rp.collision_matrix[i * TOPAY_NSPHERE + j] == -1)
                return true;
        }
    }
    return false;
}

}  // namespace oracle
