// Per-thread evaluation of a solved trajectory — PolyTrajectory<9,5> (src/planner/include/utils/
// minco.hpp:112-156, 304-313, 356-392) and MomaTraj::getState / getDState
// (src/planner/include/planner/moma_traj_opt.h:121-160). Coefficients in the solver's layout:
// row 6i+k = coefficient of t^k of piece i, 9 columns (theta, arc, q1..q7).
#pragma once
#include "hd.cuh"

struct TpPoly {
    int N;
    const double* T;   // [N]
    const double* c;   // [6N][9]
};

TP_HD double tp_poly_total(const TpPoly& p) {   // minco.hpp:304-313, summed in piece order
    double s = 0.0;
    for (int i = 0; i < p.N; i++) s += p.T[i];
    return s;
}
TP_HD int tp_poly_locate(const TpPoly& p, double& t) {   // minco.hpp:356-374
    int idx;
    double dur = 0.0;
    for (idx = 0; idx < p.N && t > (dur = p.T[idx]); idx++) t -= dur;
    if (idx == p.N) {
        idx--;
        t += p.T[idx];
    }
    return idx;
}
// position / velocity / acceleration of dimensions [d0, d0 + nd) at time t
TP_HD void tp_poly_pos(const TpPoly& p, double t, int d0, int nd, double* out) {
    const int pc = tp_poly_locate(p, t);
    const double* c = p.c + (size_t)6 * pc * 9;
    for (int d = 0; d < nd; d++) out[d] = 0.0;
    double tn = 1.0;
    for (int k = 0; k <= 5; k++) {
        for (int d = 0; d < nd; d++) out[d] += tn * c[k * 9 + d0 + d];
        tn *= t;
    }
}
TP_HD void tp_poly_vel(const TpPoly& p, double t, int d0, int nd, double* out) {
    const int pc = tp_poly_locate(p, t);
    const double* c = p.c + (size_t)6 * pc * 9;
    for (int d = 0; d < nd; d++) out[d] = 0.0;
    double tn = 1.0;
    for (int k = 1; k <= 5; k++) {
        const double f = k * tn;
        for (int d = 0; d < nd; d++) out[d] += f * c[k * 9 + d0 + d];
        tn *= t;
    }
}
TP_HD void tp_poly_acc(const TpPoly& p, double t, int d0, int nd, double* out) {
    const int pc = tp_poly_locate(p, t);
    const double* c = p.c + (size_t)6 * pc * 9;
    for (int d = 0; d < nd; d++) out[d] = 0.0;
    double tn = 1.0;
    for (int k = 2; k <= 5; k++) {
        const double f = ((k - 1) * k) * tn;
        for (int d = 0; d < nd; d++) out[d] += f * c[k * 9 + d0 + d];
        tn *= t;
    }
}

#define TP_SEQ_RES 0.1        // MomaTraj::seq_res     (moma_traj_opt.h:29)
#define TP_APPROX_RES 4       // MomaTraj::approx_res  (moma_traj_opt.h:30)

// One Simpson step of the planar pose: h/6 * (v1 f(th1) + 4 v2 f(th2) + v3 f(th3)) for cos and sin.
TP_HD void tp_simpson_xy(double h6, const double* pv1, const double* pv2, const double* pv3, double& dx, double& dy) {
    // pvX = {theta, arc', ...}: theta from pos[0], speed from vel[1]
    double s1, c1, s2, c2, s3, c3;
    sincos(pv1[0], &s1, &c1);
    sincos(pv2[0], &s2, &c2);
    sincos(pv3[0], &s3, &c3);
    dx = h6 * (pv1[1] * c1 + 4.0 * pv2[1] * c2 + pv3[1] * c3);
    dy = h6 * (pv1[1] * s1 + 4.0 * pv2[1] * s2 + pv3[1] * s3);
}
// {theta(t), arc'(t)}
TP_HD void tp_theta_speed(const TpPoly& p, double t, double out[2]) {
    double P[1], V[1];
    tp_poly_pos(p, t, 0, 1, P);
    tp_poly_vel(p, t, 1, 1, V);
    out[0] = P[0];
    out[1] = V[0];
}

// MomaTraj::getState (moma_traj_opt.h:121-149). car_seq: entries (x, y, yaw, t); n_seq of them.
TP_HD void tp_traj_state(const TpPoly& p, const double* car_seq, int n_seq, double total, double t, double state[10]) {
    t = fmin(fmax(t, 0.0), total);
    int index = (int)floor(t / TP_SEQ_RES);
    if (index > n_seq - 1) index = n_seq - 1;   // the reference indexes unchecked
    const double floor_t = index * TP_SEQ_RES;
    const double diff_t = t - floor_t;
    double a[2], b[2], c[2], dx, dy;
    tp_theta_speed(p, floor_t, a);
    tp_theta_speed(p, floor_t + diff_t / 2.0, b);
    tp_theta_speed(p, t, c);
    tp_simpson_xy(diff_t / 6.0, a, b, c, dx, dy);
    state[0] = car_seq[4 * index] + dx;
    state[1] = car_seq[4 * index + 1] + dy;
    state[2] = c[0];
    tp_poly_pos(p, t, 2, 7, state + 3);
}
// MomaTraj::getDState (moma_traj_opt.h:151-160)
TP_HD void tp_traj_dstate(const TpPoly& p, double total, double t, double ds[10]) {
    t = fmin(fmax(t, 0.0), total);
    double V[9];
    tp_poly_vel(p, t, 0, 9, V);
    ds[0] = V[1];
    ds[1] = V[0];
    ds[2] = 0.0;
    for (int i = 0; i < 7; i++) ds[3 + i] = V[2 + i];
}
