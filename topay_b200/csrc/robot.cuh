// Forward kinematics of the 12 arm collision spheres and its reverse-mode adjoint.
//
// Same mathematics as MomaParam::getColliPts / getColliGrads
// (src/simulator/fake_moma/include/fake_moma/moma_param.h:203-247, 249-337) but
// organised for one GPU thread per sample: the reference's O(dof^3) explicit chain
// products (moma_param.h:312-328) are replaced by an O(dof) Jacobian-transpose
// sweep. relative_R is not orthonormal (0.7071068 is not sqrt(1/2), moma_param.h:123)
// so the sweep is done in the frame AFTER relative_R, where every factor is a pure
// rotation: with B = Rz(yaw) * relative_R and W_i = R_1 ... R_i,
//     z_i = B * W_i e_z,            d z_i / d q_j = B (w_j x W_i e_z),   w_j = W_j a_j
// and d z_i / d yaw = e_z x z_i exactly.
#pragma once
#include "hd.cuh"

struct TpFK {
    double B[9];                       // Rz(yaw) * relative_R, row-major
    double p0[3];                      // base of the arm: (x, y, h) + Rz(yaw) * relative_t
    double w[TOPAY_DOF + 1][3];        // W_i e_z (frame z axis before B)
    double ax[TOPAY_DOF][3];           // joint axes W_j a_j (before B)
    double z[TOPAY_DOF + 1][3];        // world z axis of each frame, z_i = B w_i
    double sy, cy;                     // sin / cos of yaw
};

// pos = (x, y, yaw, q1..q7). Fills fk and the sphere centres pts[c] = (x, y, z).
TP_HD void tp_fk(const TpParams& P, const double* pos, TpFK& fk, double pts[][3]) {
    const topay_robot_params& rp = P.robot;
    double s, c;
    s = sin(pos[2]);
    c = cos(pos[2]);
    fk.sy = s;
    fk.cy = c;
    const double* R = rp.relative_R;
    // B = Rz * relative_R
    for (int j = 0; j < 3; j++) {
        fk.B[0 + j] = c * R[0 + j] + (-s) * R[3 + j];
        fk.B[3 + j] = s * R[0 + j] + c * R[3 + j];
        fk.B[6 + j] = R[6 + j];
    }
    fk.p0[0] = pos[0] + (c * rp.relative_t[0] + (-s) * rp.relative_t[1]);
    fk.p0[1] = pos[1] + (s * rp.relative_t[0] + c * rp.relative_t[1]);
    fk.p0[2] = rp.chassis_height + rp.relative_t[2];
    // W = I; columns kept separately
    double c0[3] = {1, 0, 0}, c1[3] = {0, 1, 0}, c2[3] = {0, 0, 1};
    double pcur[3] = {fk.p0[0], fk.p0[1], fk.p0[2]};
    int n = 0;
    for (int i = 0; i < TOPAY_DOF + 1; i++) {
        for (int d = 0; d < 3; d++) fk.w[i][d] = c2[d];
        for (int d = 0; d < 3; d++)
            fk.z[i][d] = fk.B[3 * d + 0] * c2[0] + fk.B[3 * d + 1] * c2[1] + fk.B[3 * d + 2] * c2[2];
        while (n < P.n_sphere && P.sphere_frame[n] == i) {
            for (int d = 0; d < 3; d++) pts[n][d] = pcur[d] + fk.z[i][d] * P.sphere_off[n];
            n++;
        }
        if (i == TOPAY_DOF) break;
        for (int d = 0; d < 3; d++) pcur[d] += fk.z[i][d] * rp.colli_length[i];
        double sq, cq;
        sq = sin(pos[3 + i]);
        cq = cos(pos[3 + i]);
        if (i % 2 == 0) {  // Rz(q): axis e_z
            for (int d = 0; d < 3; d++) fk.ax[i][d] = c2[d];
            for (int d = 0; d < 3; d++) {
                const double a = c0[d], b = c1[d];
                c0[d] = a * cq + b * sq;
                c1[d] = a * (-sq) + b * cq;
            }
        } else {  // Ry(q): axis e_y
            for (int d = 0; d < 3; d++) fk.ax[i][d] = c1[d];
            for (int d = 0; d < 3; d++) {
                const double a = c0[d], b = c2[d];
                c0[d] = a * cq + b * (-sq);
                c2[d] = a * sq + b * cq;
            }
        }
    }
}

// g[c] = d cost / d sphere centre c. out = d cost / d (x, y, yaw, q1..q7).
TP_HD void tp_fk_adjoint(const TpParams& P, const TpFK& fk, const double g[][3], double* out) {
    const topay_robot_params& rp = P.robot;
    // per-frame sums
    double gp[TOPAY_DOF + 1][3], gz[TOPAY_DOF + 1][3];
    for (int i = 0; i < TOPAY_DOF + 1; i++)
        for (int d = 0; d < 3; d++) gp[i][d] = gz[i][d] = 0.0;
    for (int n = 0; n < P.n_sphere; n++) {
        const int i = P.sphere_frame[n];
        for (int d = 0; d < 3; d++) {
            gp[i][d] += g[n][d];
            gz[i][d] += P.sphere_off[n] * g[n][d];
        }
    }
    double acc[3] = {0, 0, 0};   // sum of position gradients of frames > i
    double S[3] = {0, 0, 0};     // suffix sum of w_i x (B^T G_i)
    double gyaw = 0.0;
    for (int d = 0; d < 3 + TOPAY_DOF; d++) out[d] = 0.0;
    for (int i = TOPAY_DOF; i >= 0; i--) {
        double G[3];
        for (int d = 0; d < 3; d++) G[d] = gz[i][d] + (i < TOPAY_DOF ? rp.colli_length[i] * acc[d] : 0.0);
        // yaw: e_z . (z_i x G_i)
        gyaw += fk.z[i][0] * G[1] - fk.z[i][1] * G[0];
        // H = B^T G
        double H[3];
        for (int d = 0; d < 3; d++) H[d] = fk.B[0 + d] * G[0] + fk.B[3 + d] * G[1] + fk.B[6 + d] * G[2];
        const double* w = fk.w[i];
        S[0] += w[1] * H[2] - w[2] * H[1];
        S[1] += w[2] * H[0] - w[0] * H[2];
        S[2] += w[0] * H[1] - w[1] * H[0];
        if (i > 0) {
            // joint i-1 rotates frames >= i
            const double* a = fk.ax[i - 1];
            out[3 + (i - 1)] = a[0] * S[0] + a[1] * S[1] + a[2] * S[2];
        }
        for (int d = 0; d < 3; d++) acc[d] += gp[i][d];
    }
    // base: p0 = (x, y, h) + Rz(yaw) t  => d p0 / d yaw = dRz t
    const double dtx = -fk.sy * rp.relative_t[0] - fk.cy * rp.relative_t[1];
    const double dty = fk.cy * rp.relative_t[0] - fk.sy * rp.relative_t[1];
    gyaw += dtx * acc[0] + dty * acc[1];
    out[0] = acc[0];
    out[1] = acc[1];
    out[2] = gyaw;
}
