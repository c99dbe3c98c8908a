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
//
// Register discipline: every loop has a compile-time trip count and every private array
// is indexed statically after unrolling; the only dynamically indexed storage is the
// sphere store (shared memory in the kernels). The backward sweep does not keep the
// frames of the forward pass: it un-rotates W step by step with the saved sin/cos.
#pragma once
#include "hd.cuh"

// sin and cos of the same angle with one range reduction (device: sincos(), same results as sin() and cos())
TP_HD void tp_sincos(double a, double& s, double& c) {
#if defined(__CUDA_ARCH__)
    sincos(a, &s, &c);
#else
    s = sin(a);
    c = cos(a);
#endif
}

// Per-thread storage of 12 x 3 doubles with a runtime sphere index. Kernels place it in
// shared memory (element stride = block size, so lanes never conflict); the CPU harness
// uses a plain array.
struct TpSphereStoreLocal {
    double a[TOPAY_NSPHERE * 3];
    TP_HD double& at(int n, int d) { return a[n * 3 + d]; }
};
struct TpSphereStoreStrided {
    double* base;
    int stride;
    TP_HD double& at(int n, int d) { return base[(n * 3 + d) * stride]; }
};

struct TpFK {
    double B[9];                 // Rz(yaw) * relative_R, row-major
    double sq[TOPAY_DOF], cq[TOPAY_DOF];
    double c0[3], c1[3], c2[3];  // columns of W_7 after the forward pass
    double sy, cy;               // sin / cos of yaw
};

// pos = (x, y, yaw, q1..q7). Writes the sphere centres to pts.
template <class Store>
TP_HD void tp_fk(const TpParams& P, const double* pos, TpFK& fk, Store& pts) {
    const topay_robot_params& rp = P.robot;
    double s, c;
    tp_sincos(pos[2], s, c);
    fk.sy = s;
    fk.cy = c;
    const double* R = rp.relative_R;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        fk.B[0 + j] = c * R[0 + j] + (-s) * R[3 + j];
        fk.B[3 + j] = s * R[0 + j] + c * R[3 + j];
        fk.B[6 + j] = R[6 + j];
    }
    double pcur[3];
    pcur[0] = pos[0] + (c * rp.relative_t[0] + (-s) * rp.relative_t[1]);
    pcur[1] = pos[1] + (s * rp.relative_t[0] + c * rp.relative_t[1]);
    pcur[2] = rp.chassis_height + rp.relative_t[2];
    double c0[3] = {1, 0, 0}, c1[3] = {0, 1, 0}, c2[3] = {0, 0, 1};
#pragma unroll
    for (int i = 0; i < TOPAY_DOF + 1; i++) {
        double z[3];
#pragma unroll
        for (int d = 0; d < 3; d++) z[d] = fk.B[3 * d + 0] * c2[0] + fk.B[3 * d + 1] * c2[1] + fk.B[3 * d + 2] * c2[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int n = P.slot_sphere[2 * i + j];
            if (n >= 0) {
                const double off = rp.colli_points[2 * i + j];
#pragma unroll
                for (int d = 0; d < 3; d++) pts.at(n, d) = pcur[d] + z[d] * off;
            }
        }
        if (i < TOPAY_DOF) {
#pragma unroll
            for (int d = 0; d < 3; d++) pcur[d] += z[d] * rp.colli_length[i];
            double sq, cq;
            tp_sincos(pos[3 + i], sq, cq);
            fk.sq[i] = sq;
            fk.cq[i] = cq;
            if (i % 2 == 0) {  // Rz(q)
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double a = c0[d], b = c1[d];
                    c0[d] = a * cq + b * sq;
                    c1[d] = a * (-sq) + b * cq;
                }
            } else {  // Ry(q)
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double a = c0[d], b = c2[d];
                    c0[d] = a * cq + b * (-sq);
                    c2[d] = a * sq + b * cq;
                }
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        fk.c0[d] = c0[d];
        fk.c1[d] = c1[d];
        fk.c2[d] = c2[d];
    }
}

// g.at(n, :) = d cost / d sphere centre n. out = d cost / d (x, y, yaw, q1..q7).
template <class Store>
TP_HD void tp_fk_adjoint(const TpParams& P, const TpFK& fk, Store& g, double* out) {
    const topay_robot_params& rp = P.robot;
    double c0[3], c1[3], c2[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        c0[d] = fk.c0[d];
        c1[d] = fk.c1[d];
        c2[d] = fk.c2[d];
    }
    double acc[3] = {0, 0, 0};   // sum of position gradients of frames > i
    double S[3] = {0, 0, 0};     // suffix sum of w_i x (B^T G_i)
    double gyaw = 0.0;
#pragma unroll
    for (int i = TOPAY_DOF; i >= 0; i--) {
        // gradients of the spheres attached to frame i
        double gp[3] = {0, 0, 0}, gz[3] = {0, 0, 0};
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int n = P.slot_sphere[2 * i + j];
            if (n >= 0) {
                const double off = rp.colli_points[2 * i + j];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double v = g.at(n, d);
                    gp[d] += v;
                    gz[d] += off * v;
                }
            }
        }
        double G[3];
#pragma unroll
        for (int d = 0; d < 3; d++) G[d] = gz[d] + (i < TOPAY_DOF ? rp.colli_length[i < TOPAY_DOF ? i : 0] * acc[d] : 0.0);
        double z[3];
#pragma unroll
        for (int d = 0; d < 3; d++) z[d] = fk.B[3 * d + 0] * c2[0] + fk.B[3 * d + 1] * c2[1] + fk.B[3 * d + 2] * c2[2];
        gyaw += z[0] * G[1] - z[1] * G[0];
        double H[3];
#pragma unroll
        for (int d = 0; d < 3; d++) H[d] = fk.B[0 + d] * G[0] + fk.B[3 + d] * G[1] + fk.B[6 + d] * G[2];
        S[0] += c2[1] * H[2] - c2[2] * H[1];
        S[1] += c2[2] * H[0] - c2[0] * H[2];
        S[2] += c2[0] * H[1] - c2[1] * H[0];
        if (i > 0) {
            // joint i-1 (rotation R_i) moves frames >= i; its axis is invariant under R_i
            const int jt = i - 1;
            const double sq = fk.sq[jt], cq = fk.cq[jt];
            if (jt % 2 == 0) {
                out[3 + jt] = c2[0] * S[0] + c2[1] * S[1] + c2[2] * S[2];
#pragma unroll
                for (int d = 0; d < 3; d++) {   // W_{i-1} = W_i Rz(q)^T
                    const double a = c0[d], b = c1[d];
                    c0[d] = a * cq - b * sq;
                    c1[d] = a * sq + b * cq;
                }
            } else {
                out[3 + jt] = c1[0] * S[0] + c1[1] * S[1] + c1[2] * S[2];
#pragma unroll
                for (int d = 0; d < 3; d++) {   // W_{i-1} = W_i Ry(q)^T
                    const double a = c0[d], b = c2[d];
                    c0[d] = a * cq + b * sq;
                    c2[d] = -a * sq + b * cq;
                }
            }
        }
#pragma unroll
        for (int d = 0; d < 3; d++) acc[d] += gp[d];
    }
    const double dtx = -fk.sy * rp.relative_t[0] - fk.cy * rp.relative_t[1];
    const double dty = fk.cy * rp.relative_t[0] - fk.sy * rp.relative_t[1];
    gyaw += dtx * acc[0] + dty * acc[1];
    out[0] = acc[0];
    out[1] = acc[1];
    out[2] = gyaw;
}
