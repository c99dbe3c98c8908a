// Shared definitions for the device code of libtopay_b200.
//
// The per-thread arithmetic lives in headers whose functions are declared TP_HD
// (__host__ __device__ under nvcc). The product library only ever calls them from
// kernels; tests/host_harness.cpp instantiates the same functions on the CPU so
// the lane-level maths can be unit-tested without a GPU. There is no CPU path in
// the product.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/topay_b200.h"

#if defined(__CUDACC__)
#define TP_HD __host__ __device__ __forceinline__
#else
#define TP_HD inline
#endif

#define TP_MAX_K 32          // int_K supported by the warp-per-piece kernels
#define TP_LBFGS_MAX_PAST 16 // pf ring (reference uses past <= 8)

// Everything the kernels need that is constant for the life of a solver/field:
// robot constants (moma_param.h:36-143), optimiser weights (optimizer.yaml via
// moma_traj_opt.h:845-941) and the dense grid geometry (grid_map.cpp:33-54).
// View of the ROG-Map ring buffers (rog_map::ESDFMap, src/rog_map/include/rog_map/esdf_map.h:119-123);
// the lookups are in rog_query.cuh.
struct TpRog {
    double res, res_inv;
    int half[3], size[3];
    const double* dist3;    // distance_buffer,       x*Sy*Sz + y*Sz + z
    const double* crit;     // distance_buffer_2d,    x*Sy + y
    const double* flat;     // distance_buffer_flat,  x*Sy + y
};

struct TpGrid {
    int32_t kind;           // 0: dense GridMap buffers below; 1: `rog` (GridMap's use_rog branches)
    int32_t pad_;
    TpRog rog;
    double resolution, resolution_inv;
    double origin[3];
    double min_boundary[3], max_boundary[3];
    int32_t dims[3];
    int32_t ready;
    const double* esdf3d;
    const double* esdf2d;          // flat (z < chassis height)
    const double* esdf2d_inflate;
    const double* esdf2d_critical;
};

struct TpParams {
    topay_robot_params robot;
    topay_opt_params opt;
    // derived
    double smooth_f3c, smooth_f4c, smooth_d2c, smooth_d3c, smooth_half;
    int32_t sphere_frame[TOPAY_NSPHERE];   // frame index of each sphere (colli_link_map)
    double  sphere_off[TOPAY_NSPHERE];     // offset along the frame's z axis
    double  sphere_r[TOPAY_NSPHERE];
    int32_t n_sphere;
    int32_t slot_sphere[2 * (TOPAY_DOF + 1)];  // sphere index of (frame i, slot j) or -1
    uint32_t pair_mask[TOPAY_NSPHERE];     // bit c2 set => pair (c, c2), c2 > c, is checked
};

TP_HD void tp_derive_params(TpParams& p) {
    const double pe = p.opt.relu_mu;
    p.smooth_half = 0.5 * pe;
    p.smooth_f3c = 1.0 / (pe * pe);
    p.smooth_f4c = -0.5 * p.smooth_f3c / pe;
    p.smooth_d2c = 3.0 * p.smooth_f3c;
    p.smooth_d3c = 4.0 * p.smooth_f4c;
    int n = 0;
    for (int i = 0; i < TOPAY_DOF + 1; i++)
        for (int j = 0; j < 2; j++) {
            p.slot_sphere[i * 2 + j] = -1;
            if (p.robot.colli_points[i * 2 + j] == 0.0) continue;
            if (n < TOPAY_NSPHERE) {
                p.slot_sphere[i * 2 + j] = n;
                p.sphere_frame[n] = i;
                p.sphere_off[n] = p.robot.colli_points[i * 2 + j];
                p.sphere_r[n] = p.robot.colli_point_radius[i * 2 + j];
            }
            n++;
        }
    p.n_sphere = n < TOPAY_NSPHERE ? n : TOPAY_NSPHERE;
    for (int c = 0; c < TOPAY_NSPHERE; c++) {
        uint32_t m = 0;
        for (int c2 = c + 1; c2 < p.n_sphere; c2++)
            if (p.robot.collision_matrix[c * TOPAY_NSPHERE + c2] == -1) m |= (1u << c2);
        p.pair_mask[c] = c < p.n_sphere ? m : 0u;
    }
}

// ---- scalar maps, moma_traj_opt.h:744-830 ----
TP_HD double tp_expC2(double tau) {
    return tau > 0.0 ? ((0.5 * tau + 1.0) * tau + 1.0) : 1.0 / ((0.5 * tau - 1.0) * tau + 1.0);
}
TP_HD double tp_logC2(double T) {
    return T > 1.0 ? (sqrt(2.0 * T - 1.0) - 1.0) : (1.0 - sqrt(2.0 / T - 1.0));
}
TP_HD double tp_dT_dtau(double tau) {
    if (tau > 0) return tau + 1.0;
    const double den = (0.5 * tau - 1.0) * tau + 1.0;
    return (1.0 - tau) / (den * den);
}
TP_HD double tp_sigmoidC2(double vq, double max_q) {
    const double e = tp_expC2(vq);
    return 2.0 * max_q * e / (1.0 + e) - max_q;
}
TP_HD double tp_dq_dvq(double vq, double max_q) {
    const double e1 = tp_expC2(vq) + 1.0;
    return 2.0 * max_q * tp_dT_dtau(vq) / (e1 * e1);
}
// smoothL1Penalty, moma_traj_opt.h:810-830 (only ever called with x > 0)
TP_HD void tp_smoothL1(const TpParams& p, double x, double& f, double& df) {
    if (x < p.opt.relu_mu) {
        f = (p.smooth_f4c * x + p.smooth_f3c) * x * x * x;
        df = (p.smooth_d3c * x + p.smooth_d2c) * x * x;
    } else {
        f = x - p.smooth_half;
        df = 1.0;
    }
}
