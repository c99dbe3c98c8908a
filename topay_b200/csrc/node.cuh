// Per-sample arithmetic of the penalty cost and its gradient — one call per sample
// node, no cross-thread communication. The kernels in solver_kernels.cu map lanes
// onto nodes and do the reductions; tests/host_harness.cpp calls the same functions
// on the CPU to check them against the oracle.
//
// Reference: MomaTrajOpt::calSecondStagePenalGrad (src/planner/src/moma_traj_opt.cpp:
// 1200-1829) and calFirstStagePenalGrad (:957-1198). Node index j runs 0..2K inside a
// piece; even j are penalty nodes, odd j are Simpson midpoints.
#pragma once
#include "rog_query.cuh"
#include "hd.cuh"
#include "robot.cuh"

// beta rows at local time s1 (moma_traj_opt.cpp:1263-1270)
TP_HD void tp_basis(double s1, double* b0, double* b1, double* b2, double* b3) {
    const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s3 * s2;
    b0[0] = 1.0; b0[1] = s1; b0[2] = s2; b0[3] = s3; b0[4] = s4; b0[5] = s5;
    b1[0] = 0.0; b1[1] = 1.0; b1[2] = 2.0 * s1; b1[3] = 3.0 * s2; b1[4] = 4.0 * s3; b1[5] = 5.0 * s4;
    b2[0] = 0.0; b2[1] = 0.0; b2[2] = 2.0; b2[3] = 6.0 * s1; b2[4] = 12.0 * s2; b2[5] = 20.0 * s3;
    if (b3) {
        b3[0] = 0.0; b3[1] = 0.0; b3[2] = 0.0; b3[3] = 6.0; b3[4] = 24.0 * s1; b3[5] = 60.0 * s2;
    }
}

// c is the 6 x 9 coefficient block of one piece, row k = coefficient of t^k.
// out[d] = sum_k c[k][d] * b[k] for d in [d0, d1).
template <int D0, int D1>
TP_HD void tp_rows(const double* c, const double* b, double* out) {
#pragma unroll
    for (int d = D0; d < D1; d++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) s += c[k * 9 + d] * b[k];
        out[d] = s;
    }
}

// Yaw / arc-length part of a node (all j): what the Simpson prefix and its adjoint need.
struct TpSlot {
    double cy, sy;          // cos / sin of yaw
    double dth, ds;         // yaw rate, arc-length rate
    double d2th, d2s;
};
TP_HD void tp_slot(const double* c, double t, TpSlot& o, double* b0, double* b1, double* b2) {
    tp_basis(t, b0, b1, b2, nullptr);
    double st[2], d1[2], d2[2];
    tp_rows<0, 2>(c, b0, st);
    tp_rows<0, 2>(c, b1, d1);
    tp_rows<0, 2>(c, b2, d2);
    tp_sincos(st[0], o.sy, o.cy);
    o.dth = d1[0];
    o.ds = d1[1];
    o.d2th = d2[0];
    o.d2s = d2[1];
}

// Chain-rule contraction of one node slot (moma_traj_opt.cpp:1293-1300, 1744-1749,
// 1812-1822): coefX / coefY are the chain weights of this slot already multiplied by
// the fixed Simpson pattern [1,4,2,...,4,1]. Accumulates into the yaw and arc columns of
// the piece's gdC and into its gdT.
TP_HD void tp_chain_slot(const TpSlot& s, const double* b0, const double* b1, double T, int K, int j,
                         double coefX, double coefY, double* g_theta, double* g_arc, double& gdT) {
    const double step = T / K;
    const double coeff = step / 6.0;
    const double alpha = 1.0 / (2 * K) * j;
    const int int_6K = 6 * K;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const double xth = (-s.ds * b0[k] * s.sy) * coeff;
        const double xar = (b1[k] * s.cy) * coeff;
        const double yth = (s.ds * b0[k] * s.cy) * coeff;
        const double yar = (b1[k] * s.sy) * coeff;
        g_arc[k] += xar * coefX + yar * coefY;
        g_theta[k] += xth * coefX + yth * coefY;
    }
    const double xgt = (s.d2s * s.cy - s.ds * s.dth * s.sy) * alpha * coeff + s.ds * s.cy / int_6K;
    const double ygt = (s.d2s * s.sy + s.ds * s.dth * s.cy) * alpha * coeff + s.ds * s.sy / int_6K;
    gdT += xgt * coefX + ygt * coefY;
}

// What one penalty node contributes.
struct TpNodeOut {
    double G0[9], G1[9], G2[9];   // d cost / d (state, dstate, d2state) rows of gradBeta
    double gdT;                   // contribution to gdT of the piece
    double gx, gy;                // xy adjoint added to the chain head (all nodes up to this one)
    double terms[TOPAY_NTERMS];
};

TP_HD void tp_node_clear(TpNodeOut& o) {
#pragma unroll
    for (int d = 0; d < 9; d++) o.G0[d] = o.G1[d] = o.G2[d] = 0.0;
    o.gdT = 0.0;
    o.gx = o.gy = 0.0;
#pragma unroll
    for (int t = 0; t < TOPAY_NTERMS; t++) o.terms[t] = 0.0;
}

// Velocity-polytope ("moment"), linear and angular acceleration limits on the base —
// common to both stages (moma_traj_opt.cpp:1059-1115 and :1334-1462).
TP_HD void tp_base_limits(const TpParams& P, double w_m, double w_a, double w_dw, double omg, double step, int K,
                          double real_alpha, const double* dst, const double* d2st, const double* d3st,
                          TpNodeOut& o) {
    const topay_robot_params& rp = P.robot;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const double sw = half == 0 ? 1.0 : -1.0;   // sign of the max_w term
#pragma unroll
        for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
            const double v = omg_sym * rp.max_v * dst[0] + sw * rp.max_w * dst[1] - rp.max_v * rp.max_w;
            if (v > 0) {
                double f, df;
                tp_smoothL1(P, v, f, df);
                const double gmt = real_alpha * (omg_sym * rp.max_v * d2st[0] + sw * rp.max_w * d2st[1]);
                o.G1[0] += omg * step * w_m * df * omg_sym * rp.max_v;
                o.G1[1] += sw * (omg * step * w_m * df * rp.max_w);
                o.gdT += omg * w_m * (df * gmt * step + f / K);
                o.terms[TOPAY_TERM_MOMENT] += omg * step * w_m * f;
            }
        }
    }
    const double violaAcc = d2st[1] * d2st[1] - rp.max_a * rp.max_a;
    const double violaAlp = d2st[0] * d2st[0] - rp.max_dw * rp.max_dw;
    if (violaAcc > 0) {
        double f, df;
        tp_smoothL1(P, violaAcc, f, df);
        const double gat = 2.0 * real_alpha * d2st[1] * d3st[1];
        o.G2[1] += omg * step * w_a * df * 2.0 * d2st[1];
        o.gdT += omg * w_a * (df * gat * step + f / K);
        o.terms[TOPAY_TERM_ACC] += omg * step * w_a * f;
    }
    if (violaAlp > 0) {
        double f, df;
        tp_smoothL1(P, violaAlp, f, df);
        const double gdot = 2.0 * real_alpha * d2st[0] * d3st[0];
        o.G2[0] += omg * step * w_dw * df * 2.0 * d2st[0];
        o.gdT += omg * w_dw * (df * gdot * step + f / K);
        o.terms[TOPAY_TERM_DOMEGA] += omg * step * w_dw * f;
    }
}

// Stage-1 penalty node (even j), moma_traj_opt.cpp:1016-1135.
TP_HD void tp_node_stage1(const TpParams& P, const double* c, double T, int K, int j, TpNodeOut& o, double* b0,
                          double* b1, double* b2) {
    double b3[6];
    const double step = T / K;
    const double t = j * (step / 2.0);
    tp_basis(t, b0, b1, b2, b3);
    double dst[2], d2st[2], d3st[2];
    tp_rows<0, 2>(c, b1, dst);
    tp_rows<0, 2>(c, b2, d2st);
    tp_rows<0, 2>(c, b3, d3st);
    const double omg = (j == 0 || j == 2 * K) ? 0.5 : 1.0;
    const double real_alpha = 1.0 / K * ((double)j / 2.0);
    tp_node_clear(o);
    tp_base_limits(P, P.opt.s1_moment_weight, P.opt.s1_acc_weight, P.opt.s1_domega_weight, omg, step, K, real_alpha,
                   dst, d2st, d3st, o);
}

// Stage-2 penalty node (even j), moma_traj_opt.cpp:1261-1713. xy is the Simpson
// prefix position CurrentXY at this node. pts / pg are per-thread sphere stores (centres and
// their gradients).
template <class Store>
TP_HD void tp_node_stage2(const TpParams& P, const TpGrid& g, const double* c, double T, int K, int j,
                          const double* xy, TpNodeOut& o, double* b0, double* b1, double* b2, Store& pts,
                          Store& pg) {
    const topay_robot_params& rp = P.robot;
    const topay_opt_params& op = P.opt;
    double b3[6];
    const double step = T / K;
    const double t = j * (step / 2.0);
    tp_basis(t, b0, b1, b2, b3);
    // Only the state and the yaw / arc derivatives are needed before the arm's FK, lookups, pair tests and adjoint;
    // the joint derivatives are evaluated after them (below), so that 42 doubles do not sit in registers across
    // the heaviest part of the node.
    double st[9], dst[2], d2st[2], d3st[2];
    tp_rows<0, 9>(c, b0, st);
    tp_rows<0, 2>(c, b1, dst);
    tp_rows<0, 2>(c, b2, d2st);
    tp_rows<0, 2>(c, b3, d3st);
    const double omg = (j == 0 || j == 2 * K) ? 0.5 : 1.0;
    const double real_alpha = 1.0 / K * ((double)j / 2.0);
    tp_node_clear(o);

    // chassis vs the flat 2-D field (:1304-1332)
    {
        double sdf, gs[2];
        tp_field_query2d_flat(g, xy, sdf, gs);
        const double v = rp.chassis_colli_radius * 1.05 - sdf;
        if (v > 0) {
            double f, df;
            tp_smoothL1(P, v, f, df);
            o.gx += -omg * step * op.s2_collision_weight * df * gs[0];
            o.gy += -omg * step * op.s2_collision_weight * df * gs[1];
            o.gdT += omg * op.s2_collision_weight * (f / K);
            o.terms[TOPAY_TERM_CHASSIS_COLLI] += omg * step * op.s2_collision_weight * f;
        }
    }
    // base limits (:1334-1462)
    tp_base_limits(P, op.s2_moment_weight, op.s2_acc_weight, op.s2_domega_weight, omg, step, K, real_alpha, dst,
                   d2st, d3st, o);

    // arm (:1467-1713)
    double pos[10];
    pos[0] = xy[0];
    pos[1] = xy[1];
    pos[2] = st[0];
#pragma unroll
    for (int q = 0; q < TOPAY_DOF; q++) pos[3 + q] = st[2 + q];
    TpFK fk;
    tp_fk(P, pos, fk, pts);
    const double cost_scale = 10.0;
    const double w_mc = op.s2_mani_colli_weight, w_sc = op.s2_self_colli_weight;
    // sphere vs the 3-D field (:1477-1520) and vs the chassis top (:1525-1564)
    // `active`: some sphere carries a gradient. Spheres are almost always clear of their margins, and with every
    // sphere gradient zero the FK adjoint below returns exact zeros — it is skipped then.
    bool active = false;
    for (int ci = 0; ci < P.n_sphere; ci++) {
        const double pc[3] = {pts.at(ci, 0), pts.at(ci, 1), pts.at(ci, 2)};
        double sdf, gp[3];
        TpTaps3 taps;
        const bool dense = g.kind == 0;
        if (dense) sdf = tp_query3d_taps(g, pc, taps);      // the gradient only when the hinge is active
        else tp_field_query3d(g, pc, sdf, gp);
        const double v = P.sphere_r[ci] * cost_scale * 1.1 - sdf * cost_scale;
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        if (v > 0) {
            if (dense) tp_query3d_grad(g, taps, gp);
            active = true;
            double f, df;
            tp_smoothL1(P, v, f, df);
            const double k = -omg * step * w_mc * df;
            g0 = k * gp[0] * cost_scale;
            g1 = k * gp[1] * cost_scale;
            g2 = k * gp[2] * cost_scale;
            o.gdT += omg * w_mc * (f / K);
            o.terms[TOPAY_TERM_MANI_COLLI] += omg * step * w_mc * f;
        }
        if (ci > 2) {
            const double height = rp.chassis_height + rp.relative_t[2] + P.sphere_r[ci] - pc[2];
            if (height > 0) {
                active = true;
                double f, df;
                tp_smoothL1(P, height, f, df);
                g2 += -omg * step * w_sc * df;
                o.gdT += omg * w_sc * (f / K);
                o.terms[TOPAY_TERM_SELF_COLLI] += omg * step * w_sc * f;
            }
        }
        pg.at(ci, 0) = g0;
        pg.at(ci, 1) = g1;
        pg.at(ci, 2) = g2;
    }
    // link vs link (:1566-1611). Penetrating pairs are rare, so the 66 pair distances are evaluated branch-free
    // first (unrolled by four: the shared-memory loads of a group are issued ahead of its arithmetic), and only the
    // pairs that do penetrate — kept as a bit mask, walked in the reference's order — enter the penalty and its
    // gradient.
    for (int ci = 0; ci < P.n_sphere; ci++) {
        const uint32_t mask = P.pair_mask[ci];
        if (mask == 0u) continue;
        const double pi0 = pts.at(ci, 0), pi1 = pts.at(ci, 1), pi2 = pts.at(ci, 2);
        const double ri = P.sphere_r[ci];
        uint32_t hit = 0u;
#pragma unroll 4
        for (int cj = ci + 1; cj < TOPAY_NSPHERE; cj++) {
            const double dx = pi0 - pts.at(cj, 0), dy = pi1 - pts.at(cj, 1), dz = pi2 - pts.at(cj, 2);
            const double rs = ri + P.sphere_r[cj];
            const double dist = rs * rs - (dx * dx + dy * dy + dz * dz);
            hit |= dist > 0 ? (1u << cj) : 0u;
        }
        hit &= mask;          // pair_mask holds bits c2 > ci, c2 < n_sphere only
        if (hit == 0u) continue;
        active = true;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int cj = ci + 1; cj < P.n_sphere; cj++) {
            if (!((hit >> cj) & 1u)) continue;
            const double dx = pi0 - pts.at(cj, 0), dy = pi1 - pts.at(cj, 1), dz = pi2 - pts.at(cj, 2);
            const double rs = ri + P.sphere_r[cj];
            const double dist = rs * rs - (dx * dx + dy * dy + dz * dz);
            double f, df;
            tp_smoothL1(P, dist, f, df);
            const double k = -omg * step * w_sc * df;
            const double h0 = k * dx * 2.0, h1 = k * dy * 2.0, h2 = k * dz * 2.0;
            o.gdT += omg * w_sc * (f / K);
            o.terms[TOPAY_TERM_SELF_COLLI] += omg * step * w_sc * f;
            a0 += h0;
            a1 += h1;
            a2 += h2;
            pg.at(cj, 0) -= h0;
            pg.at(cj, 1) -= h1;
            pg.at(cj, 2) -= h2;
        }
        pg.at(ci, 0) += a0;
        pg.at(ci, 1) += a1;
        pg.at(ci, 2) += a2;
    }
    double mu[10];
    if (active) {
        tp_fk_adjoint(P, fk, pg, mu);
    } else {
#pragma unroll
        for (int q = 0; q < 10; q++) mu[q] = 0.0;
    }

    // joint position limits (:1616-1666)
    const double w_mp = op.s2_mani_pos_weight;
#pragma unroll
    for (int ji = 0; ji < TOPAY_DOF; ji++) {
        double v = pos[ji + 3] - rp.joint_pos_limit_max[ji];
        if (v > 0) {
            double f, df;
            tp_smoothL1(P, v, f, df);
            mu[ji + 3] += omg * step * w_mp * df;
            o.gdT += omg * w_mp * (f / K);
            o.terms[TOPAY_TERM_MANI_POS] += omg * step * w_mp * f;
        }
        v = -rp.joint_pos_limit_max[ji] - pos[ji + 3];
        if (v > 0) {
            double f, df;
            tp_smoothL1(P, v, f, df);
            mu[ji + 3] -= omg * step * w_mp * df;
            o.gdT += omg * w_mp * (f / K);
            o.terms[TOPAY_TERM_MANI_POS] += omg * step * w_mp * f;
        }
    }
    // unconditional chain add (:1667-1668) and state-level gradients (:1669-1672)
    o.gx += mu[0];
    o.gy += mu[1];
    o.G0[0] += mu[2];
    o.gdT += mu[2] * dst[0] * real_alpha;
    // joint derivatives: the basis is evaluated again from an opaque copy of t (the compiler would otherwise keep
    // the first evaluation alive instead); same operations, same bits
    double t2 = t;
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+d"(t2));
#endif
    tp_basis(t2, b0, b1, b2, b3);
    double dstq[9], d2stq[9], d3stq[9];
    tp_rows<2, 9>(c, b1, dstq);
    tp_rows<2, 9>(c, b2, d2stq);
    tp_rows<2, 9>(c, b3, d3stq);
    double dsum = 0.0;
#pragma unroll
    for (int q = 0; q < TOPAY_DOF; q++) {
        o.G0[2 + q] = mu[3 + q];
        dsum += mu[3 + q] * dstq[2 + q];
    }
    o.gdT += dsum * real_alpha;
    // joint velocity / acceleration limits (:1674-1710)
    const double w_mv = op.s2_mani_vel_weight, w_ma = op.s2_mani_acc_weight;
#pragma unroll
    for (int q = 0; q < TOPAY_DOF; q++) {
        const double dq = dstq[2 + q], d2q = d2stq[2 + q], d3q = d3stq[2 + q];
        const double vdq = dq * dq - rp.joint_vel_limit[q] * rp.joint_vel_limit[q];
        const double vd2q = d2q * d2q - rp.joint_acc_limit[q] * rp.joint_acc_limit[q];
        if (vdq > 0) {
            double f, df;
            tp_smoothL1(P, vdq, f, df);
            const double gvt = 2.0 * real_alpha * dq * d2q;
            o.G1[2 + q] += omg * step * w_mv * df * 2.0 * dq;
            o.gdT += omg * w_mv * (df * gvt * step + f / K);
            o.terms[TOPAY_TERM_MANI_VEL] += omg * step * w_mv * f;
        }
        if (vd2q > 0) {
            double f, df;
            tp_smoothL1(P, vd2q, f, df);
            const double gat = 2.0 * real_alpha * d2q * d3q;
            o.G2[2 + q] += omg * step * w_ma * df * 2.0 * d2q;
            o.gdT += omg * w_ma * (df * gat * step + f / K);
            o.terms[TOPAY_TERM_MANI_ACC] += omg * step * w_ma * f;
        }
    }
}
