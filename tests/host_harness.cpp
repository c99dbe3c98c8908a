// TEST HARNESS ONLY — compiled by tests/test_device_math_host.py into a scratch library.
//
// Instantiates the TP_HD per-thread functions of topay_b200/csrc/*.cuh on the CPU and walks
// them in the same dataflow as the kernels (k_integrate -> k_penalty -> k_chain), serially,
// so that the lane-level maths can be compared with the oracle without a GPU. This is not
// a CPU path of the product: libtopay_b200.so neither contains nor calls any of it.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../topay_b200/csrc/hd.cuh"
#include "../topay_b200/csrc/edt_line.cuh"
#include "../topay_b200/csrc/field_query.cuh"
#include "../topay_b200/csrc/node.cuh"
#include "../topay_b200/csrc/robot.cuh"
#include "../topay_b200/csrc/rog_query.cuh"
#include "../topay_b200/csrc/traj.cuh"

extern "C" {

void hh_make_params(const topay_opt_params* opt, const topay_robot_params* rp, TpParams* out) {
    std::memset(out, 0, sizeof(*out));
    out->robot = *rp;
    out->opt = *opt;
    tp_derive_params(*out);
}
int hh_params_size() { return (int)sizeof(TpParams); }

void hh_make_grid(const topay_grid_desc* d, const double* e3, const double* e2, const double* e2i, const double* e2c,
                  TpGrid* g) {
    g->resolution = d->resolution;
    g->resolution_inv = 1.0 / d->resolution;
    for (int i = 0; i < 3; i++) {
        g->min_boundary[i] = -d->map_size[i] / 2.0;
        g->max_boundary[i] = d->map_size[i] / 2.0;
    }
    g->min_boundary[2] = 0.0;
    g->max_boundary[2] = d->map_size[2];
    for (int i = 0; i < 3; i++) {
        g->origin[i] = g->min_boundary[i];
        g->dims[i] = (int)ceil(d->map_size[i] / d->resolution);
    }
    g->ready = 1;
    g->esdf3d = e3;
    g->esdf2d = e2;
    g->esdf2d_inflate = e2i;
    g->esdf2d_critical = e2c;
}
int hh_grid_size() { return (int)sizeof(TpGrid); }

void hh_query3d(const TpGrid* g, const double* pos, int64_t n, double* dist, double* grad) {
    for (int64_t i = 0; i < n; i++) tp_query3d(*g, pos + 3 * i, dist[i], grad + 3 * i);
}
void hh_query2d(const TpGrid* g, int which, const double* pos, int64_t n, double* dist, double* grad) {
    const double* buf = which == 2 ? g->esdf2d_critical : which == 1 ? g->esdf2d_inflate : g->esdf2d;
    for (int64_t i = 0; i < n; i++) tp_query2d(*g, buf, pos + 2 * i, dist[i], grad + 2 * i);
}

void hh_line_visib(const TpGrid* g, const double* p1, const double* p2, int64_t n, double thresh, int critical,
                   int8_t* visible, double* pc) {
    for (int64_t i = 0; i < n; i++) visible[i] = tp_line_visib(*g, p1 + 3 * i, p2 + 3 * i, thresh, critical != 0, pc + 3 * i) ? 1 : 0;
}

void hh_fk(const TpParams* P, const double* pos10, double* pts36) {
    TpFK fk;
    TpSphereStoreLocal pts;
    tp_fk(*P, pos10, fk, pts);
    std::memcpy(pts36, pts.a, sizeof(pts.a));
}
void hh_fk_adjoint(const TpParams* P, const double* pos10, const double* g36, double* out10) {
    TpFK fk;
    TpSphereStoreLocal pts, g;
    tp_fk(*P, pos10, fk, pts);
    std::memcpy(g.a, g36, sizeof(g.a));
    tp_fk_adjoint(*P, fk, g, out10);
}

// Penalty part of one evaluation of one candidate, walking the kernels' dataflow serially.
// coeff: 6N x 9, T: N. Outputs: gdC (6N x 9), gdT (N) [both without the jerk part, mean-time
// included in gdT], terms[13] (penalty terms only), final_xy[2].
void hh_penalty_eval(const TpParams* Pp, const TpGrid* Gp, int stage, int N, const double* coeff, const double* T,
                     const double* start_xy, const double* end_xy, const double* init_inner_xy, const double* lambda,
                     const double* rho, double* gdC, double* gdT, double* terms, double* final_xy) {
    const TpParams& P = *Pp;
    const TpGrid& G = *Gp;
    const int K = P.opt.int_K, L = 2 * K + 1;
    std::vector<double> Ix((size_t)N * K), Iy((size_t)N * K), totx(N), toty(N);
    double b0[6], b1[6], b2[6];
    // ---- k_integrate
    for (int i = 0; i < N; i++) {
        const double* c = coeff + (size_t)6 * i * 9;
        const double step = T[i] / K, half_step = step / 2.0, cf = step / 6.0;
        std::vector<TpSlot> sl(L);
        for (int j = 0; j < L; j++) tp_slot(c, j * half_step, sl[j], b0, b1, b2);
        double sx = 0, sy = 0;
        for (int k = 0; k < K; k++) {
            const int j = 2 * k;
            double ix = cf * sl[j].ds * sl[j].cy, iy = cf * sl[j].ds * sl[j].sy;
            ix += 4 * cf * sl[j + 1].ds * sl[j + 1].cy;
            iy += 4 * cf * sl[j + 1].ds * sl[j + 1].sy;
            ix += cf * sl[j + 2].ds * sl[j + 2].cy;
            iy += cf * sl[j + 2].ds * sl[j + 2].sy;
            Ix[(size_t)i * K + k] = ix;
            Iy[(size_t)i * K + k] = iy;
            sx += ix;
            sy += iy;
        }
        totx[i] = sx;
        toty[i] = sy;
    }
    // ---- k_penalty
    std::memset(gdC, 0, sizeof(double) * 6 * N * 9);
    std::memset(gdT, 0, sizeof(double) * N);
    for (int t = 0; t < TOPAY_NTERMS; t++) terms[t] = 0.0;
    std::vector<double> gx((size_t)N * (K + 1), 0.0), gy((size_t)N * (K + 1), 0.0);
    for (int i = 0; i < N; i++) {
        const double* c = coeff + (size_t)6 * i * 9;
        double px = start_xy[0], py = start_xy[1];
        for (int p = 0; p < i; p++) {
            px += totx[p];
            py += toty[p];
        }
        for (int jn = 0; jn <= K; jn++) {
            double xy[2] = {px, py};
            for (int k = 0; k < jn; k++) {
                xy[0] += Ix[(size_t)i * K + k];
                xy[1] += Iy[(size_t)i * K + k];
            }
            TpNodeOut o;
            if (stage == 1)
                tp_node_stage1(P, c, T[i], K, 2 * jn, o, b0, b1, b2);
            else {
                TpSphereStoreLocal pts, pg;
                tp_node_stage2(P, G, c, T[i], K, 2 * jn, xy, o, b0, b1, b2, pts, pg);
            }
            for (int k = 0; k < 6; k++)
                for (int d = 0; d < 9; d++)
                    gdC[((size_t)6 * i + k) * 9 + d] += b0[k] * o.G0[d] + b1[k] * o.G1[d] + b2[k] * o.G2[d];
            gdT[i] += o.gdT;
            for (int t = 0; t < TOPAY_NTERMS; t++) terms[t] += o.terms[t];
            gx[(size_t)i * (K + 1) + jn] = o.gx;
            gy[(size_t)i * (K + 1) + jn] = o.gy;
        }
    }
    // ---- k_chain
    std::vector<double> Fx(N + 1), Fy(N + 1);
    Fx[0] = start_xy[0];
    Fy[0] = start_xy[1];
    for (int i = 0; i < N; i++) {
        Fx[i + 1] = Fx[i] + totx[i];
        Fy[i + 1] = Fy[i] + toty[i];
    }
    final_xy[0] = Fx[N] - end_xy[0];
    final_xy[1] = Fy[N] - end_xy[1];
    for (int i = 0; i < N; i++) {
        double wx = 0, wy = 0;
        if (stage == 1) {
            for (int p = i + 1; p < N; p++) {
                wx += P.opt.s1_path_pos_weight * 2.0 * (Fx[p + 1] - init_inner_xy[2 * p]);
                wy += P.opt.s1_path_pos_weight * 2.0 * (Fy[p + 1] - init_inner_xy[2 * p + 1]);
            }
        } else {
            wx = rho[0] * (final_xy[0] + lambda[0] / rho[0]);
            wy = rho[1] * (final_xy[1] + lambda[1] / rho[1]);
            for (int p = i + 1; p < N; p++)
                for (int jn = 0; jn <= K; jn++) {
                    wx += gx[(size_t)p * (K + 1) + jn];
                    wy += gy[(size_t)p * (K + 1) + jn];
                }
        }
        const double* c = coeff + (size_t)6 * i * 9;
        const double half_step = (T[i] / K) / 2.0;
        double gth[6] = {0}, gar[6] = {0}, gdt = 0.0;
        for (int j = 0; j < L; j++) {
            double sx = 0, sy = 0;   // adjoints of this piece's nodes at or after slot j
            for (int jn = 0; jn <= K; jn++)
                if (2 * jn >= j) {
                    sx += gx[(size_t)i * (K + 1) + jn];
                    sy += gy[(size_t)i * (K + 1) + jn];
                }
            const double icc = (j == 0 || j == 2 * K) ? 1.0 : (j % 2 ? 4.0 : 2.0);
            TpSlot sl;
            tp_slot(c, j * half_step, sl, b0, b1, b2);
            tp_chain_slot(sl, b0, b1, T[i], K, j, (wx + sx) * icc, (wy + sy) * icc, gth, gar, gdt);
        }
        for (int k = 0; k < 6; k++) {
            gdC[((size_t)6 * i + k) * 9 + 0] += gth[k];
            gdC[((size_t)6 * i + k) * 9 + 1] += gar[k];
        }
        gdT[i] += gdt;
    }
    // ---- k_cand's share of the penalty: end-point / path terms and the mean-time penalty
    if (stage == 1) {
        double cp = 0.0;
        for (int i = 0; i < N; i++) {
            const double ex = Fx[i + 1] - init_inner_xy[2 * i], ey = Fy[i + 1] - init_inner_xy[2 * i + 1];
            cp += P.opt.s1_path_pos_weight * (ex * ex + ey * ey);
        }
        terms[TOPAY_TERM_ENDP] = cp;
    } else {
        double avg = 0.0;
        for (int i = 0; i < N; i++) avg += T[i];
        avg /= N;
        const double w = P.opt.s2_mean_time_weight;
        double all = 0.0;
        for (int i = 0; i < N; i++) {
            if (T[i] < avg * 0.5) {
                terms[TOPAY_TERM_MEAN_TIME] += w * (T[i] - avg * 0.5) * (T[i] - avg * 0.5);
                all += w * 2.0 * (T[i] - avg * 0.5) * (-0.5 / N);
                gdT[i] += w * 2.0 * (T[i] - avg * 0.5);
            }
            if (T[i] > avg * 2.0) {
                terms[TOPAY_TERM_MEAN_TIME] += w * (T[i] - avg * 2.0) * (T[i] - avg * 2.0);
                all += w * 2.0 * (T[i] - avg * 2.0) * (-2.0 / N);
                gdT[i] += w * 2.0 * (T[i] - avg * 2.0);
            }
        }
        for (int i = 0; i < N; i++) gdT[i] += all;
        terms[TOPAY_TERM_ENDP] = 0.5 * (rho[0] * (final_xy[0] + lambda[0] / rho[0]) * (final_xy[0] + lambda[0] / rho[0]) +
                                        rho[1] * (final_xy[1] + lambda[1] / rho[1]) * (final_xy[1] + lambda[1] / rho[1]));
    }
}

// ---- ROG-Map ring lookups (rog_query.cuh)
void hh_make_rog_grid(double res, const int* half, const int* size, const double* dist3, const double* crit,
                      const double* flat, TpGrid* g) {
    std::memset(g, 0, sizeof(*g));
    g->kind = 1;
    g->rog.res = res;
    g->rog.res_inv = 1.0 / res;
    for (int i = 0; i < 3; i++) {
        g->rog.half[i] = half[i];
        g->rog.size[i] = size[i];
    }
    g->rog.dist3 = dist3;
    g->rog.crit = crit;
    g->rog.flat = flat;
    g->resolution = res;
    g->resolution_inv = 1.0 / res;
    g->ready = 1;
}
void hh_rog_query(const TpGrid* g, int kind, const double* pos, int64_t n, double* dist, double* grad) {
    const TpRog& r = g->rog;
    for (int64_t i = 0; i < n; i++) {
        const double* p = pos + 3 * i;
        double gg[3] = {0, 0, 0};
        switch (kind) {
            case TOPAY_ROG_Q_EDT: tp_rog_value_grad(r, p, dist[i], gg); break;
            case TOPAY_ROG_Q_FLAT: tp_rog_value_grad2d(r, r.flat, p, dist[i], gg); break;
            case TOPAY_ROG_Q_CRITICAL: tp_rog_value_grad2d(r, r.crit, p, dist[i], gg); break;
            case TOPAY_ROG_Q_CELL: dist[i] = tp_rog_cell3(r, p); break;
            case TOPAY_ROG_Q_CELL_FLAT: dist[i] = tp_rog_cell2(r, r.flat, p); break;
            default: dist[i] = tp_rog_cell2(r, r.crit, p); break;
        }
        for (int k = 0; k < 3; k++) grad[3 * i + k] = gg[k];
    }
}
void hh_rog_line_free(const TpGrid* g, const double* s, const double* e, int64_t n, double thr, int8_t* out) {
    for (int64_t i = 0; i < n; i++) {
        const long cap = labs((long)floor(e[2 * i] / g->rog.res) - (long)floor(s[2 * i] / g->rog.res)) +
                         labs((long)floor(e[2 * i + 1] / g->rog.res) - (long)floor(s[2 * i + 1] / g->rog.res)) + 1;
        out[i] = tp_rog_line_free2d(g->rog, s + 2 * i, e + 2 * i, thr, (int)cap) ? 1 : 0;
    }
}
// the field-kind dispatch the solver and the gate use
void hh_field_dispatch(const TpGrid* g, const double* pos, int64_t n, double* d2, double* g2, double* d3, double* g3,
                       double* v2, double* v3) {
    for (int64_t i = 0; i < n; i++) {
        tp_field_query2d_flat(*g, pos + 3 * i, d2[i], g2 + 2 * i);
        tp_field_query3d(*g, pos + 3 * i, d3[i], g3 + 3 * i);
        v2[i] = tp_field_distance2d(*g, pos + 3 * i);
        v3[i] = tp_field_distance3d(*g, pos + 3 * i);
    }
}
// ---- dense-field predicates and coarse lookups (field_query.cuh)
void hh_field_misc(const TpGrid* g, const double* p2, const double* q2, const double* p3, const int32_t* idx, int64_t n,
                   double thr, int critical, int8_t* c2, int8_t* c3, int8_t* line, double* coarse_d, double* coarse_i) {
    for (int64_t i = 0; i < n; i++) {
        c2[i] = tp_is_collision2d(*g, p2 + 2 * i, thr);
        c3[i] = tp_is_collision3d(*g, p3 + 3 * i, thr);
        line[i] = tp_line_collision_grid2d(*g, p2 + 2 * i, q2 + 2 * i, thr);
        int id[2];
        tp_pos_to_index2(*g, p2 + 2 * i, id);
        coarse_d[i] = tp_dist_coarse2i(*g, id[0], id[1], critical != 0);
        coarse_i[i] = tp_dist_coarse2i(*g, idx[2 * i], idx[2 * i + 1], critical != 0);
    }
}
// ---- trajectory evaluation (traj.cuh): states at m times given the pose table
void hh_traj_sample(int N, const double* T, const double* coeff, const double* car_seq, int n_seq, const double* t, int m,
                    double* state, double* dstate, double* pva) {
    TpPoly p{N, T, coeff};
    const double total = tp_poly_total(p);
    for (int j = 0; j < m; j++) {
        tp_traj_state(p, car_seq, n_seq, total, t[j], state + 10 * j);
        tp_traj_dstate(p, total, t[j], dstate + 10 * j);
        tp_poly_pos(p, t[j], 0, 9, pva + 27 * j);
        tp_poly_vel(p, t[j], 0, 9, pva + 27 * j + 9);
        tp_poly_acc(p, t[j], 0, 9, pva + 27 * j + 18);
    }
}
// ---- one line of the strided EDT passes (edt_line.cuh) in k_edt_line's schedule, serially: skip distances of the
// "no source" runs, chunk boundaries against the whole line, chunk insides by halving. Sign-packed line in, positive
// and negative transform out.
void hh_edt_line(const int32_t* f_in, int n, int CH, int32_t* out_pos, int32_t* out_neg) {
    std::vector<int> f(f_in, f_in + n), arg(n);
    {
        int nextc = n, lastc = -1;
        for (int v = n - 1; v >= 0; v--) {
            if (f[v] < TP_INF32) nextc = v;
            else f[v] = tp_skip_make(0, nextc - v);
        }
        for (int v = 0; v < n; v++) {
            if (f[v] < TP_INF32) lastc = v;
            else f[v] |= (v - lastc) << TP_SKIP_BITS;
        }
    }
    auto emit = [&](int q, int bp) {
        out_pos[q] = bp >= TP_INF32 ? TP_INF32 : bp;
        out_neg[q] = tp_neg_search<1>(f.data(), n, q);
    };
    const int nb = tp_edt_boundaries(n, CH);
    for (int b = 0; b < nb; b++) {
        const int q = tp_edt_boundary(b, n, CH);
        int a;
        const int bp = tp_dc_query<1>(f.data(), q, 0, n - 1, a);
        arg[q] = a;
        emit(q, bp);
    }
    for (int ch = 0; ch < nb - 1; ch++) {
        const int lo = tp_edt_boundary(ch, n, CH), hi = tp_edt_boundary(ch + 1, n, CH);
        for (int h = tp_dc_top(hi - lo) >> 1; h >= 1; h >>= 1)
            for (int q = lo + h; q < hi; q += 2 * h) {
                const int qr = q + h < hi ? q + h : hi;
                int a;
                const int bp = tp_dc_query<1>(f.data(), q, arg[q - h], arg[qr], a);
                arg[q] = a;
                emit(q, bp);
            }
    }
}
int hh_sq16(int v) { return tp_sq16(v); }
// ---- the same line in k_edt_scan's schedule: the line cut into NS segments of L cells, one envelope per segment
// (its sources against the whole line), the positive transform of a free cell = minimum over the envelopes; an
// occupied cell takes the outward search of the negative transform.
void hh_edt_scan_line(const int32_t* f, int n, int NS, int32_t* out_pos, int32_t* out_neg, int32_t* depth) {
    const int L = (((n + NS - 1) / NS) + 31) & ~31;
    std::vector<TpEnvEntry> stk(L);
    int deepest = 0;
    for (int u = 0; u < n; u++) {
        out_pos[u] = f[u] > 0 ? TP_INF32 : 0;
        out_neg[u] = 0;
    }
    for (int j = 0; j < NS; j++) {
        const int s0 = j * L, s1 = std::min(s0 + L, n);
        if (s0 >= s1) continue;
        TpEnv e{-1, 0, 0, 0};
        int prev = s0 > 0 ? tp_env_val<false>(f[s0 - 1]) : 1, cur = tp_env_val<false>(f[s0]);
        for (int u = s0; u < s1; u++) {
            const int nxt = u + 1 < n ? tp_env_val<false>(f[u + 1]) : 1;
            tp_env_push(e, stk.data(), n, u, cur, prev, nxt);
            deepest = std::max(deepest, e.q + 1);
            prev = cur;
            cur = nxt;
        }
        for (int u = n - 1; u >= 0; u--) {
            const int v = tp_env_pop(e, stk.data(), u);
            if (f[u] > 0) out_pos[u] = std::min(out_pos[u], v);
        }
    }
    for (int u = 0; u < n; u++) {
        if (f[u] >= 0) continue;
        int bn = tp_env_val<true>(f[u]);
        for (int d = 1; d < n; d++) {
            const int dd = d * d;
            if (dd >= bn) break;
            if (u - d >= 0) bn = std::min(bn, dd + tp_env_val<true>(f[u - d]));
            if (u + d < n) bn = std::min(bn, dd + tp_env_val<true>(f[u + d]));
        }
        out_neg[u] = bn;
    }
    *depth = deepest;
}
}  // extern "C"
