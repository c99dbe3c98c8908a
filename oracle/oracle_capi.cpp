// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Flat C entry points over the CPU restatement, bound with ctypes from tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// PARITY UNPINNED — see oracle_field.hpp.
#include <atomic>
#include <cstring>
#include <thread>

#include "oracle_field.hpp"
#include "oracle_front.hpp"
#include "oracle_robot.hpp"
#include "oracle_rog.hpp"
#include "oracle_solve.hpp"
#include "oracle_traj.hpp"

using namespace oracle;

static bool g_exact_chain = false;  // debug: see TrajOpt::exact_chain
static const RogEsdf* g_rog = nullptr;   // when set, the solve / gate read the ROG ring instead of `field`

extern "C" {

void oracle_set_exact_chain(int on) { g_exact_chain = on != 0; }
void oracle_use_rog(void* rog_handle) { g_rog = (const RogEsdf*)rog_handle; }

void oracle_robot_params_default(topay_robot_params* out) { robot_defaults(out); }
void oracle_opt_params_default(topay_opt_params* out) { opt_defaults(out); }

// ------------------------------------------------------------------ field
void* oracle_field_create(const topay_grid_desc* d) {
    Field* f = new Field();
    f->init(*d);
    return f;
}
void oracle_field_destroy(void* h) { delete (Field*)h; }
void oracle_field_dims(void* h, int32_t* dims) {
    Field* f = (Field*)h;
    for (int i = 0; i < 3; i++) dims[i] = f->voxel_num[i];
}
void oracle_field_set_occupancy(void* h, const int8_t* occ3d, const int8_t* occ2d, const int8_t* occ2dc) {
    Field* f = (Field*)h;
    if (occ3d) f->occ_buffer_3d.assign((const char*)occ3d, (const char*)occ3d + f->buffer_size_3d);
    if (occ2d) f->occ_buffer_2d.assign((const char*)occ2d, (const char*)occ2d + f->buffer_size_2d);
    if (occ2dc) f->occ_buffer_2d_critical.assign((const char*)occ2dc, (const char*)occ2dc + f->buffer_size_2d);
}
void oracle_field_clear(void* h, int clear_critical) { ((Field*)h)->clear(clear_critical != 0); }
void oracle_field_rasterize(void* h, const float* xyz, int64_t n) { ((Field*)h)->rasterize(xyz, n); }
void oracle_field_rebuild(void* h) { ((Field*)h)->update_esdf(); }
void oracle_field_query3d(void* h, const double* pos, int64_t n, double* dist, double* grad) {
    Field* f = (Field*)h;
    for (int64_t i = 0; i < n; i++) {
        double g[3];
        f->dis_with_grad_3d(pos + 3 * i, dist[i], g);
        if (grad) std::memcpy(grad + 3 * i, g, sizeof(g));
    }
}
void oracle_field_query2d(void* h, const double* pos, int64_t n, int which, double* dist, double* grad) {
    Field* f = (Field*)h;
    for (int64_t i = 0; i < n; i++) {
        double g[2];
        f->dis_with_grad_2d(pos + 2 * i, dist[i], g, which == TOPAY_MAP2D_INFLATE, which == TOPAY_MAP2D_CRITICAL);
        if (grad) std::memcpy(grad + 2 * i, g, sizeof(g));
    }
}
void oracle_field_distance3d(void* h, const double* pos, int64_t n, double* dist) {
    Field* f = (Field*)h;
    for (int64_t i = 0; i < n; i++) f->distance3d(pos + 3 * i, dist[i]);
}
void oracle_field_distance2d(void* h, const double* pos, int64_t n, double* dist) {
    Field* f = (Field*)h;
    for (int64_t i = 0; i < n; i++) f->distance2d(pos + 2 * i, dist[i]);
}
void oracle_field_whole_body_collision(void* h, const topay_robot_params* rp, const double* states, int64_t n,
                                       int8_t* out) {
    Field* f = (Field*)h;
    for (int64_t i = 0; i < n; i++) out[i] = whole_body_collision(*f, *rp, states + 10 * i) ? 1 : 0;
}
static const std::vector<double>& esdf_of(Field* f, int which) {
    switch (which) {
        case TOPAY_MAP2D_FLAT: return f->esdf_buffer_2d;
        case TOPAY_MAP2D_INFLATE: return f->esdf_buffer_2d_inflate;
        case TOPAY_MAP2D_CRITICAL: return f->esdf_buffer_2d_critical;
        default: return f->esdf_buffer_3d;
    }
}
void oracle_field_is_collision(void* h, int dim, const double* pos, int64_t n, double thr, int8_t* out) {
    Field* f = (Field*)h;
    for (int64_t i = 0; i < n; i++)
        out[i] = dim == 2 ? f->is_collision2d(pos + 2 * i, thr) : f->is_collision3d(pos + 3 * i, thr);
}
void oracle_field_dist_coarse2d(void* h, const double* pos, int64_t n, int critical, double* out) {
    for (int64_t i = 0; i < n; i++) out[i] = ((Field*)h)->dist_coarse2d(pos + 2 * i, critical != 0);
}
void oracle_field_dist_coarse2i(void* h, const int32_t* idx, int64_t n, int critical, double* out) {
    for (int64_t i = 0; i < n; i++) out[i] = ((Field*)h)->dist_coarse2i(idx + 2 * i, critical != 0);
}
void oracle_field_line_collision2d(void* h, const double* p1, const double* p2, int64_t n, double thr, int8_t* out) {
    for (int64_t i = 0; i < n; i++) out[i] = ((Field*)h)->is_line_collision_grid2d(p1 + 2 * i, p2 + 2 * i, thr);
}
void oracle_field_download(void* h, int which, double* out) {
    const std::vector<double>& b = esdf_of((Field*)h, which);
    std::memcpy(out, b.data(), b.size() * sizeof(double));
}
void oracle_field_download_sqdist(void* h, int which, int32_t* pos_sq, int32_t* neg_sq) {
    Field* f = (Field*)h;
    if (pos_sq) std::memcpy(pos_sq, f->sq_pos[which].data(), f->sq_pos[which].size() * sizeof(int32_t));
    if (neg_sq) std::memcpy(neg_sq, f->sq_neg[which].data(), f->sq_neg[which].size() * sizeof(int32_t));
}
void oracle_field_download_occupancy(void* h, int which, int8_t* out) {
    Field* f = (Field*)h;
    const std::vector<char>& b = which == TOPAY_MAP3D ? f->occ_buffer_3d
                                 : which == TOPAY_MAP2D_CRITICAL ? f->occ_buffer_2d_critical : f->occ_buffer_2d;
    std::memcpy(out, b.data(), b.size());
}

// ------------------------------------------------------------------ robot
int oracle_colli_pts(const topay_robot_params* rp, const double* pos10, double* out48) {
    double pts[TOPAY_NSPHERE][4];
    int n = get_colli_pts(*rp, pos10, pts);
    std::memcpy(out48, pts, sizeof(pts));
    return n;
}
void oracle_colli_grads(const topay_robot_params* rp, const double* pos10, const double* grads36, double* out10) {
    double g[TOPAY_NSPHERE][3];
    std::memcpy(g, grads36, sizeof(g));
    get_colli_grads(*rp, pos10, g, out10);
}

// ------------------------------------------------------------------ minco / banded / lbfgs
// inner: 9 x (N-1) column-major. coeff_out: 6N x 9 row-major.
void oracle_minco_generate(int N, const double* ew, const double* head, const double* tail, const double* inner,
                           const double* T, double* coeff_out, double* jerk_cost, double* gdC_jerk, double* gdT_jerk) {
    MinJerk9 m;
    m.reset(N, ew);
    m.generate(head, tail, inner, T);
    std::memcpy(coeff_out, m.c.data(), m.c.size() * sizeof(double));
    if (jerk_cost) *jerk_cost = m.getTrajJerkCost();
    if (gdC_jerk && gdT_jerk) {
        Vec a, b;
        m.calJerkGradCT(a, b);
        std::memcpy(gdC_jerk, a.data(), a.size() * sizeof(double));
        std::memcpy(gdT_jerk, b.data(), b.size() * sizeof(double));
    }
}
// (gdC, gdT) -> (gdP 9 x (N-1) col-major, gdTail 9 x 3 row-major, gdT updated)
void oracle_minco_backprop(int N, const double* ew, const double* head, const double* tail, const double* inner,
                           const double* T, const double* gdC, double* gdT, double* gdP, double* gdTail) {
    MinJerk9 m;
    m.reset(N, ew);
    m.generate(head, tail, inner, T);
    Vec gc(gdC, gdC + (size_t)6 * N * 9), gt(gdT, gdT + N), gp, gtail;
    m.calGradCTtoQT(gc, gt, gp, gtail);
    std::memcpy(gdT, gt.data(), N * sizeof(double));
    std::memcpy(gdP, gp.data(), gp.size() * sizeof(double));
    std::memcpy(gdTail, gtail.data(), gtail.size() * sizeof(double));
}
// Dense (n x n row-major, band p/q) -> banded LU solve of A X = B and A^T X = B; B is n x m.
void oracle_banded_solve(int n, int p, int q, const double* dense, int m, double* B, int adjoint) {
    Banded A;
    A.create(n, p, q);
    for (int i = 0; i < n; i++)
        for (int j = std::max(0, i - p); j <= std::min(n - 1, i + q); j++) A(i, j) = dense[(size_t)i * n + j];
    A.factorizeLU();
    if (adjoint)
        A.solveAdj(B, m);
    else
        A.solve(B, m);
}
// L-BFGS on f(x) = sum_i [ a_i (x_i - c_i)^2 ] + b * sum_i (x_{i+1} - x_i^2)^2  (known-answer problems)
int oracle_lbfgs_test_problem(int n, const double* a, const double* c, double b, const topay_lbfgs_params* p,
                              double* x, double* f_out, int* iters, int* evals) {
    Vec xv(x, x + n);
    LbfgsStats st;
    double f = 0;
    int r = lbfgs_optimize(
        xv, f,
        [&](const Vec& xx, Vec& g) {
            double s = 0;
            for (int i = 0; i < n; i++) {
                g[i] = 2 * a[i] * (xx[i] - c[i]);
                s += a[i] * (xx[i] - c[i]) * (xx[i] - c[i]);
            }
            for (int i = 0; i + 1 < n; i++) {
                double t = xx[i + 1] - xx[i] * xx[i];
                s += b * t * t;
                g[i + 1] += 2 * b * t;
                g[i] += -4 * b * t * xx[i];
            }
            return s;
        },
        nullptr, *p, &st);
    std::memcpy(x, xv.data(), n * sizeof(double));
    *f_out = f;
    if (iters) *iters = st.iters;
    if (evals) *evals = st.evals;
    return r;
}

// ------------------------------------------------------------------ solve
int oracle_prepare_candidate(const topay_opt_params* opt, const topay_robot_params* rp, const double* init_path,
                             int path_len, const double* bvel, const double* bacc, int max_pieces,
                             int32_t* piece_num, double* head, double* tail, double* sxy, double* exy,
                             double* inner_xy, double* x0, int32_t* s1_past) {
    TrajOpt t;
    t.opt = *opt;
    t.rp = *rp;
    Vec x = t.prepare(init_path, path_len, bvel, bacc);
    *piece_num = t.piece_num;
    if (t.piece_num > max_pieces) return TOPAY_ERR_TOO_LARGE;
    std::memcpy(head, t.minco_start_state, sizeof(t.minco_start_state));
    std::memcpy(tail, t.minco_end_state, sizeof(t.minco_end_state));
    sxy[0] = t.start_state[0];
    sxy[1] = t.start_state[1];
    exy[0] = t.end_state[0];
    exy[1] = t.end_state[1];
    std::memcpy(inner_xy, t.init_inner_xy.data(), t.init_inner_xy.size() * sizeof(double));
    std::memcpy(x0, x.data(), x.size() * sizeof(double));
    if (s1_past) *s1_past = t.s1_past;
    return TOPAY_OK;
}

// One evaluation of one candidate (the parity hook's checker).
void oracle_eval(const topay_opt_params* opt, const topay_robot_params* rp, void* field, int stage, int N,
                 const double* head, const double* tail, const double* sxy, const double* exy,
                 const double* inner_xy, const double* lambda, const double* rho, const double* x, double* cost,
                 double* grad, double* terms, double* coeff_out, double* final_xy) {
    TrajOpt t;
    t.opt = *opt;
    t.rp = *rp;
    t.grid = (Field*)field;
    t.rog = g_rog;
    t.exact_chain = g_exact_chain;
    t.set_problem(N, head, tail, sxy, exy, inner_xy, lambda, rho);
    const int n = topay_num_vars(N);
    Vec xv(x, x + n), g(n, 0.0);
    *cost = t.cost_callback(stage, xv, g);
    std::memcpy(grad, g.data(), n * sizeof(double));
    if (terms) std::memcpy(terms, t.terms, sizeof(t.terms));
    if (coeff_out) std::memcpy(coeff_out, t.minco.c.data(), t.minco.c.size() * sizeof(double));
    if (final_xy) {
        final_xy[0] = t.final_xy_error[0];
        final_xy[1] = t.final_xy_error[1];
    }
}

// Penalty part only (calFirst/SecondStagePenalGrad outputs) for given coefficients: used to check
// the device's per-node maths in isolation. coeff: 6N x 9, T: N.
void oracle_penalty_only(const topay_opt_params* opt, const topay_robot_params* rp, void* field, int stage, int N,
                         const double* coeff, const double* T, const double* sxy, const double* exy,
                         const double* inner_xy, const double* lambda, const double* rho, double* cost, double* gdC,
                         double* gdT, double* terms, double* final_xy) {
    TrajOpt t;
    t.opt = *opt;
    t.rp = *rp;
    t.grid = (Field*)field;
    t.rog = g_rog;
    t.exact_chain = g_exact_chain;
    double head[27] = {0}, tail[27] = {0};
    t.set_problem(N, head, tail, sxy, exy, inner_xy, lambda, rho);
    t.times.assign(T, T + N);
    t.minco.c.assign(coeff, coeff + (size_t)6 * N * 9);
    for (int i = 0; i < TOPAY_NTERMS; i++) t.terms[i] = 0.0;
    Vec gc, gt;
    if (stage == 1)
        t.first_stage_penalty(*cost, gc, gt);
    else
        t.second_stage_penalty(*cost, gc, gt);
    std::memcpy(gdC, gc.data(), gc.size() * sizeof(double));
    std::memcpy(gdT, gt.data(), gt.size() * sizeof(double));
    std::memcpy(terms, t.terms, sizeof(t.terms));
    final_xy[0] = t.final_xy_error[0];
    final_xy[1] = t.final_xy_error[1];
}

struct OracleSolveOut {
    int32_t status, lbfgs_code, piece_num, iters, evals, alm_rounds;
    double cost, duration, final_xy_err[2];
};

// Full optimizeTraj of one candidate. T_out [max_pieces], coeff_out [6*max_pieces*9], x_out [10*max_pieces-8]
// may be NULL. trace_out (4 doubles per accepted iteration: f, step, k, ls) up to trace_cap entries.
int oracle_solve(const topay_opt_params* opt, const topay_robot_params* rp, void* field, const double* init_path,
                 int path_len, const double* bvel, const double* bacc, double wall_cap_s, OracleSolveOut* out,
                 double* T_out, double* coeff_out, double* x_out, double* trace_out, int trace_cap,
                 int* trace_len) {
    TrajOpt t;
    t.opt = *opt;
    t.rp = *rp;
    t.grid = (Field*)field;
    t.rog = g_rog;
    std::vector<double> tr;
    if (trace_out) t.trace = &tr;
    Vec x;
    bool ok = t.optimize_traj(init_path, path_len, bvel, bacc, &x, wall_cap_s);
    out->status = ok ? 1 : 0;
    out->lbfgs_code = t.last_code;
    out->piece_num = t.piece_num;
    out->iters = t.stats.iters;
    out->evals = t.stats.evals;
    out->alm_rounds = t.alm_rounds;
    out->cost = t.traj_cost;
    double dur = 0;
    for (int i = 0; i < t.piece_num; i++) dur += t.minco.T1[i];
    out->duration = dur;
    out->final_xy_err[0] = t.final_xy_error[0];
    out->final_xy_err[1] = t.final_xy_error[1];
    if (T_out) std::memcpy(T_out, t.minco.T1.data(), t.piece_num * sizeof(double));
    if (coeff_out) std::memcpy(coeff_out, t.minco.c.data(), t.minco.c.size() * sizeof(double));
    if (x_out) std::memcpy(x_out, x.data(), x.size() * sizeof(double));
    if (trace_out) {
        int n = std::min((int)tr.size(), trace_cap);
        std::memcpy(trace_out, tr.data(), n * sizeof(double));
        if (trace_len) *trace_len = n;
    }
    return 0;
}

// The reference's thread-per-candidate pattern (planner.cpp:921-925): n_threads workers pull candidates.
// init_paths: candidates back to back; out[n_cand].
int oracle_solve_batch(const topay_opt_params* opt, const topay_robot_params* rp, void* field, int n_cand,
                       const int32_t* path_len, const double* init_paths, const double* bvel, const double* bacc,
                       double wall_cap_s, int n_threads, OracleSolveOut* out) {
    std::vector<size_t> off(n_cand + 1, 0);
    for (int i = 0; i < n_cand; i++) off[i + 1] = off[i] + (size_t)path_len[i] * 10;
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_cand) break;
            oracle_solve(opt, rp, field, init_paths + off[i], path_len[i], bvel + (size_t)i * 20,
                         bacc + (size_t)i * 20, wall_cap_s, &out[i], nullptr, nullptr, nullptr, nullptr, 0, nullptr);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, n_threads); t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    return 0;
}

// Timed evaluations for the CPU baseline of the evaluation microbenchmark: repeats `reps` stage-`stage`
// evaluations of each candidate on n_threads threads.
int oracle_eval_batch(const topay_opt_params* opt, const topay_robot_params* rp, void* field, int stage, int n_cand,
                      const int32_t* piece_num, int max_pieces, const double* head, const double* tail,
                      const double* sxy, const double* exy, const double* inner_xy, const double* lambda,
                      const double* rho, const double* x, int x_stride, int reps, int n_threads, double* cost,
                      double* grad) {
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_cand) break;
            for (int r = 0; r < reps; r++)
                oracle_eval(opt, rp, field, stage, piece_num[i], head + 27 * i, tail + 27 * i, sxy + 2 * i,
                            exy + 2 * i, inner_xy + (size_t)2 * max_pieces * i, lambda ? lambda + 2 * i : nullptr,
                            rho ? rho + 2 * i : nullptr, x + (size_t)x_stride * i, cost + i,
                            grad + (size_t)x_stride * i, nullptr, nullptr, nullptr);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, n_threads); t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    return 0;
}

// ---------------------------------------------------------------- ROG-Map field (oracle_rog.hpp)
void* oracle_rog_create(const topay_rog_desc* d) {
    RogEsdf* r = new RogEsdf();
    r->init(d->half_prob_map_size_i, d->prob_resolution, d->esdf_resolution, d->local_update_box,
            d->map_sliding_en != 0, d->fix_map_origin, d->unk_thresh);
    return r;
}
void oracle_rog_destroy(void* h) { delete (RogEsdf*)h; }
void oracle_rog_geometry(void* h, int32_t* half, int32_t* size, double* res, int32_t* origin_i, int32_t* half_box) {
    RogEsdf* r = (RogEsdf*)h;
    for (int i = 0; i < 3; i++) {
        half[i] = r->half[i];
        size[i] = r->size[i];
        origin_i[i] = r->origin_i[i];
        half_box[i] = r->half_box_i[i];
    }
    *res = r->res;
}
void oracle_rog_slide(void* h, const double* odom) { ((RogEsdf*)h)->slide(odom); }
void oracle_rog_update_counters(void* h, const double* pos, const uint8_t* from, const uint8_t* to, int64_t n) {
    for (int64_t i = 0; i < n; i++) ((RogEsdf*)h)->update_counter(pos + 3 * i, from[i], to[i]);
}
void oracle_rog_set_occupied_cnt(void* h, const int16_t* cnt) {
    RogEsdf* r = (RogEsdf*)h;
    std::memcpy(r->occupied_cnt.data(), cnt, r->vox * sizeof(int16_t));
}
void oracle_rog_download_counters(void* h, int16_t* occ, int16_t* unk) {
    RogEsdf* r = (RogEsdf*)h;
    if (occ) std::memcpy(occ, r->occupied_cnt.data(), r->vox * sizeof(int16_t));
    if (unk) std::memcpy(unk, r->unknown_cnt.data(), r->vox * sizeof(int16_t));
}
void oracle_rog_update_esdf(void* h, const double* odom) { ((RogEsdf*)h)->update_esdf(odom); }
void oracle_rog_query(void* h, int kind, const double* pos, int64_t n, double* dist, double* grad) {
    RogEsdf* r = (RogEsdf*)h;
    for (int64_t i = 0; i < n; i++) {
        const double* p = pos + 3 * i;
        double d = 0, g[3] = {0, 0, 0};
        bool has_grad = true;
        switch (kind) {
            case TOPAY_ROG_Q_EDT: r->value_grad(p, d, g); break;
            case TOPAY_ROG_Q_FLAT: r->value_grad_2d(p, false, d, g); break;
            case TOPAY_ROG_Q_CRITICAL: r->value_grad_2d(p, true, d, g); break;
            case TOPAY_ROG_Q_CELL: d = r->get_distance(p); has_grad = false; break;
            case TOPAY_ROG_Q_CELL_FLAT: d = r->get_distance2d(p); has_grad = false; break;
            default: d = r->get_critical_distance(p); has_grad = false; break;
        }
        dist[i] = d;
        if (grad && has_grad)
            for (int k = 0; k < 3; k++) grad[3 * i + k] = g[k];
    }
}
void oracle_rog_evaluate_edt(void* h, const double* pos, int64_t n, double* dist) {
    for (int64_t i = 0; i < n; i++) dist[i] = ((RogEsdf*)h)->evaluate_edt(pos + 3 * i);
}
void oracle_rog_is_line_free2d(void* h, const double* s, const double* e, int64_t n, double thr, int8_t* out) {
    for (int64_t i = 0; i < n; i++) out[i] = ((RogEsdf*)h)->is_line_free_2d(s + 2 * i, e + 2 * i, thr) ? 1 : 0;
}
void oracle_rog_download(void* h, int which, double* out) {
    RogEsdf* r = (RogEsdf*)h;
    const std::vector<double>& v = which == TOPAY_ROG_BUF_DIST3 ? r->dist3
                                 : which == TOPAY_ROG_BUF_NEG3 ? r->tmp1
                                 : which == TOPAY_ROG_BUF_CRITICAL ? r->dist_crit : r->dist_flat;
    std::memcpy(out, v.data(), v.size() * sizeof(double));
}

// ---------------------------------------------------------------- MomaTraj + feasibility (oracle_traj.hpp)
static MomaTrajO make_traj(const topay_traj_batch* b, int i) {
    PolyTraj p;
    p.N = b->piece_num[i];
    p.T = b->T + (size_t)i * b->max_pieces;
    p.c = b->coeff + (size_t)i * 6 * b->max_pieces * 9;
    MomaTrajO t;
    t.init(p, b->start_se2 + 3 * i);
    return t;
}
int oracle_traj_car_seq(const topay_traj_batch* b, int cap, double* car_seq, int32_t* len) {
    for (int i = 0; i < b->n_traj; i++) {
        MomaTrajO t = make_traj(b, i);
        const int n = (int)(t.car_seq.size() / 4);
        len[i] = n;
        if (n > cap) return -5;
        std::memcpy(car_seq + (size_t)i * cap * 4, t.car_seq.data(), t.car_seq.size() * sizeof(double));
    }
    return 0;
}
void oracle_traj_sample(const topay_traj_batch* b, const double* tt, int m, double* state, double* dstate) {
    for (int i = 0; i < b->n_traj; i++) {
        MomaTrajO t = make_traj(b, i);
        for (int j = 0; j < m; j++) {
            if (state) t.get_state(tt[(size_t)i * m + j], state + ((size_t)i * m + j) * 10);
            if (dstate) t.get_dstate(tt[(size_t)i * m + j], dstate + ((size_t)i * m + j) * 10);
        }
    }
}
void oracle_check_feasible(void* field, const topay_robot_params* rp, const topay_traj_batch* b,
                           topay_feasibility* out) {
    for (int i = 0; i < b->n_traj; i++) {
        MomaTrajO t = make_traj(b, i);
        Feasibility F = check_feasible(t, *rp, (Field*)field, g_rog);
        out->feasible[i] = F.feasible;
        if (out->feasible_print) out->feasible_print[i] = F.feasible_print;
        if (out->n_samples) out->n_samples[i] = F.n_samples;
        if (out->max_vel) out->max_vel[i] = F.max_vel;
        if (out->max_acc) out->max_acc[i] = F.max_acc;
        if (out->max_domega) out->max_domega[i] = F.max_domega;
        if (out->max_d2omega) out->max_d2omega[i] = F.max_d2omega;
        for (int k = 0; k < TOPAY_DOF; k++) {
            if (out->max_q) out->max_q[i * TOPAY_DOF + k] = F.max_q[k];
            if (out->max_dq) out->max_dq[i * TOPAY_DOF + k] = F.max_dq[k];
            if (out->max_d2q) out->max_d2q[i * TOPAY_DOF + k] = F.max_d2q[k];
        }
        if (out->min_dist) out->min_dist[i] = F.min_dist;
        if (out->min_dist_mani)
            for (int k = 0; k < TOPAY_NSPHERE; k++) out->min_dist_mani[i * TOPAY_NSPHERE + k] = F.min_dist_mani[k];
    }
}
int oracle_select_shortest(const int32_t* succ, const double* dur, int n) { return select_shortest(succ, dur, n); }


// ---- ProbMap (row N3)
void* oracle_prob_create(void* esdf, const topay_rog_desc* d, const topay_prob_desc* p) {
    RogProb* m = new RogProb();
    const float pp[6] = {p->p_hit, p->p_miss, p->p_min, p->p_max, p->p_occ, p->p_free};
    m->init((RogEsdf*)esdf, d->half_prob_map_size_i, d->prob_resolution, d->map_sliding_en != 0, p->map_sliding_thresh,
            d->fix_map_origin, pp, p->raycast_range_min, p->raycast_range_max, p->virtual_ceil_height,
            p->virtual_ground_height, p->inflation_resolution, p->inflation_step, p->local_update_box, p->point_filt_num, p->batch_update_size, p->intensity_thresh,
            p->raycasting_en != 0);
    return m;
}
void oracle_prob_destroy(void* h) { delete (RogProb*)h; }
void oracle_prob_update(void* h, const float* cloud, int64_t n, const double* pos) { ((RogProb*)h)->update(cloud, n, pos); }
void oracle_prob_set_first_frame(void* h, int armed) { ((RogProb*)h)->first_frame = armed != 0; }
void oracle_prob_download(void* h, float* occ, int32_t* origin_i) {
    RogProb* m = (RogProb*)h;
    std::memcpy(occ, m->occupancy.data(), m->occupancy.size() * sizeof(float));
    for (int i = 0; i < 3; i++) origin_i[i] = m->origin_i[i];
}
void oracle_prob_size(void* h, int32_t* size) {
    for (int i = 0; i < 3; i++) size[i] = ((RogProb*)h)->size[i];
}

// ------------------------------------------------------------------ front-end pieces (row N2)
int oracle_dense_path(const double* raw_xy, int n, double step_size, double start_yaw, double end_yaw, double v_max,
                      double w_max, double* out, int cap) {
    const auto r = dense_path(raw_xy, n, step_size, start_yaw, end_yaw, v_max, w_max);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++)
        for (int k = 0; k < 4; k++) out[4 * i + k] = r[i][k];
    return (int)r.size();
}
void oracle_line_visib(void* h, const double* p1, const double* p2, int64_t n, double thresh, int use_critical,
                       int8_t* visible, double* pc) {
    const Field* f = (const Field*)h;
    for (int64_t i = 0; i < n; i++) visible[i] = line_visib(*f, p1 + 3 * i, p2 + 3 * i, thresh, use_critical != 0, pc + 3 * i) ? 1 : 0;
}

int oracle_discretize_path(const double* path, int n, int pt_num, double* out) {
    const auto r = discretize_path(path, n, pt_num);
    for (size_t i = 0; i < r.size(); i++)
        for (int k = 0; k < 3; k++) out[3 * i + k] = r[i][k];
    return (int)r.size();
}
double oracle_path_length(const double* path, int n) { return path_length(path, n); }
int oracle_same_topo_path(void* h, const double* p1, int n1, const double* p2, int n2, double thresh, int use_critical) {
    return same_topo_path(*(const Field*)h, p1, n1, p2, n2, thresh, use_critical != 0) ? 1 : 0;
}

}  // extern "C"
