// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C entry points over the REFERENCE'S OWN hot-path headers, compiled unmodified from where they lie
// under /root/reference (oracle/Makefile target `_ref`; include paths only, nothing is copied):
//   src/planner/include/utils/lbfgs.hpp            lbfgs_optimize, line_search_lewisoverton
//   src/planner/include/utils/banded_system.hpp    BandedSystem
//   src/planner/include/utils/minco.hpp            nmoma_planner::MinJerkOpt<9> (+ root_finder.hpp)
//   src/simulator/fake_moma/include/fake_moma/moma_param.h   MomaParam::getColliPts / getColliGrads
// against the Eigen / ROS stand-ins of oracle/ref_stubs (the image has neither). Output:
// oracle/_ref/libtopay_ref.so, used by tests/test_ref_pin.py to pin the oracle restatement: the entry
// points mirror oracle_capi.cpp's (same argument layouts), so the two can be compared bit for bit.
#include <cstring>
#include <vector>

#include "utils/banded_system.hpp"
#include "utils/lbfgs.hpp"
#include "utils/minco.hpp"
#include "fake_moma/moma_param.h"

#include "../include/topay_b200.h"

namespace {

lbfgs::lbfgs_parameter_t to_ref(const topay_lbfgs_params& p) {
    lbfgs::lbfgs_parameter_t r;
    r.mem_size = p.mem_size;
    r.g_epsilon = p.g_epsilon;
    r.past = p.past;
    r.delta = p.delta;
    r.max_iterations = p.max_iterations;
    r.max_linesearch = p.max_linesearch;
    r.min_step = p.min_step;
    r.max_step = p.max_step;
    r.f_dec_coeff = p.f_dec_coeff;
    r.s_curv_coeff = p.s_curv_coeff;
    r.cautious_factor = p.cautious_factor;
    r.machine_prec = p.machine_prec;
    return r;
}

// f(x) = sum_i a_i (x_i - c_i)^2 + b sum_i (x_{i+1} - x_i^2)^2, the same expression order as
// oracle_lbfgs_test_problem
struct Problem {
    int n;
    const double *a, *c;
    double b;
    int evals = 0, iters = 0;
    double* trace = nullptr;   // (fx, step, k, ls) per accepted iteration
    int trace_cap = 0, trace_len = 0;
    int cancel_after = 0;      // progress callback returns 1 once k exceeds this (0 = never)
};

double eval_cb(void* inst, const Eigen::VectorXd& x, Eigen::VectorXd& g) {
    Problem& P = *(Problem*)inst;
    P.evals++;
    double s = 0;
    for (int i = 0; i < P.n; i++) {
        g[i] = 2 * P.a[i] * (x[i] - P.c[i]);
        s += P.a[i] * (x[i] - P.c[i]) * (x[i] - P.c[i]);
    }
    for (int i = 0; i + 1 < P.n; i++) {
        double t = x[i + 1] - x[i] * x[i];
        s += P.b * t * t;
        g[i + 1] += 2 * P.b * t;
        g[i] += -4 * P.b * t * x[i];
    }
    return s;
}

int progress_cb(void* inst, const Eigen::VectorXd&, const Eigen::VectorXd&, const double fx, const double step,
                const int k, const int ls) {
    Problem& P = *(Problem*)inst;
    P.iters = k;
    if (P.trace && P.trace_len < P.trace_cap) {
        double* t = P.trace + 4 * P.trace_len++;
        t[0] = fx;
        t[1] = step;
        t[2] = k;
        t[3] = ls;
    }
    return P.cancel_after > 0 && k > P.cancel_after;
}

MomaParam make_robot() { return MomaParam(); }

}  // namespace

extern "C" {

const char* ref_sources() {
    return "lbfgs.hpp banded_system.hpp minco.hpp(MinJerkOpt<9>) moma_param.h(getColliPts,getColliGrads) "
           "compiled unmodified from /root/reference against oracle/ref_stubs";
}

// ------------------------------------------------------------------ lbfgs.hpp
int ref_lbfgs_test_problem(int n, const double* a, const double* c, double b, const topay_lbfgs_params* p, double* x,
                           double* f_out, int* iters, int* evals, double* trace, int trace_cap, int* trace_len,
                           int cancel_after) {
    Eigen::VectorXd xv(n);
    for (int i = 0; i < n; i++) xv[i] = x[i];
    Problem P{n, a, c, b};
    P.trace = trace;
    P.trace_cap = trace_cap;
    P.cancel_after = cancel_after;
    double f = 0;
    const int r = lbfgs::lbfgs_optimize(xv, f, eval_cb, nullptr, progress_cb, &P, to_ref(*p));
    for (int i = 0; i < n; i++) x[i] = xv[i];
    *f_out = f;
    if (iters) *iters = P.iters;
    if (evals) *evals = P.evals;
    if (trace_len) *trace_len = P.trace_len;
    return r;
}

// ------------------------------------------------------------------ banded_system.hpp
// dense: n x n row-major; B: n x m row-major, overwritten with the solution
void ref_banded_solve(int n, int p, int q, const double* dense, int m, double* B, int adjoint) {
    BandedSystem A;
    A.create(n, p, q);
    for (int i = 0; i < n; i++)
        for (int j = std::max(0, i - p); j <= std::min(n - 1, i + q); j++) A(i, j) = dense[(size_t)i * n + j];
    A.factorizeLU();
    Eigen::MatrixXd b(n, m);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) b(i, j) = B[(size_t)i * m + j];
    if (adjoint)
        A.solveAdj(b);
    else
        A.solve(b);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) B[(size_t)i * m + j] = b(i, j);
    A.destroy();
}

// ------------------------------------------------------------------ minco.hpp, MinJerkOpt<9>
static void load_minco(nmoma_planner::MinJerkOpt<9>& m, int N, const double* ew, const double* head, const double* tail,
                       const double* inner, const double* T) {
    Eigen::VectorXd w(9), ts(N);
    for (int d = 0; d < 9; d++) w[d] = ew[d];
    for (int i = 0; i < N; i++) ts[i] = T[i];
    Eigen::MatrixXd H(9, 3), Tl(9, 3), In(9, std::max(N - 1, 0));
    for (int d = 0; d < 9; d++)
        for (int k = 0; k < 3; k++) {
            H(d, k) = head[d * 3 + k];
            Tl(d, k) = tail[d * 3 + k];
        }
    for (int i = 0; i < N - 1; i++)
        for (int d = 0; d < 9; d++) In(d, i) = inner[(size_t)i * 9 + d];
    m.reset(N, w);
    m.generate(H, Tl, In, ts);
}
static void store_rowmajor(const Eigen::MatrixXd& M, double* out) {
    for (Eigen::Index i = 0; i < M.rows(); i++)
        for (Eigen::Index j = 0; j < M.cols(); j++) out[(size_t)i * M.cols() + j] = M(i, j);
}

void ref_minco_generate(int N, const double* ew, const double* head, const double* tail, const double* inner,
                        const double* T, double* coeff_out, double* jerk_cost, double* gdC_jerk, double* gdT_jerk) {
    nmoma_planner::MinJerkOpt<9> m;
    load_minco(m, N, ew, head, tail, inner, T);
    store_rowmajor(m.getCoeffs(), coeff_out);
    if (jerk_cost) *jerk_cost = m.getTrajJerkCost();
    if (gdC_jerk && gdT_jerk) {
        Eigen::MatrixXd gc;
        Eigen::VectorXd gt;
        m.calJerkGradCT(gc, gt);
        store_rowmajor(gc, gdC_jerk);
        for (int i = 0; i < N; i++) gdT_jerk[i] = gt[i];
    }
}

// (gdC 6N x 9 row-major, gdT) -> gdP (N-1) x 9 (row i = inner point i), gdTail 9 x 3 row-major, gdT updated
void ref_minco_backprop(int N, const double* ew, const double* head, const double* tail, const double* inner,
                        const double* T, const double* gdC, double* gdT, double* gdP, double* gdTail) {
    nmoma_planner::MinJerkOpt<9> m;
    load_minco(m, N, ew, head, tail, inner, T);
    Eigen::MatrixXd gc(6 * N, 9), gp, gtail;
    Eigen::VectorXd gt(N);
    for (int i = 0; i < 6 * N; i++)
        for (int d = 0; d < 9; d++) gc(i, d) = gdC[(size_t)i * 9 + d];
    for (int i = 0; i < N; i++) gt[i] = gdT[i];
    m.calGradCTtoQT(gc, gt, gp, gtail);
    for (int i = 0; i < N; i++) gdT[i] = gt[i];
    for (int i = 0; i < N - 1; i++)
        for (int d = 0; d < 9; d++) gdP[(size_t)i * 9 + d] = gp(d, i);
    store_rowmajor(gtail, gdTail);
}

// ------------------------------------------------------------------ moma_param.h
int ref_colli_pts(const double* pos10, double* out48) {
    MomaParam rp = make_robot();
    Eigen::VectorXd pos(10);
    for (int i = 0; i < 10; i++) pos[i] = pos10[i];
    const std::vector<Eigen::Vector4d> pts = rp.getColliPts(pos);
    for (size_t c = 0; c < pts.size() && c < 12; c++)
        for (int k = 0; k < 4; k++) out48[c * 4 + k] = pts[c][k];
    return (int)pts.size();
}
void ref_colli_grads(const double* pos10, const double* grads36, int n_pts, double* out10) {
    MomaParam rp = make_robot();
    Eigen::VectorXd pos(10);
    for (int i = 0; i < 10; i++) pos[i] = pos10[i];
    std::vector<Eigen::Vector3d> g;
    for (int c = 0; c < n_pts; c++) g.push_back(Eigen::Vector3d(grads36[3 * c], grads36[3 * c + 1], grads36[3 * c + 2]));
    const Eigen::VectorXd out = rp.getColliGrads(pos, g);
    for (int i = 0; i < 10; i++) out10[i] = out[i];
}
// robot constants as the reference's constructor computes them, in topay_robot_params layout order
void ref_robot_constants(double* colli_length8, double* colli_points16, double* colli_radius16, int* link_map,
                         double* limits /*max_v,max_a,max_w,max_dw,chassis_r,chassis_h*/, double* joint_pos_max7,
                         double* relative_t3, double* relative_R9, int* collision_matrix144) {
    MomaParam rp = make_robot();
    for (int i = 0; i < 8; i++) colli_length8[i] = rp.colli_length[i];
    for (int i = 0; i < 16; i++) {
        colli_points16[i] = rp.colli_points[i];
        colli_radius16[i] = rp.colli_point_radius[i];
    }
    for (Eigen::Index i = 0; i < rp.colli_link_map.size() && i < 12; i++) link_map[i] = rp.colli_link_map[i];
    limits[0] = rp.max_v; limits[1] = rp.max_a; limits[2] = rp.max_w; limits[3] = rp.max_dw;
    limits[4] = rp.chassis_colli_radius; limits[5] = rp.chassis_height;
    for (int i = 0; i < 7; i++) joint_pos_max7[i] = rp.joint_pos_limit_max[i];
    for (int i = 0; i < 3; i++) relative_t3[i] = rp.relative_t[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) relative_R9[i * 3 + j] = rp.relative_R(i, j);
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 12; j++) collision_matrix144[i * 12 + j] = rp.collision_matrix(i, j);
}

}  // extern "C"
