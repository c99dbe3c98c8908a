// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the reference's NLP solve in plain C++17, in the reference's
// operation order (scalar loops where the reference uses Eigen reductions; Eigen's
// packet-wise reduction order may differ from these at the last ulp).
//
// PARITY UNPINNED (see oracle_field.hpp): no reference test pins this path. Pinned
// instead by finite differences of the oracle's own cost, dense-solver checks of
// the banded LU, and known-answer L-BFGS problems (tests/test_oracle_solve.py).
//
// Reference files restated here (paths relative to /root/reference):
//   src/planner/include/utils/banded_system.hpp:14-146        -> Banded
//   src/planner/include/utils/minco.hpp:772-1070              -> MinJerk9
//   src/planner/include/utils/lbfgs.hpp:276-389, 439-722      -> line_search_lewisoverton, lbfgs_optimize
//   src/planner/include/planner/moma_traj_opt.h:676-842       -> trapezoid timing, C2 maps, smoothL1
//   src/planner/src/moma_traj_opt.cpp:142-498                 -> TrajOpt::optimize_traj
//   src/planner/src/moma_traj_opt.cpp:817-955                 -> TrajOpt::cost_callback
//   src/planner/src/moma_traj_opt.cpp:957-1198                -> TrajOpt::first_stage_penalty
//   src/planner/src/moma_traj_opt.cpp:1200-1829               -> TrajOpt::second_stage_penalty
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <functional>
#include <vector>

#include "../include/topay_b200.h"
#include "oracle_field.hpp"
#include "oracle_robot.hpp"
#include "oracle_rog.hpp"

namespace oracle {

typedef std::vector<double> Vec;

// ---------------------------------------------------------------- banded LU
// banded_system.hpp: storage ptr[(i - j + upperBw) * N + j], no pivoting, exact
// zeros skipped.
struct Banded {
    int N = 0, lowerBw = 0, upperBw = 0;
    std::vector<double> data;
    void create(int n, int p, int q) {
        N = n;
        lowerBw = p;
        upperBw = q;
        data.assign((size_t)N * (lowerBw + upperBw + 1), 0.0);
    }
    void reset() { std::fill(data.begin(), data.end(), 0.0); }
    inline double& operator()(int i, int j) { return data[(size_t)(i - j + upperBw) * N + j]; }
    inline double operator()(int i, int j) const { return data[(size_t)(i - j + upperBw) * N + j]; }

    void factorizeLU() {
        int iM, jM;
        double cVl;
        for (int k = 0; k <= N - 2; k++) {
            iM = std::min(k + lowerBw, N - 1);
            cVl = (*this)(k, k);
            for (int i = k + 1; i <= iM; i++)
                if ((*this)(i, k) != 0.0) (*this)(i, k) /= cVl;
            jM = std::min(k + upperBw, N - 1);
            for (int j = k + 1; j <= jM; j++) {
                cVl = (*this)(k, j);
                if (cVl != 0.0)
                    for (int i = k + 1; i <= iM; i++)
                        if ((*this)(i, k) != 0.0) (*this)(i, j) -= (*this)(i, k) * cVl;
            }
        }
    }
    // b is N x m row-major
    void solve(double* b, int m) const {
        int iM;
        for (int j = 0; j <= N - 1; j++) {
            iM = std::min(j + lowerBw, N - 1);
            for (int i = j + 1; i <= iM; i++)
                if ((*this)(i, j) != 0.0) {
                    const double a = (*this)(i, j);
                    for (int c = 0; c < m; c++) b[i * m + c] -= a * b[j * m + c];
                }
        }
        for (int j = N - 1; j >= 0; j--) {
            const double dj = (*this)(j, j);
            for (int c = 0; c < m; c++) b[j * m + c] /= dj;
            iM = std::max(0, j - upperBw);
            for (int i = iM; i <= j - 1; i++)
                if ((*this)(i, j) != 0.0) {
                    const double a = (*this)(i, j);
                    for (int c = 0; c < m; c++) b[i * m + c] -= a * b[j * m + c];
                }
        }
    }
    void solveAdj(double* b, int m) const {
        int iM;
        for (int j = 0; j <= N - 1; j++) {
            const double dj = (*this)(j, j);
            for (int c = 0; c < m; c++) b[j * m + c] /= dj;
            iM = std::min(j + upperBw, N - 1);
            for (int i = j + 1; i <= iM; i++)
                if ((*this)(j, i) != 0.0) {
                    const double a = (*this)(j, i);
                    for (int c = 0; c < m; c++) b[i * m + c] -= a * b[j * m + c];
                }
        }
        for (int j = N - 1; j >= 0; j--) {
            iM = std::max(0, j - lowerBw);
            for (int i = iM; i <= j - 1; i++)
                if ((*this)(j, i) != 0.0) {
                    const double a = (*this)(j, i);
                    for (int c = 0; c < m; c++) b[i * m + c] -= a * b[j * m + c];
                }
        }
    }
};

// ---------------------------------------------------------------- MINCO
// minco.hpp:772-1070 with Dim = 9. c is (6N) x 9 row-major.
struct MinJerk9 {
    static const int Dim = 9;
    int N = 0;
    Banded A;
    Vec c;
    Vec T1, T2, T3, T4, T5;
    double ew[Dim];

    void reset(int pieceNum, const double* energy_weights) {
        N = pieceNum;
        A.create(6 * N, 6, 6);
        c.assign((size_t)6 * N * Dim, 0.0);
        T1.assign(N, 0);
        T2.assign(N, 0);
        T3.assign(N, 0);
        T4.assign(N, 0);
        T5.assign(N, 0);
        for (int d = 0; d < Dim; d++) ew[d] = energy_weights[d];
    }
    inline double& C(int r, int d) { return c[(size_t)r * Dim + d]; }
    inline double C(int r, int d) const { return c[(size_t)r * Dim + d]; }

    // headPVA/tailPVA: 9 x 3 row-major; inPs: 9 x (N-1) column-major (inPs[i*9+d]).
    void generate(const double* headPVA, const double* tailPVA, const double* inPs, const double* ts) {
        for (int i = 0; i < N; i++) {
            T1[i] = ts[i];
            T2[i] = T1[i] * T1[i];
            T3[i] = T2[i] * T1[i];
            T4[i] = T2[i] * T2[i];
            T5[i] = T4[i] * T1[i];
        }
        A.reset();
        std::fill(c.begin(), c.end(), 0.0);
        A(0, 0) = 1.0;
        A(1, 1) = 1.0;
        A(2, 2) = 2.0;
        for (int d = 0; d < Dim; d++) {
            C(0, d) = headPVA[d * 3 + 0];
            C(1, d) = headPVA[d * 3 + 1];
            C(2, d) = headPVA[d * 3 + 2];
        }
        for (int i = 0; i < N - 1; i++) {
            A(6 * i + 3, 6 * i + 3) = 6.0;
            A(6 * i + 3, 6 * i + 4) = 24.0 * T1[i];
            A(6 * i + 3, 6 * i + 5) = 60.0 * T2[i];
            A(6 * i + 3, 6 * i + 9) = -6.0;
            A(6 * i + 4, 6 * i + 4) = 24.0;
            A(6 * i + 4, 6 * i + 5) = 120.0 * T1[i];
            A(6 * i + 4, 6 * i + 10) = -24.0;
            A(6 * i + 5, 6 * i) = 1.0;
            A(6 * i + 5, 6 * i + 1) = T1[i];
            A(6 * i + 5, 6 * i + 2) = T2[i];
            A(6 * i + 5, 6 * i + 3) = T3[i];
            A(6 * i + 5, 6 * i + 4) = T4[i];
            A(6 * i + 5, 6 * i + 5) = T5[i];
            A(6 * i + 6, 6 * i) = 1.0;
            A(6 * i + 6, 6 * i + 1) = T1[i];
            A(6 * i + 6, 6 * i + 2) = T2[i];
            A(6 * i + 6, 6 * i + 3) = T3[i];
            A(6 * i + 6, 6 * i + 4) = T4[i];
            A(6 * i + 6, 6 * i + 5) = T5[i];
            A(6 * i + 6, 6 * i + 6) = -1.0;
            A(6 * i + 7, 6 * i + 1) = 1.0;
            A(6 * i + 7, 6 * i + 2) = 2 * T1[i];
            A(6 * i + 7, 6 * i + 3) = 3 * T2[i];
            A(6 * i + 7, 6 * i + 4) = 4 * T3[i];
            A(6 * i + 7, 6 * i + 5) = 5 * T4[i];
            A(6 * i + 7, 6 * i + 7) = -1.0;
            A(6 * i + 8, 6 * i + 2) = 2.0;
            A(6 * i + 8, 6 * i + 3) = 6 * T1[i];
            A(6 * i + 8, 6 * i + 4) = 12 * T2[i];
            A(6 * i + 8, 6 * i + 5) = 20 * T3[i];
            A(6 * i + 8, 6 * i + 8) = -2.0;
            for (int d = 0; d < Dim; d++) C(6 * i + 5, d) = inPs[(size_t)i * Dim + d];
        }
        A(6 * N - 3, 6 * N - 6) = 1.0;
        A(6 * N - 3, 6 * N - 5) = T1[N - 1];
        A(6 * N - 3, 6 * N - 4) = T2[N - 1];
        A(6 * N - 3, 6 * N - 3) = T3[N - 1];
        A(6 * N - 3, 6 * N - 2) = T4[N - 1];
        A(6 * N - 3, 6 * N - 1) = T5[N - 1];
        A(6 * N - 2, 6 * N - 5) = 1.0;
        A(6 * N - 2, 6 * N - 4) = 2 * T1[N - 1];
        A(6 * N - 2, 6 * N - 3) = 3 * T2[N - 1];
        A(6 * N - 2, 6 * N - 2) = 4 * T3[N - 1];
        A(6 * N - 2, 6 * N - 1) = 5 * T4[N - 1];
        A(6 * N - 1, 6 * N - 4) = 2;
        A(6 * N - 1, 6 * N - 3) = 6 * T1[N - 1];
        A(6 * N - 1, 6 * N - 2) = 12 * T2[N - 1];
        A(6 * N - 1, 6 * N - 1) = 20 * T3[N - 1];
        for (int d = 0; d < Dim; d++) {
            C(6 * N - 3, d) = tailPVA[d * 3 + 0];
            C(6 * N - 2, d) = tailPVA[d * 3 + 1];
            C(6 * N - 1, d) = tailPVA[d * 3 + 2];
        }
        A.factorizeLU();
        A.solve(c.data(), Dim);
    }

    // (row_a * W) . row_b
    inline double wdot(int ra, int rb) const {
        double s = 0.0;
        for (int d = 0; d < Dim; d++) s += (C(ra, d) * ew[d]) * C(rb, d);
        return s;
    }

    // minco.hpp:923-942
    double getTrajJerkCost() const {
        double energy = 0.0;
        for (int i = 0; i < N; i++) {
            energy += 36.0 * wdot(6 * i + 3, 6 * i + 3) * T1[i] +
                      144.0 * wdot(6 * i + 4, 6 * i + 3) * T2[i] +
                      192.0 * wdot(6 * i + 4, 6 * i + 4) * T3[i] +
                      240.0 * wdot(6 * i + 5, 6 * i + 3) * T3[i] +
                      720.0 * wdot(6 * i + 5, 6 * i + 4) * T4[i] +
                      720.0 * wdot(6 * i + 5, 6 * i + 5) * T5[i];
        }
        return energy;
    }

    // minco.hpp:951-996. gdC: 6N x 9 row-major, gdT: N.
    void calJerkGradCT(Vec& gdC, Vec& gdT) const {
        gdC.assign((size_t)6 * N * Dim, 0.0);
        for (int i = 0; i < N; i++)
            for (int d = 0; d < Dim; d++) {
                const double c3 = C(6 * i + 3, d), c4 = C(6 * i + 4, d), c5 = C(6 * i + 5, d);
                gdC[(size_t)(6 * i + 5) * Dim + d] =
                    240.0 * c3 * ew[d] * T3[i] + 720.0 * c4 * ew[d] * T4[i] + 1440.0 * c5 * ew[d] * T5[i];
                gdC[(size_t)(6 * i + 4) * Dim + d] =
                    144.0 * c3 * ew[d] * T2[i] + 384.0 * c4 * ew[d] * T3[i] + 720.0 * c5 * ew[d] * T4[i];
                gdC[(size_t)(6 * i + 3) * Dim + d] =
                    72.0 * c3 * ew[d] * T1[i] + 144.0 * c4 * ew[d] * T2[i] + 240.0 * c5 * ew[d] * T3[i];
            }
        gdT.assign(N, 0.0);
        for (int i = 0; i < N; i++) {
            gdT[i] = 36.0 * wdot(6 * i + 3, 6 * i + 3) + 288.0 * wdot(6 * i + 4, 6 * i + 3) * T1[i] +
                     576.0 * wdot(6 * i + 4, 6 * i + 4) * T2[i] + 720.0 * wdot(6 * i + 5, 6 * i + 3) * T2[i] +
                     2880.0 * wdot(6 * i + 5, 6 * i + 4) * T3[i] + 3600.0 * wdot(6 * i + 5, 6 * i + 5) * T4[i];
        }
    }

    // minco.hpp:1000-1069. gdP: 9 x (N-1) column-major; gdTail: 9 x 3 row-major.
    void calGradCTtoQT(const Vec& gdC, Vec& gdT, Vec& gdP, Vec& gdTail) const {
        gdP.assign((size_t)Dim * std::max(N - 1, 0), 0.0);
        Vec adj = gdC;
        A.solveAdj(adj.data(), Dim);
        auto AD = [&](int r, int d) { return adj[(size_t)r * Dim + d]; };
        for (int i = 0; i < N - 1; i++)
            for (int d = 0; d < Dim; d++) gdP[(size_t)i * Dim + d] = AD(6 * i + 5, d);
        gdTail.assign(Dim * 3, 0.0);
        for (int d = 0; d < Dim; d++)
            for (int r = 0; r < 3; r++) gdTail[d * 3 + r] = AD(6 * N - 3 + r, d);
        double B1[6][Dim], B2[3][Dim];
        for (int i = 0; i < N - 1; i++) {
            for (int d = 0; d < Dim; d++) {
                B1[2][d] = -(C(i * 6 + 1, d) + 2.0 * T1[i] * C(i * 6 + 2, d) + 3.0 * T2[i] * C(i * 6 + 3, d) +
                             4.0 * T3[i] * C(i * 6 + 4, d) + 5.0 * T4[i] * C(i * 6 + 5, d));
                B1[3][d] = B1[2][d];
                B1[4][d] = -(2.0 * C(i * 6 + 2, d) + 6.0 * T1[i] * C(i * 6 + 3, d) +
                             12.0 * T2[i] * C(i * 6 + 4, d) + 20.0 * T3[i] * C(i * 6 + 5, d));
                B1[5][d] = -(6.0 * C(i * 6 + 3, d) + 24.0 * T1[i] * C(i * 6 + 4, d) + 60.0 * T2[i] * C(i * 6 + 5, d));
                B1[0][d] = -(24.0 * C(i * 6 + 4, d) + 120.0 * T1[i] * C(i * 6 + 5, d));
                B1[1][d] = -120.0 * C(i * 6 + 5, d);
            }
            // Eigen's sum() over a column-major 6x9 expression: column by column.
            double s = 0.0;
            for (int d = 0; d < Dim; d++)
                for (int r = 0; r < 6; r++) s += B1[r][d] * AD(6 * i + 3 + r, d);
            gdT[i] += s;
        }
        for (int d = 0; d < Dim; d++) {
            B2[0][d] = -(C(6 * N - 5, d) + 2.0 * T1[N - 1] * C(6 * N - 4, d) + 3.0 * T2[N - 1] * C(6 * N - 3, d) +
                         4.0 * T3[N - 1] * C(6 * N - 2, d) + 5.0 * T4[N - 1] * C(6 * N - 1, d));
            B2[1][d] = -(2.0 * C(6 * N - 4, d) + 6.0 * T1[N - 1] * C(6 * N - 3, d) +
                         12.0 * T2[N - 1] * C(6 * N - 2, d) + 20.0 * T3[N - 1] * C(6 * N - 1, d));
            B2[2][d] = -(6.0 * C(6 * N - 3, d) + 24.0 * T1[N - 1] * C(6 * N - 2, d) + 60.0 * T2[N - 1] * C(6 * N - 1, d));
        }
        double s = 0.0;
        for (int d = 0; d < Dim; d++)
            for (int r = 0; r < 3; r++) s += B2[r][d] * AD(6 * N - 3 + r, d);
        gdT[N - 1] += s;
    }
};

// ---------------------------------------------------------------- L-BFGS
typedef std::function<double(const Vec&, Vec&)> EvalFn;
typedef std::function<int(const Vec&, const Vec&, double, double, int, int)> ProgressFn;

struct LbfgsStats {
    int iters = 0;
    int evals = 0;
};

inline double vdot(const Vec& a, const Vec& b) {
    double s = 0.0;
    for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
    return s;
}
inline double vdot(const double* a, const double* b, int n) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
inline double absmax(const Vec& a) {
    double m = 0.0;
    for (double v : a) m = std::max(m, std::fabs(v));
    return m;
}

// lbfgs.hpp:276-389
inline int line_search_lewisoverton(Vec& x, double& f, Vec& g, double& stp, const Vec& s, const Vec& xp,
                                    const Vec& gp, const double stpmin, const double stpmax,
                                    const EvalFn& eval, const topay_lbfgs_params& param, LbfgsStats* st) {
    int count = 0;
    bool brackt = false, touched = false;
    double finit, dginit, dgtest, dstest;
    double mu = 0.0, nu = stpmax;
    if (!(stp > 0.0)) return TOPAY_LBFGSERR_INVALIDPARAMETERS;
    dginit = vdot(gp, s);
    if (0.0 < dginit) return TOPAY_LBFGSERR_INCREASEGRADIENT;
    finit = f;
    dgtest = param.f_dec_coeff * dginit;
    dstest = param.s_curv_coeff * dginit;
    const size_t n = x.size();
    while (true) {
        for (size_t i = 0; i < n; i++) x[i] = xp[i] + stp * s[i];
        f = eval(x, g);
        ++count;
        if (st) st->evals++;
        if (std::isinf(f) || std::isnan(f)) return TOPAY_LBFGSERR_INVALID_FUNCVAL;
        // reference-specific early accept, lbfgs.hpp:327-330
        if (param.past > 0 && std::fabs(finit - f) / (std::fabs(finit) + 1.0) < param.delta / param.past)
            return count;
        if (f > finit + stp * dgtest) {
            nu = stp;
            brackt = true;
        } else {
            if (vdot(g, s) < dstest) {
                mu = stp;
            } else {
                return count;
            }
        }
        if (param.max_linesearch <= count) return TOPAY_LBFGSERR_MAXIMUMLINESEARCH;
        if (brackt && (nu - mu) < param.machine_prec * nu) return TOPAY_LBFGSERR_WIDTHTOOSMALL;
        if (brackt)
            stp = 0.5 * (mu + nu);
        else
            stp *= 2.0;
        if (stp < stpmin) return TOPAY_LBFGSERR_MINIMUMSTEP;
        if (stp > stpmax) {
            if (touched) return TOPAY_LBFGSERR_MAXIMUMSTEP;
            touched = true;
            stp = stpmax;
        }
    }
}

// lbfgs.hpp:439-722 (no step-bound callback: the reference passes nullptr)
inline int lbfgs_optimize(Vec& x, double& f, const EvalFn& eval, const ProgressFn& progress,
                          const topay_lbfgs_params& param, LbfgsStats* st = nullptr) {
    int ret, i, j, k, ls, end, bound;
    double step, step_min, step_max, fx, ys, yy;
    double gnorm_inf, xnorm_inf, beta, rate, cau;
    const int n = (int)x.size();
    const int m = param.mem_size;
    if (n <= 0) return TOPAY_LBFGSERR_INVALID_N;
    if (m <= 0) return TOPAY_LBFGSERR_INVALID_MEMSIZE;
    if (param.g_epsilon < 0.0) return TOPAY_LBFGSERR_INVALID_GEPSILON;
    if (param.past < 0) return TOPAY_LBFGSERR_INVALID_TESTPERIOD;
    if (param.delta < 0.0) return TOPAY_LBFGSERR_INVALID_DELTA;
    if (param.min_step < 0.0) return TOPAY_LBFGSERR_INVALID_MINSTEP;
    if (param.max_step < param.min_step) return TOPAY_LBFGSERR_INVALID_MAXSTEP;
    if (!(param.f_dec_coeff > 0.0 && param.f_dec_coeff < 1.0)) return TOPAY_LBFGSERR_INVALID_FDECCOEFF;
    if (!(param.s_curv_coeff < 1.0 && param.s_curv_coeff > param.f_dec_coeff))
        return TOPAY_LBFGSERR_INVALID_SCURVCOEFF;
    if (!(param.machine_prec > 0.0)) return TOPAY_LBFGSERR_INVALID_MACHINEPREC;
    if (param.max_linesearch <= 0) return TOPAY_LBFGSERR_INVALID_MAXLINESEARCH;

    Vec xp(n), g(n), gp(n), d(n), pf(std::max(1, param.past));
    Vec lm_alpha(m, 0.0), lm_ys(m, 0.0);
    std::vector<double> lm_s((size_t)n * m, 0.0), lm_y((size_t)n * m, 0.0);  // column j at [j*n]

    fx = eval(x, g);
    if (st) st->evals++;
    pf[0] = fx;
    for (i = 0; i < n; i++) d[i] = -g[i];
    gnorm_inf = absmax(g);
    xnorm_inf = absmax(x);
    k = 0;
    if (gnorm_inf / std::max(1.0, xnorm_inf) < param.g_epsilon) {
        ret = TOPAY_LBFGS_CONVERGENCE;
    } else {
        step = 1.0 / std::sqrt(vdot(d, d));
        k = 1;
        end = 0;
        bound = 0;
        while (true) {
            xp = x;
            gp = g;
            step_min = param.min_step;
            step_max = param.max_step;
            ls = line_search_lewisoverton(x, fx, g, step, d, xp, gp, step_min, step_max, eval, param, st);
            if (ls < 0) {
                x = xp;
                g = gp;
                ret = ls;
                break;
            }
            if (progress) {
                if (progress(x, g, fx, step, k, ls)) {
                    ret = TOPAY_LBFGS_CANCELED;
                    break;
                }
            }
            gnorm_inf = absmax(g);
            xnorm_inf = absmax(x);
            if (gnorm_inf / std::max(1.0, xnorm_inf) < param.g_epsilon) {
                ret = TOPAY_LBFGS_CONVERGENCE;
                break;
            }
            if (0 < param.past) {
                if (param.past <= k) {
                    rate = std::fabs(pf[k % param.past] - fx) / std::max(1.0, std::fabs(fx));
                    if (rate < param.delta) {
                        ret = TOPAY_LBFGS_STOP;
                        break;
                    }
                }
                pf[k % param.past] = fx;
            }
            if (param.max_iterations != 0 && param.max_iterations <= k) {
                ret = TOPAY_LBFGSERR_MAXIMUMITERATION;
                break;
            }
            ++k;
            double* se = &lm_s[(size_t)end * n];
            double* ye = &lm_y[(size_t)end * n];
            for (i = 0; i < n; i++) {
                se[i] = x[i] - xp[i];
                ye[i] = g[i] - gp[i];
            }
            ys = vdot(ye, se, n);
            yy = vdot(ye, ye, n);
            lm_ys[end] = ys;
            for (i = 0; i < n; i++) d[i] = -g[i];
            cau = vdot(se, se, n) * std::sqrt(vdot(gp, gp)) * param.cautious_factor;
            if (ys > cau) {
                ++bound;
                bound = m < bound ? m : bound;
                end = (end + 1) % m;
                j = end;
                for (i = 0; i < bound; ++i) {
                    j = (j + m - 1) % m;
                    lm_alpha[j] = vdot(&lm_s[(size_t)j * n], d.data(), n) / lm_ys[j];
                    const double a = -lm_alpha[j];
                    const double* yj = &lm_y[(size_t)j * n];
                    for (int t = 0; t < n; t++) d[t] += a * yj[t];
                }
                const double sc = ys / yy;
                for (int t = 0; t < n; t++) d[t] *= sc;
                for (i = 0; i < bound; ++i) {
                    beta = vdot(&lm_y[(size_t)j * n], d.data(), n) / lm_ys[j];
                    const double a = lm_alpha[j] - beta;
                    const double* sj = &lm_s[(size_t)j * n];
                    for (int t = 0; t < n; t++) d[t] += a * sj[t];
                    j = (j + 1) % m;
                }
            }
            step = 1.0;
        }
    }
    f = fx;
    if (st) st->iters += k;
    return ret;
}

// ---------------------------------------------------------------- optimizer
struct TrajOpt {
    topay_opt_params opt;
    topay_robot_params rp;
    const Field* grid = nullptr;
    const RogEsdf* rog = nullptr;   // set => GridMap's use_rog branches (grid_map.h:364-392, 443-461)

    // data (moma_traj_opt.h:626-639)
    int piece_num = 0;
    double start_state[10], end_state[10];
    Vec times;
    double minco_start_state[27], minco_end_state[27];  // 9 x 3 row-major
    Vec inner_pts;                                       // 9 x (N-1) column-major
    Vec init_inner_xy;                                   // N x 2
    MinJerk9 minco;
    double alm_lambda[2] = {0, 0}, alm_rho[2] = {1, 1};
    double final_xy_error[2] = {0, 0};
    double traj_cost = 0.0;
    double terms[TOPAY_NTERMS];
    int s1_past = 0;
    // solve statistics
    int last_code = 0, alm_rounds = 0;
    LbfgsStats stats;
    // optional L-BFGS trace (f after each accepted iteration, step, ls count)
    std::vector<double>* trace = nullptr;
    // DEBUG ONLY (never set for parity or baselines): replaces reference quirks 1 and 2 of
    // SURVEY.md §8a by the exact adjoint of the Simpson prefix, so that the oracle's own
    // gradient can be validated against finite differences of its cost.
    bool exact_chain = false;

    // ---- scalar maps, moma_traj_opt.h:744-830 ----
    static inline double expC2(double tau) {
        return tau > 0.0 ? ((0.5 * tau + 1.0) * tau + 1.0) : 1.0 / ((0.5 * tau - 1.0) * tau + 1.0);
    }
    static inline double logC2(double T) {
        return T > 1.0 ? (std::sqrt(2.0 * T - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / T - 1.0));
    }
    static inline double getTtoTauGrad(double tau) {
        if (tau > 0)
            return tau + 1.0;
        else {
            double denSqrt = (0.5 * tau - 1.0) * tau + 1.0;
            return (1.0 - tau) / (denSqrt * denSqrt);
        }
    }
    static inline double sigmoidC2(double vq, double max_q) {
        double e_ang = expC2(vq);
        return 2.0 * max_q * e_ang / (1.0 + e_ang) - max_q;
    }
    static inline double invSigmoidC2(double q, double max_q) {
        double b = 0.5 * (max_q + q) / max_q;
        return logC2(b / (1 - b));
    }
    static inline double getQtoVqGrad(double vq, double max_q) {
        double e_ang_1 = expC2(vq) + 1.0;
        return 2.0 * max_q * getTtoTauGrad(vq) / (e_ang_1 * e_ang_1);
    }
    inline void smoothL1Penalty(const double x, double& f, double& df) const {
        const double pe = opt.relu_mu;
        const double half = 0.5 * pe;
        const double f3c = 1.0 / (pe * pe);
        const double f4c = -0.5 * f3c / pe;
        const double d2c = 3.0 * f3c;
        const double d3c = 4.0 * f4c;
        if (x < pe) {
            f = (f4c * x + f3c) * x * x * x;
            df = (d3c * x + d2c) * x * x;
        } else {
            f = x - half;
            df = 1.0;
        }
    }

    // ---- trapezoid timing, moma_traj_opt.h:676-733 ----
    static double getDurationTrapezoid(double length, double startV, double endV, double maxV, double maxA) {
        double critical_len;
        double startv2 = startV * startV;
        double endv2 = endV * endV;
        double maxv2 = maxV * maxV;
        if (startV > maxV) startv2 = maxv2;
        if (endV > maxV) endv2 = maxv2;
        critical_len = (maxv2 - startv2) / (2 * maxA) + (maxv2 - endv2) / (2 * maxA);
        if (length >= critical_len)
            return (maxV - startV) / maxA + (maxV - endV) / maxA + (length - critical_len) / maxV;
        else {
            double tmpv = std::sqrt(0.5 * (startv2 + endv2 + 2 * maxA * length));
            return (tmpv - startV) / maxA + (tmpv - endV) / maxA;
        }
    }
    static double getArcTrapezoid(double curt, double locallength, double startV, double endV, double maxV,
                                  double maxA) {
        double critical_len;
        double startv2 = startV * startV;
        double endv2 = endV * endV;
        double maxv2 = maxV * maxV;
        if (startV > maxV) startv2 = maxv2;
        if (endV > maxV) endv2 = maxv2;
        critical_len = (maxv2 - startv2) / (2 * maxA) + (maxv2 - endv2) / (2 * maxA);
        if (locallength >= critical_len) {
            double t1 = (maxV - startV) / maxA;
            double t2 = t1 + (locallength - critical_len) / maxV;
            if (curt <= t1)
                return startV * curt + 0.5 * maxA * (curt * curt);
            else if (curt <= t2)
                return startV * t1 + 0.5 * maxA * (t1 * t1) + (curt - t1) * maxV;
            else
                return startV * t1 + 0.5 * maxA * (t1 * t1) + (t2 - t1) * maxV + maxV * (curt - t2) -
                       0.5 * maxA * (curt - t2) * (curt - t2);
        } else {
            double tmpv = std::sqrt(0.5 * (startv2 + endv2 + 2 * maxA * locallength));
            double tmpt = (tmpv - startV) / maxA;
            if (curt <= tmpt)
                return startV * curt + 0.5 * maxA * (curt * curt);
            else
                return startV * tmpt + 0.5 * maxA * (tmpt * tmpt) + tmpv * (curt - tmpt) -
                       0.5 * maxA * (curt - tmpt) * (curt - tmpt);
        }
    }
    static void normalizeAngle(const double ref_angle, double& angle) {
        while (ref_angle - angle > M_PI) angle += 2 * M_PI;
        while (ref_angle - angle < -M_PI) angle -= 2 * M_PI;
    }

    // moma_traj_opt.cpp:146-344: waypoints -> problem data + x0. init_path is
    // len x 10; bvel / bacc are 10 x 2 row-major. Returns x0.
    Vec prepare(const double* init_path, int len, const double* bvel, const double* bacc) {
        auto BV = [&](int r, int c) { return bvel[r * 2 + c]; };
        auto BA = [&](int r, int c) { return bacc[r * 2 + c]; };
        for (int i = 0; i < 10; i++) {
            start_state[i] = init_path[i];
            end_state[i] = init_path[(size_t)(len - 1) * 10 + i];
        }
        // sampled_path rows: x y theta delta_theta delta_arc q(7)
        std::vector<std::array<double, 12>> sampled_path;
        std::array<double, 12> st{};
        st.fill(0.0);
        for (int i = 0; i < 3; i++) st[i] = init_path[i];
        for (int i = 0; i < 7; i++) st[5 + i] = init_path[3 + i];
        sampled_path.push_back(st);
        for (int i = 1; i < len; i++) {
            const double* cur = &init_path[(size_t)i * 10];
            const double* prv = &init_path[(size_t)(i - 1) * 10];
            st.fill(0.0);
            double dx = cur[0] - prv[0], dy = cur[1] - prv[1];
            double arc_len = std::sqrt(dx * dx + dy * dy);
            double now_theta = cur[2];
            normalizeAngle(sampled_path.back()[2], now_theta);
            double theta_diff = now_theta - sampled_path.back()[2];
            if (std::fabs(theta_diff) > 1e-2) {
                if (arc_len < 1e-2) {
                    st[0] = cur[0];
                    st[1] = cur[1];
                    st[2] = now_theta;
                    st[3] = theta_diff;
                    st[4] = 0.0;
                    for (int q = 0; q < 7; q++) st[5 + q] = cur[3 + q];
                    sampled_path.push_back(st);
                } else {
                    st = sampled_path.back();
                    double direct_theta = std::atan2(cur[1] - sampled_path.back()[1], cur[0] - sampled_path.back()[0]);
                    normalizeAngle(sampled_path.back()[2], direct_theta);
                    theta_diff = direct_theta - sampled_path.back()[2];
                    st[2] = direct_theta;
                    st[3] = theta_diff;
                    st[4] = 0.0;
                    sampled_path.push_back(st);

                    st[0] = cur[0];
                    st[1] = cur[1];
                    st[2] = direct_theta;
                    st[3] = 0.0;
                    st[4] = arc_len;
                    for (int q = 0; q < 7; q++) st[5 + q] = cur[3 + q];
                    sampled_path.push_back(st);

                    normalizeAngle(sampled_path.back()[2], now_theta);
                    theta_diff = now_theta - sampled_path.back()[2];
                    st[2] = now_theta;
                    st[3] = theta_diff;
                    st[4] = 0.0;
                    sampled_path.push_back(st);
                }
            } else {
                if (arc_len > 1e-2) {
                    st[0] = cur[0];
                    st[1] = cur[1];
                    st[2] = now_theta;
                    st[3] = 0.0;
                    st[4] = arc_len;
                    for (int q = 0; q < 7; q++) st[5 + q] = cur[3 + q];
                    sampled_path.push_back(st);
                }
            }
        }

        std::vector<double> path_arcs, weighted_path_arcs;
        double total_len = 0, weighted_total_len = 0;
        const size_t path_num = sampled_path.size();
        path_arcs.push_back(0);
        weighted_path_arcs.push_back(0);
        for (size_t idx = 1; idx < path_num; idx++) {
            const auto& node = sampled_path[idx];
            total_len += node[4];
            path_arcs.push_back(total_len);
            weighted_total_len += 0.2 * std::fabs(node[3]) + 1.4 * std::fabs(node[4]);
            weighted_path_arcs.push_back(weighted_total_len);
        }
        double total_time = getDurationTrapezoid(weighted_total_len, BV(0, 0), 0.0, rp.max_v, rp.max_a);
        std::vector<std::array<double, 9>> vector_inner_pts;
        double sample_interval =
            total_time / std::max(int(total_time / opt.sample_interval + 0.5), (int)opt.min_piece_num);
        size_t now_idx = 1;
        init_inner_xy.clear();
        for (double t = sample_interval; t < total_time - 1e-3; t += sample_interval) {
            double arc = getArcTrapezoid(t, weighted_total_len, BV(0, 0), 0.0, rp.max_v, rp.max_a);
            for (size_t k = now_idx; k < path_num; k++) {
                const auto& node = sampled_path[k];
                const auto& pre = sampled_path[k - 1];
                double tmp_arc = weighted_path_arcs[k];
                if (tmp_arc >= arc) {
                    now_idx = k;
                    double l1 = tmp_arc - arc;
                    double l = weighted_path_arcs[k] - weighted_path_arcs[k - 1];
                    std::array<double, 9> pts{};
                    pts[0] = pre[2] + (l - l1) / l * (node[3]);
                    pts[1] = path_arcs[k - 1] + (l - l1) / l * (node[4]);
                    for (int q = 0; q < 7; q++) pts[2 + q] = pre[5 + q] + (l - l1) / l * (node[5 + q] - pre[5 + q]);
                    vector_inner_pts.push_back(pts);
                    double interp_x = l1 / l * pre[0] + (l - l1) / l * (node[0]);
                    double interp_y = l1 / l * pre[1] + (l - l1) / l * (node[1]);
                    init_inner_xy.push_back(interp_x);
                    init_inner_xy.push_back(interp_y);
                    break;
                }
            }
        }
        init_inner_xy.push_back(end_state[0]);
        init_inner_xy.push_back(end_state[1]);

        for (int i = 0; i < 27; i++) minco_start_state[i] = minco_end_state[i] = 0.0;
        auto SS = [&](int r, int c) -> double& { return minco_start_state[r * 3 + c]; };
        auto ES = [&](int r, int c) -> double& { return minco_end_state[r * 3 + c]; };
        SS(0, 0) = sampled_path[0][2];
        SS(0, 1) = BV(1, 0);
        SS(0, 2) = BA(1, 0);
        SS(1, 1) = BV(0, 0);
        SS(1, 2) = BA(0, 0);
        for (int q = 0; q < 7; q++) {
            SS(2 + q, 0) = sampled_path[0][5 + q];
            SS(2 + q, 1) = BV(3 + q, 0);
            SS(2 + q, 2) = BA(3 + q, 0);
        }
        ES(0, 0) = sampled_path.back()[2];
        ES(1, 0) = path_arcs.back();
        for (int q = 0; q < 7; q++) {
            ES(2 + q, 0) = sampled_path.back()[5 + q];
            ES(2 + q, 1) = BV(3 + q, 1);
            ES(2 + q, 2) = BA(3 + q, 1);
        }

        piece_num = (int)vector_inner_pts.size() + 1;
        times.assign(piece_num, sample_interval);
        inner_pts.assign((size_t)9 * (piece_num - 1), 0.0);
        for (size_t i = 0; i < vector_inner_pts.size(); i++)
            for (int d = 0; d < 9; d++) inner_pts[i * 9 + d] = vector_inner_pts[i][d];

        minco.reset(piece_num, opt.energy_weights);
        minco.generate(minco_start_state, minco_end_state, inner_pts.data(), times.data());

        const int N = piece_num;
        Vec x(topay_num_vars(N));
        double* Tau = x.data();
        double* Theta = Tau + N;
        double* Arc = Theta + (N - 1);
        double* Vq = Arc + N;  // 7 x (N-1) column-major
        for (int i = 0; i < N - 1; i++) {
            Tau[i] = logC2(times[i]);
            Theta[i] = inner_pts[(size_t)i * 9 + 0];
            Arc[i] = inner_pts[(size_t)i * 9 + 1];
            for (int j = 0; j < 7; j++)
                Vq[(size_t)i * 7 + j] = invSigmoidC2(inner_pts[(size_t)i * 9 + 2 + j], rp.joint_pos_limit_max[j]);
        }
        Tau[N - 1] = logC2(times[N - 1]);
        Arc[N - 1] = ES(1, 0);
        // moma_traj_opt.cpp:354-357
        if (std::fabs(ES(1, 0)) < opt.s1_shot_path_horizon)
            s1_past = opt.s1_lbfgs_shot_path_past;
        else
            s1_past = opt.s1_lbfgs_normal_past;
        return x;
    }

    // Install problem data directly (used by the per-evaluation parity hook).
    void set_problem(int N, const double* head, const double* tail, const double* sxy, const double* exy,
                     const double* inner_xy, const double* lambda, const double* rho) {
        piece_num = N;
        for (int i = 0; i < 27; i++) {
            minco_start_state[i] = head[i];
            minco_end_state[i] = tail[i];
        }
        for (int i = 0; i < 10; i++) start_state[i] = end_state[i] = 0.0;
        start_state[0] = sxy[0];
        start_state[1] = sxy[1];
        end_state[0] = exy[0];
        end_state[1] = exy[1];
        init_inner_xy.assign(inner_xy, inner_xy + 2 * N);
        if (lambda) {
            alm_lambda[0] = lambda[0];
            alm_lambda[1] = lambda[1];
        }
        if (rho) {
            alm_rho[0] = rho[0];
            alm_rho[1] = rho[1];
        }
        times.assign(N, 0.0);
        inner_pts.assign((size_t)9 * (N - 1), 0.0);
        minco.reset(N, opt.energy_weights);
    }

    // firstStageCostCallback / secondStageCostCallback, moma_traj_opt.cpp:817-955
    double cost_callback(int stage, const Vec& x, Vec& grad) {
        const int N = piece_num;
        const double* Tau = x.data();
        const double* Theta = Tau + N;
        const double* Arc = Theta + (N - 1);
        const double* Vq = Arc + N;
        double* gradTau = grad.data();
        double* gradTheta = gradTau + N;
        double* gradArc = gradTheta + (N - 1);
        double* gradVq = gradArc + N;
        for (int i = 0; i < TOPAY_NTERMS; i++) terms[i] = 0.0;

        times.resize(N);
        for (int i = 0; i < N; i++) times[i] = expC2(Tau[i]);
        minco_end_state[1 * 3 + 0] = Arc[N - 1];
        inner_pts.resize((size_t)9 * (N - 1));
        for (int i = 0; i < N - 1; i++) {
            inner_pts[(size_t)i * 9 + 0] = Theta[i];
            inner_pts[(size_t)i * 9 + 1] = Arc[i];
            for (int j = 0; j < 7; j++)
                inner_pts[(size_t)i * 9 + 2 + j] = sigmoidC2(Vq[(size_t)i * 7 + j], rp.joint_pos_limit_max[j]);
        }
        minco.generate(minco_start_state, minco_end_state, inner_pts.data(), times.data());

        Vec gdC_jerk, gdT_jerk;
        minco.calJerkGradCT(gdC_jerk, gdT_jerk);
        double jerk_cost = minco.getTrajJerkCost();

        double penalty_cost = 0.0;
        Vec gdC_pen, gdT_pen;
        if (stage == 1)
            first_stage_penalty(penalty_cost, gdC_pen, gdT_pen);
        else
            second_stage_penalty(penalty_cost, gdC_pen, gdT_pen);

        Vec gdC(gdC_jerk.size()), gdT(N);
        for (size_t i = 0; i < gdC.size(); i++) gdC[i] = gdC_jerk[i] + gdC_pen[i];
        for (int i = 0; i < N; i++) gdT[i] = gdT_jerk[i] + gdT_pen[i];
        Vec gdP, gdP_tail;
        minco.calGradCTtoQT(gdC, gdT, gdP, gdP_tail);

        const double tw = stage == 1 ? opt.s1_time_weight : opt.s2_time_weight;
        double tsum = 0.0;
        for (int i = 0; i < N; i++) tsum += times[i];
        double time_cost = tw * tsum;

        for (int i = 0; i < N - 1; i++) {
            gradTheta[i] = gdP[(size_t)i * 9 + 0];
            gradArc[i] = gdP[(size_t)i * 9 + 1];
            for (int j = 0; j < 7; j++)
                gradVq[(size_t)i * 7 + j] =
                    gdP[(size_t)i * 9 + 2 + j] * getQtoVqGrad(Vq[(size_t)i * 7 + j], rp.joint_pos_limit_max[j]);
        }
        for (int i = 0; i < N; i++) gradTau[i] = (gdT[i] + tw) * getTtoTauGrad(Tau[i]);
        gradArc[N - 1] = gdP_tail[1 * 3 + 0];
        terms[TOPAY_TERM_JERK] += jerk_cost;
        terms[TOPAY_TERM_TIME] += time_cost;
        return jerk_cost + penalty_cost + time_cost;
    }

    // Basis rows at local time s1 (moma_traj_opt.cpp:1263-1270)
    static inline void basis(double s1, double* b0, double* b1, double* b2, double* b3) {
        double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s3 * s2;
        b0[0] = 1.0; b0[1] = s1; b0[2] = s2; b0[3] = s3; b0[4] = s4; b0[5] = s5;
        b1[0] = 0.0; b1[1] = 1.0; b1[2] = 2.0 * s1; b1[3] = 3.0 * s2; b1[4] = 4.0 * s3; b1[5] = 5.0 * s4;
        b2[0] = 0.0; b2[1] = 0.0; b2[2] = 2.0; b2[3] = 6.0 * s1; b2[4] = 12.0 * s2; b2[5] = 20.0 * s3;
        if (b3) {
            b3[0] = 0.0; b3[1] = 0.0; b3[2] = 0.0; b3[3] = 6.0; b3[4] = 24.0 * s1; b3[5] = 60.0 * s2;
        }
    }
    // c.transpose() * beta for piece i: out[d] = sum_k c(6i+k, d) * beta[k]
    inline void ct_beta(int i, const double* beta, double* out) const {
        for (int d = 0; d < 9; d++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s += minco.C(6 * i + k, d) * beta[k];
            out[d] = s;
        }
    }

    struct ChainStore {
        // per piece: X/Y GradCTheta, GradCArc (6 x (2K+1), column j at [j*6]) and GradT (2K+1)
        Vec XgCTheta, XgCArc, XgT, YgCTheta, YgCArc, YgT;
    };

    // Midpoint / Jacobian part shared by both stages (all j): fills columns j of the
    // Single* arrays. moma_traj_opt.cpp:1293-1300 and :1734-1740.
    inline void jacobians(int j, double alpha, double coeff, int int_6K, const double* beta0, const double* beta1,
                          const double* dstate, const double* d2state, double cyaw, double syaw, ChainStore& cs) const {
        for (int k = 0; k < 6; k++) {
            cs.XgCTheta[(size_t)j * 6 + k] = -dstate[1] * beta0[k] * syaw;
            cs.XgCArc[(size_t)j * 6 + k] = beta1[k] * cyaw;
            cs.YgCTheta[(size_t)j * 6 + k] = dstate[1] * beta0[k] * cyaw;
            cs.YgCArc[(size_t)j * 6 + k] = beta1[k] * syaw;
        }
        cs.XgT[j] = (d2state[1] * cyaw - dstate[1] * dstate[0] * syaw) * alpha * coeff + dstate[1] * cyaw / int_6K;
        cs.YgT[j] = (d2state[1] * syaw + dstate[1] * dstate[0] * cyaw) * alpha * coeff + dstate[1] * syaw / int_6K;
    }

    // moma_traj_opt.cpp:957-1198
    void first_stage_penalty(double& cost, Vec& gdC, Vec& gdT) {
        const int N = piece_num, K = opt.int_K;
        cost = 0.0;
        gdC.assign((size_t)6 * N * 9, 0.0);
        gdT.assign(N, 0.0);
        double beta0[6], beta1[6], beta2[6], beta3[6];
        double state[9], dstate[9], d2state[9], d3state[9];
        double alpha, omg;
        const int inner_num = 2 * K;
        const int int_6K = K * 6;
        double violaMom, violaMomPena, violaMomPenaD;
        Vec IntegralChainCoeff(inner_num + 1, 0.0);
        for (int i = 0; i < K; i++) {
            IntegralChainCoeff[2 * i] += 1.0;
            IntegralChainCoeff[2 * i + 1] += 4.0;
            IntegralChainCoeff[2 * i + 2] += 1.0;
        }
        std::vector<ChainStore> store(N);
        double CurrentXY[2] = {start_state[0], start_state[1]};
        const size_t S = (size_t)N * (inner_num + 1);
        Vec ChainX(S, 0.0), ChainY(S, 0.0);
        Vec FinalXY(2 * (N + 1));
        FinalXY[0] = CurrentXY[0];
        FinalXY[1] = CurrentXY[1];
        double cost_path = 0, cost_moment = 0, cost_acc = 0, cost_domega = 0;
        const double w_m = opt.s1_moment_weight, w_a = opt.s1_acc_weight, w_dw = opt.s1_domega_weight;

        for (int i = 0; i < N; i++) {
            double step = times[i] / K;
            double half_step = step / 2.0;
            double coeff = step / 6.0;
            ChainStore& cs = store[i];
            cs.XgCTheta.assign((size_t)6 * (inner_num + 1), 0.0);
            cs.XgCArc = cs.YgCTheta = cs.YgCArc = cs.XgCTheta;
            cs.XgT.assign(inner_num + 1, 0.0);
            cs.YgT = cs.XgT;
            Vec IntegralX(K, 0.0), IntegralY(K, 0.0);
            double s1 = 0.0;
            for (int j = 0; j <= inner_num; j++) {
                if (j % 2 == 0) {
                    basis(s1, beta0, beta1, beta2, beta3);
                    s1 += half_step;
                    alpha = 1.0 / inner_num * j;
                    omg = (j == 0 || j == inner_num) ? 0.5 : 1.0;
                    ct_beta(i, beta0, state);
                    ct_beta(i, beta1, dstate);
                    ct_beta(i, beta2, d2state);
                    ct_beta(i, beta3, d3state);
                    double syaw = std::sin(state[0]);
                    double cyaw = std::cos(state[0]);
                    if (j != 0) {
                        IntegralX[j / 2 - 1] += coeff * dstate[1] * cyaw;
                        IntegralY[j / 2 - 1] += coeff * dstate[1] * syaw;
                    }
                    if (j != inner_num) {
                        IntegralX[j / 2] += coeff * dstate[1] * cyaw;
                        IntegralY[j / 2] += coeff * dstate[1] * syaw;
                    }
                    jacobians(j, alpha, coeff, int_6K, beta0, beta1, dstate, d2state, cyaw, syaw, cs);
                    if (j != 0) {
                        CurrentXY[0] += IntegralX[j / 2 - 1];
                        CurrentXY[1] += IntegralY[j / 2 - 1];
                    }
                    double gradViolaMt;
                    double real_alpha = 1.0 / K * ((double)j / 2.0);
                    double gradBeta[3][2] = {{0, 0}, {0, 0}, {0, 0}};
                    for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
                        violaMom = omg_sym * rp.max_v * dstate[0] + rp.max_w * dstate[1] - rp.max_v * rp.max_w;
                        if (violaMom > 0) {
                            smoothL1Penalty(violaMom, violaMomPena, violaMomPenaD);
                            gradViolaMt = real_alpha * (omg_sym * rp.max_v * d2state[0] + rp.max_w * d2state[1]);
                            gradBeta[1][0] += omg * step * w_m * violaMomPenaD * omg_sym * rp.max_v;
                            gradBeta[1][1] += omg * step * w_m * violaMomPenaD * rp.max_w;
                            gdT[i] += omg * w_m * (violaMomPenaD * gradViolaMt * step + violaMomPena / K);
                            cost += omg * step * w_m * violaMomPena;
                            cost_moment += omg * step * w_m * violaMomPena;
                        }
                    }
                    for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
                        violaMom = omg_sym * rp.max_v * dstate[0] - rp.max_w * dstate[1] - rp.max_v * rp.max_w;
                        if (violaMom > 0) {
                            smoothL1Penalty(violaMom, violaMomPena, violaMomPenaD);
                            gradViolaMt = real_alpha * (omg_sym * rp.max_v * d2state[0] - rp.max_w * d2state[1]);
                            gradBeta[1][0] += omg * step * w_m * violaMomPenaD * omg_sym * rp.max_v;
                            gradBeta[1][1] -= omg * step * w_m * violaMomPenaD * rp.max_w;
                            gdT[i] += omg * w_m * (violaMomPenaD * gradViolaMt * step + violaMomPena / K);
                            cost += omg * step * w_m * violaMomPena;
                            cost_moment += omg * step * w_m * violaMomPena;
                        }
                    }
                    double violaAcc = d2state[1] * d2state[1] - rp.max_a * rp.max_a;
                    double violaAlp = d2state[0] * d2state[0] - rp.max_dw * rp.max_dw;
                    double violaAccPena, violaAccPenaD, violaAlpPena, violaAlpPenaD;
                    if (violaAcc > 0) {
                        smoothL1Penalty(violaAcc, violaAccPena, violaAccPenaD);
                        double gradViolaAT = 2.0 * real_alpha * d2state[1] * d3state[1];
                        gradBeta[2][1] += omg * step * w_a * violaAccPenaD * 2.0 * d2state[1];
                        gdT[i] += omg * w_a * (violaAccPenaD * gradViolaAT * step + violaAccPena / K);
                        cost += omg * step * w_a * violaAccPena;
                        cost_acc += omg * step * w_a * violaAccPena;
                    }
                    if (violaAlp > 0) {
                        smoothL1Penalty(violaAlp, violaAlpPena, violaAlpPenaD);
                        double gradViolaDOT = 2.0 * real_alpha * d2state[0] * d3state[0];
                        gradBeta[2][0] += omg * step * w_dw * violaAlpPenaD * 2.0 * d2state[0];
                        gdT[i] += omg * w_dw * (violaAlpPenaD * gradViolaDOT * step + violaAlpPena / K);
                        cost += omg * step * w_dw * violaAlpPena;
                        cost_domega += omg * step * w_dw * violaAlpPena;
                    }
                    for (int k = 0; k < 6; k++)
                        for (int d = 0; d < 2; d++)
                            gdC[(size_t)(i * 6 + k) * 9 + d] +=
                                beta0[k] * gradBeta[0][d] + beta1[k] * gradBeta[1][d] + beta2[k] * gradBeta[2][d];
                } else {
                    basis(s1, beta0, beta1, beta2, nullptr);
                    s1 += half_step;
                    alpha = 1.0 / inner_num * j;
                    ct_beta(i, beta0, state);
                    ct_beta(i, beta1, dstate);
                    ct_beta(i, beta2, d2state);
                    double cyaw = std::cos(state[0]), syaw = std::sin(state[0]);
                    IntegralX[j / 2] += 4 * coeff * dstate[1] * cyaw;
                    IntegralY[j / 2] += 4 * coeff * dstate[1] * syaw;
                    jacobians(j, alpha, coeff, int_6K, beta0, beta1, dstate, d2state, cyaw, syaw, cs);
                }
            }
            for (auto* v : {&cs.XgCArc, &cs.XgCTheta, &cs.YgCArc, &cs.YgCTheta})
                for (double& e : *v) e = e * coeff;
            double sx = 0.0, sy = 0.0;
            for (int k = 0; k < K; k++) {
                sx += IntegralX[k];
                sy += IntegralY[k];
            }
            FinalXY[2 * (i + 1)] = FinalXY[2 * i] + sx;
            FinalXY[2 * (i + 1) + 1] = FinalXY[2 * i + 1] + sy;
            const double ex = FinalXY[2 * (i + 1)] - init_inner_xy[2 * i];
            const double ey = FinalXY[2 * (i + 1) + 1] - init_inner_xy[2 * i + 1];
            double violaPos = ex * ex + ey * ey;
            // head(i*(inner_num+1)): pieces 0..i-1 only (reference quirk 1)
            const size_t hn = (size_t)(exact_chain ? i + 1 : i) * (inner_num + 1);
            const double ax = opt.s1_path_pos_weight * 2.0 * ex, ay = opt.s1_path_pos_weight * 2.0 * ey;
            for (size_t t = 0; t < hn; t++) {
                ChainX[t] += ax;
                ChainY[t] += ay;
            }
            cost += opt.s1_path_pos_weight * violaPos;
            cost_path += opt.s1_path_pos_weight * violaPos;
        }
        chain_contract(store, ChainX, ChainY, IntegralChainCoeff, gdC, gdT);
        final_xy_error[0] = FinalXY[2 * N] - end_state[0];
        final_xy_error[1] = FinalXY[2 * N + 1] - end_state[1];
        terms[TOPAY_TERM_MOMENT] += cost_moment;
        terms[TOPAY_TERM_ACC] += cost_acc;
        terms[TOPAY_TERM_DOMEGA] += cost_domega;
        terms[TOPAY_TERM_ENDP] += cost_path;
    }

    // exact_chain only: the node's own Simpson weight is 1 at the end of its last interval (0 for
    // j == 0), not the fixed pattern value the reference applies.
    static inline void exact_fix(Vec& CX, Vec& CY, size_t own, int j, int inner_num, double gx, double gy) {
        if (j == 0) {
            CX[own] -= gx;
            CY[own] -= gy;
        } else if (j != inner_num) {
            CX[own] -= 0.5 * gx;
            CY[own] -= 0.5 * gy;
        }
    }

    // moma_traj_opt.cpp:1812-1822 (and :1181-1191)
    void chain_contract(const std::vector<ChainStore>& store, const Vec& ChainX, const Vec& ChainY,
                        const Vec& ICC, Vec& gdC, Vec& gdT) const {
        const int N = piece_num, L = 2 * opt.int_K + 1;
        Vec CoeffX(L), CoeffY(L);
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < L; j++) {
                CoeffX[j] = ChainX[(size_t)i * L + j] * ICC[j];
                CoeffY[j] = ChainY[(size_t)i * L + j] * ICC[j];
            }
            const ChainStore& cs = store[i];
            auto matvec = [&](const Vec& M, const Vec& v, int col) {
                for (int k = 0; k < 6; k++) {
                    double s = 0.0;
                    for (int j = 0; j < L; j++) s += M[(size_t)j * 6 + k] * v[j];
                    gdC[(size_t)(i * 6 + k) * 9 + col] += s;
                }
            };
            matvec(cs.XgCArc, CoeffX, 1);
            matvec(cs.XgCTheta, CoeffX, 0);
            matvec(cs.YgCArc, CoeffY, 1);
            matvec(cs.YgCTheta, CoeffY, 0);
            double sx = 0.0, sy = 0.0;
            for (int j = 0; j < L; j++) sx += cs.XgT[j] * CoeffX[j];
            for (int j = 0; j < L; j++) sy += cs.YgT[j] * CoeffY[j];
            gdT[i] += sx;
            gdT[i] += sy;
        }
    }

    // moma_traj_opt.cpp:1200-1829
    void second_stage_penalty(double& cost, Vec& gdC, Vec& gdT) {
        const int N = piece_num, K = opt.int_K;
        cost = 0.0;
        gdC.assign((size_t)6 * N * 9, 0.0);
        gdT.assign(N, 0.0);
        double beta0[6], beta1[6], beta2[6], beta3[6];
        double state[9], dstate[9], d2state[9], d3state[9];
        double alpha, omg;
        const int inner_num = 2 * K;
        const int int_6K = K * 6;
        double avg_time = 0.0;
        for (int i = 0; i < N; i++) avg_time += times[i];
        avg_time /= N;
        double violaPos, violaPosPena, violaPosPenaD;
        double violaMom, violaMomPena, violaMomPenaD;
        Vec IntegralChainCoeff(inner_num + 1, 0.0);
        for (int i = 0; i < K; i++) {
            IntegralChainCoeff[2 * i] += 1.0;
            IntegralChainCoeff[2 * i + 1] += 4.0;
            IntegralChainCoeff[2 * i + 2] += 1.0;
        }
        std::vector<ChainStore> store(N);
        double CurrentXY[2] = {start_state[0], start_state[1]};
        const size_t S = (size_t)N * (inner_num + 1);
        Vec ChainX(S, 0.0), ChainY(S, 0.0);
        Vec FinalXY(2 * (N + 1));
        FinalXY[0] = CurrentXY[0];
        FinalXY[1] = CurrentXY[1];
        const double w_col = opt.s2_collision_weight, w_m = opt.s2_moment_weight, w_a = opt.s2_acc_weight,
                     w_dw = opt.s2_domega_weight, w_mc = opt.s2_mani_colli_weight, w_sc = opt.s2_self_colli_weight,
                     w_mp = opt.s2_mani_pos_weight, w_mv = opt.s2_mani_vel_weight, w_ma = opt.s2_mani_acc_weight,
                     w_mt = opt.s2_mean_time_weight;
        double* tm = terms;

        for (int i = 0; i < N; i++) {
            double step = times[i] / K;
            double half_step = step / 2.0;
            double coeff = step / 6.0;
            ChainStore& cs = store[i];
            cs.XgCTheta.assign((size_t)6 * (inner_num + 1), 0.0);
            cs.XgCArc = cs.YgCTheta = cs.YgCArc = cs.XgCTheta;
            cs.XgT.assign(inner_num + 1, 0.0);
            cs.YgT = cs.XgT;
            Vec IntegralX(K, 0.0), IntegralY(K, 0.0);
            double s1 = 0.0;
            for (int j = 0; j <= inner_num; j++) {
                if (j % 2 == 0) {
                    basis(s1, beta0, beta1, beta2, beta3);
                    s1 += half_step;
                    alpha = 1.0 / inner_num * j;
                    omg = (j == 0 || j == inner_num) ? 0.5 : 1.0;
                    ct_beta(i, beta0, state);
                    ct_beta(i, beta1, dstate);
                    ct_beta(i, beta2, d2state);
                    ct_beta(i, beta3, d3state);
                    double syaw = std::sin(state[0]);
                    double cyaw = std::cos(state[0]);
                    if (j != 0) {
                        IntegralX[j / 2 - 1] += coeff * dstate[1] * cyaw;
                        IntegralY[j / 2 - 1] += coeff * dstate[1] * syaw;
                    }
                    if (j != inner_num) {
                        IntegralX[j / 2] += coeff * dstate[1] * cyaw;
                        IntegralY[j / 2] += coeff * dstate[1] * syaw;
                    }
                    jacobians(j, alpha, coeff, int_6K, beta0, beta1, dstate, d2state, cyaw, syaw, cs);
                    if (j != 0) {
                        CurrentXY[0] += IntegralX[j / 2 - 1];
                        CurrentXY[1] += IntegralY[j / 2 - 1];
                    }
                    const size_t head_n = (size_t)i * (inner_num + 1) + j + 1;

                    // chassis collision (:1304-1332)
                    double sdf_value;
                    double grad_sdf[2];
                    if (rog) {   // getValueGrad2d, no inflation (inflate = critical = false)
                        const double p3[3] = {CurrentXY[0], CurrentXY[1], 0.0};
                        double g3[3];
                        rog->value_grad_2d(p3, false, sdf_value, g3);
                        grad_sdf[0] = g3[0];
                        grad_sdf[1] = g3[1];
                    } else {
                        grid->dis_with_grad_2d(CurrentXY, sdf_value, grad_sdf);
                    }
                    violaPos = rp.chassis_colli_radius * 1.05 - sdf_value;
                    if (violaPos > 0) {
                        smoothL1Penalty(violaPos, violaPosPena, violaPosPenaD);
                        const double gx = -omg * step * w_col * violaPosPenaD * grad_sdf[0];
                        const double gy = -omg * step * w_col * violaPosPenaD * grad_sdf[1];
                        for (size_t t = 0; t < head_n; t++) ChainX[t] += gx;
                        for (size_t t = 0; t < head_n; t++) ChainY[t] += gy;
                        if (exact_chain) exact_fix(ChainX, ChainY, head_n - 1, j, inner_num, gx, gy);
                        gdT[i] += omg * w_col * (violaPosPena / K);
                        double chassis_colli = omg * step * w_col * violaPosPena;
                        cost += chassis_colli;
                        tm[TOPAY_TERM_CHASSIS_COLLI] += chassis_colli;
                    }

                    // moment (:1334-1397)
                    double gradViolaMt;
                    double real_alpha = 1.0 / K * ((double)j / 2.0);
                    double gradBeta[3][2] = {{0, 0}, {0, 0}, {0, 0}};
                    for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
                        violaMom = omg_sym * rp.max_v * dstate[0] + rp.max_w * dstate[1] - rp.max_v * rp.max_w;
                        if (violaMom > 0) {
                            smoothL1Penalty(violaMom, violaMomPena, violaMomPenaD);
                            gradViolaMt = real_alpha * (omg_sym * rp.max_v * d2state[0] + rp.max_w * d2state[1]);
                            gradBeta[1][0] += omg * step * w_m * violaMomPenaD * omg_sym * rp.max_v;
                            gradBeta[1][1] += omg * step * w_m * violaMomPenaD * rp.max_w;
                            gdT[i] += omg * w_m * (violaMomPenaD * gradViolaMt * step + violaMomPena / K);
                            double moment = omg * step * w_m * violaMomPena;
                            cost += moment;
                            tm[TOPAY_TERM_MOMENT] += moment;
                        }
                    }
                    for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
                        violaMom = omg_sym * rp.max_v * dstate[0] - rp.max_w * dstate[1] - rp.max_v * rp.max_w;
                        if (violaMom > 0) {
                            smoothL1Penalty(violaMom, violaMomPena, violaMomPenaD);
                            gradViolaMt = real_alpha * (omg_sym * rp.max_v * d2state[0] - rp.max_w * d2state[1]);
                            gradBeta[1][0] += omg * step * w_m * violaMomPenaD * omg_sym * rp.max_v;
                            gradBeta[1][1] -= omg * step * w_m * violaMomPenaD * rp.max_w;
                            gdT[i] += omg * w_m * (violaMomPenaD * gradViolaMt * step + violaMomPena / K);
                            double moment = omg * step * w_m * violaMomPena;
                            cost += moment;
                            tm[TOPAY_TERM_MOMENT] += moment;
                        }
                    }

                    // acc / domega (:1413-1462)
                    double violaAcc = d2state[1] * d2state[1] - rp.max_a * rp.max_a;
                    double violaAlp = d2state[0] * d2state[0] - rp.max_dw * rp.max_dw;
                    double violaAccPena, violaAccPenaD, violaAlpPena, violaAlpPenaD;
                    if (violaAcc > 0) {
                        smoothL1Penalty(violaAcc, violaAccPena, violaAccPenaD);
                        double gradViolaAT = 2.0 * real_alpha * d2state[1] * d3state[1];
                        gradBeta[2][1] += omg * step * w_a * violaAccPenaD * 2.0 * d2state[1];
                        gdT[i] += omg * w_a * (violaAccPenaD * gradViolaAT * step + violaAccPena / K);
                        double cost_acc = omg * step * w_a * violaAccPena;
                        cost += cost_acc;
                        tm[TOPAY_TERM_ACC] += cost_acc;
                    }
                    if (violaAlp > 0) {
                        smoothL1Penalty(violaAlp, violaAlpPena, violaAlpPenaD);
                        double gradViolaDOT = 2.0 * real_alpha * d2state[0] * d3state[0];
                        gradBeta[2][0] += omg * step * w_dw * violaAlpPenaD * 2.0 * d2state[0];
                        gdT[i] += omg * w_dw * (violaAlpPenaD * gradViolaDOT * step + violaAlpPena / K);
                        double cost_domega = omg * step * w_dw * violaAlpPena;
                        cost += cost_domega;
                        tm[TOPAY_TERM_DOMEGA] += cost_domega;
                    }
                    for (int k = 0; k < 6; k++)
                        for (int d = 0; d < 2; d++)
                            gdC[(size_t)(i * 6 + k) * 9 + d] +=
                                beta0[k] * gradBeta[0][d] + beta1[k] * gradBeta[1][d] + beta2[k] * gradBeta[2][d];

                    // manipulator (:1467-1713)
                    double gradBetaQ[3][7] = {};
                    double moma_pos[10];
                    moma_pos[0] = CurrentXY[0];
                    moma_pos[1] = CurrentXY[1];
                    moma_pos[2] = state[0];
                    for (int q = 0; q < 7; q++) moma_pos[3 + q] = state[2 + q];
                    double colli_pts[TOPAY_NSPHERE][4];
                    const int ncp = get_colli_pts(rp, moma_pos, colli_pts);
                    double pos_grads[TOPAY_NSPHERE][3];
                    const double cost_scale = 10.0;
                    for (int cidx = 0; cidx < ncp; cidx++) {
                        double grad_pc[3];
                        if (rog) rog->value_grad(colli_pts[cidx], sdf_value, grad_pc);   // evaluateEDT + evaluateFirstGrad
                        else grid->dis_with_grad_3d(colli_pts[cidx], sdf_value, grad_pc);
                        violaPos = colli_pts[cidx][3] * cost_scale * 1.1 - sdf_value * cost_scale;
                        double g2p[3] = {0, 0, 0};
                        if (violaPos > 0) {
                            smoothL1Penalty(violaPos, violaPosPena, violaPosPenaD);
                            for (int d = 0; d < 3; d++)
                                g2p[d] = -omg * step * w_mc * violaPosPenaD * grad_pc[d] * cost_scale;
                            gdT[i] += omg * w_mc * (violaPosPena / K);
                            double mani_colli = omg * step * w_mc * violaPosPena;
                            cost += mani_colli;
                            tm[TOPAY_TERM_MANI_COLLI] += mani_colli;
                        }
                        for (int d = 0; d < 3; d++) pos_grads[cidx][d] = g2p[d];
                    }
                    for (int cidx = 0; cidx < ncp; cidx++) {
                        if (cidx > 2) {
                            double height = rp.chassis_height + rp.relative_t[2] + colli_pts[cidx][3] - colli_pts[cidx][2];
                            if (height > 0) {
                                double violaSelfPena, violaSelfPenaD;
                                smoothL1Penalty(height, violaSelfPena, violaSelfPenaD);
                                double grad_z = -omg * step * w_sc * violaSelfPenaD;
                                gdT[i] += omg * w_sc * (violaSelfPena / K);
                                double self_colli = omg * step * w_sc * violaSelfPena;
                                cost += self_colli;
                                tm[TOPAY_TERM_SELF_COLLI] += self_colli;
                                pos_grads[cidx][2] += grad_z;
                            }
                        }
                        for (int cj = cidx + 1; cj < ncp; cj++) {
                            if (rp.collision_matrix[cidx * TOPAY_NSPHERE + cj] != -1) continue;
                            double diff[3] = {colli_pts[cidx][0] - colli_pts[cj][0], colli_pts[cidx][1] - colli_pts[cj][1],
                                              colli_pts[cidx][2] - colli_pts[cj][2]};
                            double sqn = diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2];
                            double dist = (colli_pts[cidx][3] + colli_pts[cj][3]) * (colli_pts[cidx][3] + colli_pts[cj][3]) - sqn;
                            if (dist > 0) {
                                double violaSelfPena, violaSelfPenaD;
                                smoothL1Penalty(dist, violaSelfPena, violaSelfPenaD);
                                double grad1[3];
                                for (int d = 0; d < 3; d++) grad1[d] = -omg * step * w_sc * violaSelfPenaD * diff[d] * 2.0;
                                gdT[i] += omg * w_sc * (violaSelfPena / K);
                                double self_colli = omg * step * w_sc * violaSelfPena;
                                cost += self_colli;
                                tm[TOPAY_TERM_SELF_COLLI] += self_colli;
                                for (int d = 0; d < 3; d++) {
                                    pos_grads[cidx][d] += grad1[d];
                                    pos_grads[cj][d] -= grad1[d];
                                }
                            }
                        }
                    }
                    double moma_grad[10];
                    get_colli_grads(rp, moma_pos, pos_grads, moma_grad);

                    // joint position limits (:1616-1666)
                    for (int ji = 0; ji < 7; ji++) {
                        double violaJointPos = moma_pos[ji + 3] - rp.joint_pos_limit_max[ji];
                        double violaJointPosPena, violaJointPosPenaD;
                        if (violaJointPos > 0) {
                            smoothL1Penalty(violaJointPos, violaJointPosPena, violaJointPosPenaD);
                            moma_grad[ji + 3] += omg * step * w_mp * violaJointPosPenaD;
                            gdT[i] += omg * w_mp * (violaJointPosPena / K);
                            double c_mp = omg * step * w_mp * violaJointPosPena;
                            cost += c_mp;
                            tm[TOPAY_TERM_MANI_POS] += c_mp;
                        }
                        violaJointPos = -rp.joint_pos_limit_max[ji] - moma_pos[ji + 3];
                        if (violaJointPos > 0) {
                            smoothL1Penalty(violaJointPos, violaJointPosPena, violaJointPosPenaD);
                            moma_grad[ji + 3] -= omg * step * w_mp * violaJointPosPenaD;
                            gdT[i] += omg * w_mp * (violaJointPosPena / K);
                            double c_mp = omg * step * w_mp * violaJointPosPena;
                            cost += c_mp;
                            tm[TOPAY_TERM_MANI_POS] += c_mp;
                        }
                    }
                    // unconditional chain add (:1667-1668)
                    for (size_t t = 0; t < head_n; t++) ChainX[t] += moma_grad[0];
                    for (size_t t = 0; t < head_n; t++) ChainY[t] += moma_grad[1];
                    if (exact_chain) exact_fix(ChainX, ChainY, head_n - 1, j, inner_num, moma_grad[0], moma_grad[1]);
                    for (int k = 0; k < 6; k++) gdC[(size_t)(i * 6 + k) * 9 + 0] += beta0[k] * moma_grad[2];
                    gdT[i] += moma_grad[2] * dstate[0] * real_alpha;
                    for (int q = 0; q < 7; q++) gradBetaQ[0][q] = moma_grad[3 + q];
                    {
                        double dsum = 0.0;
                        for (int q = 0; q < 7; q++) dsum += moma_grad[3 + q] * dstate[2 + q];
                        gdT[i] += dsum * real_alpha;
                    }
                    // joint velocity / acceleration limits (:1674-1710)
                    for (int jidx = 0; jidx < 7; jidx++) {
                        const double dq = dstate[2 + jidx], d2q = d2state[2 + jidx], d3q = d3state[2 + jidx];
                        const double violaDq = dq * dq - rp.joint_vel_limit[jidx] * rp.joint_vel_limit[jidx];
                        const double violaD2q = d2q * d2q - rp.joint_acc_limit[jidx] * rp.joint_acc_limit[jidx];
                        if (violaDq > 0) {
                            double violaDqPena, violaDqPenaD;
                            smoothL1Penalty(violaDq, violaDqPena, violaDqPenaD);
                            double gradViolaVT = 2.0 * real_alpha * dq * d2q;
                            gradBetaQ[1][jidx] += omg * step * w_mv * violaDqPenaD * 2.0 * dq;
                            gdT[i] += omg * w_mv * (violaDqPenaD * gradViolaVT * step + violaDqPena / K);
                            double c_mv = omg * step * w_mv * violaDqPena;
                            cost += c_mv;
                            tm[TOPAY_TERM_MANI_VEL] += c_mv;
                        }
                        if (violaD2q > 0) {
                            double violaD2qPena, violaD2qPenaD;
                            smoothL1Penalty(violaD2q, violaD2qPena, violaD2qPenaD);
                            double gradViolaAT = 2.0 * real_alpha * d2q * d3q;
                            gradBetaQ[2][jidx] += omg * step * w_ma * violaD2qPenaD * 2.0 * d2q;
                            gdT[i] += omg * w_ma * (violaD2qPenaD * gradViolaAT * step + violaD2qPena / K);
                            double c_ma = omg * step * w_ma * violaD2qPena;
                            cost += c_ma;
                            tm[TOPAY_TERM_MANI_ACC] += c_ma;
                        }
                    }
                    for (int k = 0; k < 6; k++)
                        for (int q = 0; q < 7; q++)
                            gdC[(size_t)(i * 6 + k) * 9 + 2 + q] +=
                                beta0[k] * gradBetaQ[0][q] + beta1[k] * gradBetaQ[1][q] + beta2[k] * gradBetaQ[2][q];
                } else {
                    basis(s1, beta0, beta1, beta2, nullptr);
                    s1 += half_step;
                    alpha = 1.0 / inner_num * j;
                    ct_beta(i, beta0, state);
                    ct_beta(i, beta1, dstate);
                    ct_beta(i, beta2, d2state);
                    double cyaw = std::cos(state[0]), syaw = std::sin(state[0]);
                    IntegralX[j / 2] += 4 * coeff * dstate[1] * cyaw;
                    IntegralY[j / 2] += 4 * coeff * dstate[1] * syaw;
                    jacobians(j, alpha, coeff, int_6K, beta0, beta1, dstate, d2state, cyaw, syaw, cs);
                }
            }
            for (auto* v : {&cs.XgCArc, &cs.XgCTheta, &cs.YgCArc, &cs.YgCTheta})
                for (double& e : *v) e = e * coeff;
            double sx = 0.0, sy = 0.0;
            for (int k = 0; k < K; k++) {
                sx += IntegralX[k];
                sy += IntegralY[k];
            }
            FinalXY[2 * (i + 1)] = FinalXY[2 * i] + sx;
            FinalXY[2 * (i + 1) + 1] = FinalXY[2 * i + 1] + sy;

            // mean-time penalty with the hard-coded bounds (:1752-1769, quirk 4)
            double mean_time_lowb = 0.5;
            double mean_time_uppb = 2.0;
            if (times[i] < avg_time * mean_time_lowb) {
                double cost_meant = w_mt * (times[i] - avg_time * mean_time_lowb) * (times[i] - avg_time * mean_time_lowb);
                cost += cost_meant;
                tm[TOPAY_TERM_MEAN_TIME] += cost_meant;
                const double a = w_mt * 2.0 * (times[i] - avg_time * mean_time_lowb) * (-mean_time_lowb / N);
                for (int t = 0; t < N; t++) gdT[t] += a;
                gdT[i] += w_mt * 2.0 * (times[i] - avg_time * mean_time_lowb);
            }
            if (times[i] > avg_time * mean_time_uppb) {
                double cost_meant = w_mt * (times[i] - avg_time * mean_time_uppb) * (times[i] - avg_time * mean_time_uppb);
                cost += cost_meant;
                tm[TOPAY_TERM_MEAN_TIME] += cost_meant;
                const double a = w_mt * 2.0 * (times[i] - avg_time * mean_time_uppb) * (-mean_time_uppb / N);
                for (int t = 0; t < N; t++) gdT[t] += a;
                gdT[i] += w_mt * 2.0 * (times[i] - avg_time * mean_time_uppb);
            }
        }

        // final xy + ALM term (:1785-1810)
        final_xy_error[0] = FinalXY[2 * N] - end_state[0];
        final_xy_error[1] = FinalXY[2 * N + 1] - end_state[1];
        double cost_endp = 0.5 * (alm_rho[0] * std::pow(final_xy_error[0] + alm_lambda[0] / alm_rho[0], 2) +
                                  alm_rho[1] * std::pow(final_xy_error[1] + alm_lambda[1] / alm_rho[1], 2));
        cost += cost_endp;
        tm[TOPAY_TERM_ENDP] += cost_endp;
        bool bad = false;
        for (int t = 0; t < TOPAY_NTERMS; t++)
            if (std::isinf(tm[t]) || std::isnan(tm[t])) bad = true;
        if (bad) {  // quirk 5
            std::fill(gdC.begin(), gdC.end(), 0.0);
            std::fill(gdT.begin(), gdT.end(), 0.0);
            cost = 1.0e+22;
            return;
        }
        const double ax = alm_rho[0] * (final_xy_error[0] + alm_lambda[0] / alm_rho[0]);
        const double ay = alm_rho[1] * (final_xy_error[1] + alm_lambda[1] / alm_rho[1]);
        for (size_t t = 0; t < S; t++) {
            ChainX[t] += ax;
            ChainY[t] += ay;
        }
        chain_contract(store, ChainX, ChainY, IntegralChainCoeff, gdC, gdT);
    }

    // moma_traj_opt.cpp:142-498. The 1.0 s wall-clock cap (:403) is replaced by
    // opt.alm_max_rounds; wall_cap_s > 0 re-enables the reference's cap as well.
    bool optimize_traj(const double* init_path, int len, const double* bvel, const double* bacc, Vec* x_out,
                       double wall_cap_s = 0.0) {
        stats = LbfgsStats();
        alm_rounds = 0;
        Vec x = prepare(init_path, len, bvel, bacc);
        topay_lbfgs_params p1 = opt.s1_lbfgs;
        p1.past = s1_past;
        double cost = 0.0;
        ProgressFn tracer = nullptr;
        if (trace)
            tracer = [this](const Vec&, const Vec&, double fx, double step, int k, int ls) {
                trace->push_back(fx);
                trace->push_back(step);
                trace->push_back((double)k);
                trace->push_back((double)ls);
                return 0;
            };
        int result = lbfgs_optimize(
            x, cost, [this](const Vec& xx, Vec& gg) { return cost_callback(1, xx, gg); }, tracer, p1, &stats);
        last_code = result;
        if (!(result == TOPAY_LBFGS_CONVERGENCE || result == TOPAY_LBFGS_CANCELED || result == TOPAY_LBFGS_STOP ||
              result == TOPAY_LBFGSERR_MAXIMUMITERATION)) {
            if (x_out) *x_out = x;
            return false;
        }
        for (int i = 0; i < 2; i++) {
            alm_lambda[i] = opt.alm_init_lambda[i];
            alm_rho[i] = opt.alm_init_rho[i];
        }
        int iter_num = 0;
        bool success = false;
        const auto t0 = std::chrono::steady_clock::now();
        const int max_it = opt.s2_lbfgs.max_iterations;
        ProgressFn early_exit = [this, max_it](const Vec&, const Vec&, double fx, double step, int k, int ls) {
            if (trace) {
                trace->push_back(fx);
                trace->push_back(step);
                trace->push_back((double)k);
                trace->push_back((double)ls);
            }
            return k > max_it ? 1 : 0;  // moma_traj_opt.cpp:1873
        };
        while (true) {
            if (wall_cap_s > 0.0 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > wall_cap_s)
                break;
            if (iter_num >= opt.alm_max_rounds) break;
            iter_num++;
            result = lbfgs_optimize(
                x, cost, [this](const Vec& xx, Vec& gg) { return cost_callback(2, xx, gg); }, early_exit,
                opt.s2_lbfgs, &stats);
            last_code = result;
            if (result == TOPAY_LBFGS_CONVERGENCE || result == TOPAY_LBFGS_CANCELED || result == TOPAY_LBFGS_STOP ||
                result == TOPAY_LBFGSERR_MAXIMUMITERATION) {
            } else if (result == TOPAY_LBFGSERR_MAXIMUMLINESEARCH) {
            } else {
                success = false;
                break;
            }
            const double en = std::sqrt(final_xy_error[0] * final_xy_error[0] + final_xy_error[1] * final_xy_error[1]);
            if (en < opt.alm_tolerance) {
                success = true;
                break;
            }
            alm_lambda[0] += alm_rho[0] * final_xy_error[0];
            alm_lambda[1] += alm_rho[1] * final_xy_error[1];
            alm_rho[0] = std::min((1 + opt.alm_gamma[0]) * alm_rho[0], opt.alm_rho_max[0]);
            alm_rho[1] = std::min((1 + opt.alm_gamma[1]) * alm_rho[1], opt.alm_rho_max[1]);
        }
        alm_rounds = iter_num;
        traj_cost = cost;
        if (x_out) *x_out = x;
        return success;
    }
};

// optimizer.yaml defaults
inline void opt_defaults(topay_opt_params* o) {
    o->int_K = 12;
    o->min_piece_num = 3;
    o->relu_mu = 1.0e-3;
    o->sample_interval = 1.5;
    o->energy_weights[0] = 0.33;
    for (int i = 1; i < 9; i++) o->energy_weights[i] = 1.0;
    o->s1_time_weight = 20.0;
    o->s1_moment_weight = 1000.0;
    o->s1_acc_weight = 1000.0;
    o->s1_domega_weight = 1000.0;
    o->s1_mean_time_weight = 1000.0;
    o->s1_path_pos_weight = 200000.0;
    o->s1_lbfgs_normal_past = 2;
    o->s1_lbfgs_shot_path_past = 8;
    o->s1_shot_path_horizon = 0.5;
    topay_lbfgs_params d;
    d.mem_size = 8;
    d.g_epsilon = 1.0e-5;
    d.past = 3;
    d.delta = 1.0e-6;
    d.max_iterations = 0;
    d.max_linesearch = 64;
    d.min_step = 1.0e-20;
    d.max_step = 1.0e+20;
    d.f_dec_coeff = 1.0e-4;
    d.s_curv_coeff = 0.9;
    d.cautious_factor = 1.0e-6;
    d.machine_prec = 1.0e-16;
    o->s1_lbfgs = d;
    o->s1_lbfgs.mem_size = 256;
    o->s1_lbfgs.g_epsilon = 0.0;
    o->s1_lbfgs.min_step = 0.0;
    o->s1_lbfgs.delta = 1.0e-2;
    o->s1_lbfgs.max_iterations = 8000;
    o->s1_lbfgs.past = 2;
    o->s2_time_weight = 50.0;
    o->s2_moment_weight = 300.0;
    o->s2_acc_weight = 3000.0;
    o->s2_domega_weight = 3000.0;
    o->s2_collision_weight = 500000.0;
    o->s2_mani_colli_weight = 500000.0;
    o->s2_self_colli_weight = 500000.0;
    o->s2_mani_pos_weight = 500.0;
    o->s2_mani_vel_weight = 500.0;
    o->s2_mani_acc_weight = 500.0;
    o->s2_mean_time_weight = 5000.0;
    o->s2_lbfgs = d;
    o->s2_lbfgs.mem_size = 256;
    o->s2_lbfgs.past = 3;
    o->s2_lbfgs.g_epsilon = 0.0;
    o->s2_lbfgs.min_step = 1.0e-32;
    o->s2_lbfgs.delta = 1.0e-4;
    o->s2_lbfgs.max_iterations = 8000;
    for (int i = 0; i < 2; i++) {
        o->alm_init_lambda[i] = 0.0;
        o->alm_init_rho[i] = 10000.0;
        o->alm_rho_max[i] = 1.0e+10;
        o->alm_gamma[i] = 9.0;
    }
    o->alm_tolerance = 0.01;
    o->alm_max_rounds = 20;
}

}  // namespace oracle
