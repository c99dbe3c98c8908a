// Host side of topay_solver: allocation, candidate upload, the tick loop and downloads.
// Replaces the per-thread MomaTrajOpt instances of the reference
// (src/planner/src/planner.cpp:59-66, 847-918) with one batched device solve.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <string>
#include <vector>

#include "common_host.h"
#include "solver_kernels.cuh"
#include "traj_host.h"

#include <atomic>

// solves in flight in this process (plans driven by concurrent host threads)
static std::atomic<int> g_solves_running{0};
struct TpRunningGuard {
    int concurrent;
    TpRunningGuard() : concurrent(++g_solves_running) {}
    ~TpRunningGuard() { --g_solves_running; }
};

struct topay_solver {
    TpSolverDev dev;
    TpParams params;
    topay_field* field;          // dense GridMap field, or
    topay_rogfield* rog;         // the ROG-Map ring (GridMap's use_rog branches); exactly one is set
    int device;
    cudaStream_t stream;
    int n_cand;          // candidates of the uploaded batch
    int max_N;           // largest piece count in the batch
    // pinned host staging (one cudaMallocHost block sized for max_cand): initial per-candidate state and x,
    // and the packed problem data on their way to the device
    char* h_pin;
    TpCandState* h_state;
    double *p_head, *p_tail, *p_sxy, *p_exy, *p_ixy;
    std::vector<void*> allocs;
    int32_t* h_active;   // pinned
    unsigned long long* h_nodes;  // pinned
    int slots;
    std::vector<cudaEvent_t> ev;  // event pool for the penalty kernel timing
    cudaEvent_t ev_begin, ev_end;
    bool spin_sync;         // TOPAY_SPIN_SYNC=1: always spin on the stream (lowest latency, one busy core per plan);
                            // default: spin while at most two solves run in this process, sleep beyond that
    cudaEvent_t ev_batch;   // blocking-sync event: the host thread sleeps while a batch of ticks runs
                            // (P plans in flight x N ranks would otherwise spin on as many cores)
    topay_solver_stats stats;
    size_t smem_cand;
    // initial state kept on the host for repeated runs
    double* h_x0;
    size_t h_x0_count;
    // one batch of `slots` ticks as a CUDA graph (launch-bound small plans); rebuilt when the
    // launch geometry or a captured pointer changes
    bool timed;                 // per-launch CUDA events around k_penalty (plain launches, no graph)
    cudaGraphExec_t graph_exec;
    int graph_n_cand, graph_max_N;
    // success gate (topay_solver_check_feasible): built on first use
    TpTrajChecker* checker;
    int32_t* d_pn;
    double* d_start;
};

namespace {

void solver_grid(const topay_solver* s, TpGrid* G) {
    if (s->rog) tp_rogfield_grid(s->rog, G);
    else tp_field_grid(s->field, G);
}
bool solver_field_ready(const topay_solver* s) { return s->rog ? tp_rogfield_ready(s->rog) : tp_field_ready(s->field); }

template <typename T>
int dev_alloc(topay_solver* s, T** p, size_t count) {
    void* q = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) {
        tp_set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        return TOPAY_ERR_ALLOC;
    }
    cudaMemsetAsync(q, 0, count * sizeof(T), s->stream);
    s->allocs.push_back(q);
    *p = (T*)q;
    return TOPAY_OK;
}

int kpad_of(int K) {
    int p = 4;
    while (p < K) p <<= 1;
    return p;
}

size_t penalty_smem() {
    return (size_t)(72 * TP_PEN_WARPS * 32 + TP_PEN_WARPS * 8 * 54) * sizeof(double);
}

void launch_eval(topay_solver* s, bool timed, int tick_in_batch) {
    const TpSolverDev& D = s->dev;
    const int groups = (s->max_N + D.ppw - 1) / D.ppw;
    const int end_warps = D.end_tasks ? (s->max_N + 31) / 32 : 0;
    dim3 blk(TP_WARPS_PER_BLOCK * 32);
    dim3 g1((groups + TP_WARPS_PER_BLOCK - 1) / TP_WARPS_PER_BLOCK, s->n_cand);
    dim3 blk2(TP_PEN_WARPS * 32);
    dim3 g2((groups + end_warps + TP_PEN_WARPS - 1) / TP_PEN_WARPS, s->n_cand);
    const size_t sm_int = (size_t)TP_WARPS_PER_BLOCK * D.ppw * 3 * (2 * D.K + 1) * sizeof(double);
    const size_t sm_pen = penalty_smem();
    TpGrid G;
    solver_grid(s, &G);
    if (timed) cudaEventRecord(s->ev[5 * tick_in_batch], s->stream);
    k_integrate<<<g1, blk, sm_int, s->stream>>>(D);
    if (timed) cudaEventRecord(s->ev[5 * tick_in_batch + 1], s->stream);
    switch (D.Kpad) {
        case 4: k_penalty<4><<<g2, blk2, sm_pen, s->stream>>>(D, s->params, G, groups); break;
        case 8: k_penalty<8><<<g2, blk2, sm_pen, s->stream>>>(D, s->params, G, groups); break;
        case 16: k_penalty<16><<<g2, blk2, sm_pen, s->stream>>>(D, s->params, G, groups); break;
        default: k_penalty<32><<<g2, blk2, sm_pen, s->stream>>>(D, s->params, G, groups); break;
    }
    if (timed) cudaEventRecord(s->ev[5 * tick_in_batch + 2], s->stream);
    k_chain<<<g1, blk, 0, s->stream>>>(D, s->params);
    if (timed) cudaEventRecord(s->ev[5 * tick_in_batch + 3], s->stream);
    s->stats.kernel_launches += 3;
    s->stats.eval_launches += 1;
}

void launch_cand(topay_solver* s, int mode, int slot) {
    k_cand<<<s->n_cand, TP_CAND_THREADS, s->smem_cand, s->stream>>>(s->dev, s->params, mode, slot);
    s->stats.kernel_launches += 1;
}

}  // namespace

static int solver_create(const topay_opt_params* opt, const topay_robot_params* robot, topay_field* field,
                         topay_rogfield* rog, int max_cand, int max_pieces, topay_solver** out);

extern "C" int topay_solver_create(const topay_opt_params* opt, const topay_robot_params* robot,
                                   topay_field* field, int max_cand, int max_pieces, topay_solver** out) {
    if (!field) return TOPAY_ERR_INVALID_ARG;
    return solver_create(opt, robot, field, nullptr, max_cand, max_pieces, out);
}

extern "C" int topay_solver_create_rog(const topay_opt_params* opt, const topay_robot_params* robot,
                                       topay_rogfield* rog, int max_cand, int max_pieces, topay_solver** out) {
    if (!rog) return TOPAY_ERR_INVALID_ARG;
    return solver_create(opt, robot, nullptr, rog, max_cand, max_pieces, out);
}

static int solver_create(const topay_opt_params* opt, const topay_robot_params* robot, topay_field* field,
                         topay_rogfield* rog, int max_cand, int max_pieces, topay_solver** out) {
    if (!opt || !robot || !out || max_cand < 1 || max_pieces < 1) return TOPAY_ERR_INVALID_ARG;
    if (opt->int_K < 1 || opt->int_K > TP_MAX_K) {
        tp_set_error("int_K must be in [1, 32]");
        return TOPAY_ERR_TOO_LARGE;
    }
    if (opt->s1_lbfgs_normal_past > TP_LBFGS_MAX_PAST || opt->s1_lbfgs_shot_path_past > TP_LBFGS_MAX_PAST ||
        opt->s2_lbfgs.past > TP_LBFGS_MAX_PAST) {
        tp_set_error("lbfgs past above the supported ring size");
        return TOPAY_ERR_TOO_LARGE;
    }
    if (topay_num_vars(max_pieces) > TP_EPT * TP_CAND_THREADS || max_pieces > 64 ||
        std::max(opt->s1_lbfgs.mem_size, opt->s2_lbfgs.mem_size) > 256) {
        tp_set_error("max_pieces above 64 or lbfgs mem_size above 256 is not supported");
        return TOPAY_ERR_TOO_LARGE;
    }
    const int dev = rog ? tp_rogfield_device(rog) : tp_field_device(field);
    int rc = tp_require_device(dev);
    if (rc != TOPAY_OK) return rc;
    topay_solver* s = new topay_solver();
    s->field = field;
    s->rog = rog;
    s->device = dev;
    s->n_cand = 0;
    s->max_N = 0;
    s->slots = 16;
    s->timed = false;
    s->graph_exec = nullptr;
    s->graph_n_cand = s->graph_max_N = -1;
    s->checker = nullptr;
    s->spin_sync = getenv("TOPAY_SPIN_SYNC") && atoi(getenv("TOPAY_SPIN_SYNC")) != 0;
    s->d_pn = nullptr;
    s->d_start = nullptr;
    memset(&s->stats, 0, sizeof(s->stats));
    cudaSetDevice(s->device);
    TP_CUDA_OK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), { delete s; });
    memset(&s->params, 0, sizeof(s->params));
    s->params.robot = *robot;
    s->params.opt = *opt;
    tp_derive_params(s->params);
    TpSolverDev& D = s->dev;
    memset(&D, 0, sizeof(D));
    D.max_cand = max_cand;
    D.max_pieces = max_pieces;
    D.K = opt->int_K;
    D.Kpad = kpad_of(D.K);
    D.ppw = 32 / D.Kpad;
    D.end_tasks = D.K == D.Kpad ? 1 : 0;
    D.xs = (topay_num_vars(max_pieces) + 3) & ~3;
    D.mem = std::max(1, std::max(opt->s1_lbfgs.mem_size, opt->s2_lbfgs.mem_size));
    const size_t C = max_cand, NP = max_pieces, K = D.K;
#define ALLOC(ptr, count)                                    \
    if ((rc = dev_alloc(s, &(ptr), (count))) != TOPAY_OK) {  \
        topay_solver_destroy(s);                             \
        return rc;                                           \
    }
    ALLOC(D.st, C);
    ALLOC(D.head_pva, C * 27);
    ALLOC(D.tail_pva, C * 27);
    ALLOC(D.start_xy, C * 2);
    ALLOC(D.end_xy, C * 2);
    ALLOC(D.init_inner_xy, C * NP * 2);
    ALLOC(D.x, C * D.xs);
    ALLOC(D.g, C * D.xs);
    ALLOC(D.xp, C * D.xs);
    ALLOC(D.gp, C * D.xs);
    ALLOC(D.d, C * D.xs);
    ALLOC(D.lm_s, C * D.mem * D.xs);
    ALLOC(D.lm_y, C * D.mem * D.xs);
    ALLOC(D.lm_ys, C * D.mem);
    ALLOC(D.lm_alpha, C * D.mem);
    ALLOC(D.T, C * NP);
    ALLOC(D.coeff, C * 6 * NP * 9);
    ALLOC(D.lu, C * 6 * NP * TP_BAND);
    ALLOC(D.Ixy, C * NP * K * 2);
    ALLOC(D.tot, C * NP * 2);
    ALLOC(D.gnode, C * NP * (K + 1) * 2);
    ALLOC(D.gsum, C * NP * 2);
    ALLOC(D.gdC, C * 6 * NP * 9);
    ALLOC(D.gdC_end, C * NP * 54);
    ALLOC(D.gdT, C * NP);
    ALLOC(D.gdT_end, C * NP);
    ALLOC(D.terms, C * NP * TOPAY_NTERMS);
    ALLOC(D.terms_end, C * NP * TOPAY_NTERMS);
    ALLOC(D.f, C);
    ALLOC(D.term_out, C * TOPAY_NTERMS);
    ALLOC(D.n_active, (size_t)s->slots);
    ALLOC(D.node_count, 2);
#undef ALLOC
    {
        const size_t C_ = (size_t)max_cand;
        const size_t b_state = ((C_ * sizeof(TpCandState) + 15) / 16) * 16;
        const size_t n_dbl = C_ * D.xs + C_ * (27 + 27 + 2 + 2) + C_ * NP * 2;
        TP_CUDA_OK(cudaMallocHost(&s->h_pin, b_state + n_dbl * sizeof(double)), { topay_solver_destroy(s); });
        s->h_state = reinterpret_cast<TpCandState*>(s->h_pin);
        double* dp = reinterpret_cast<double*>(s->h_pin + b_state);
        s->h_x0 = dp;            dp += C_ * D.xs;
        s->p_head = dp;          dp += C_ * 27;
        s->p_tail = dp;          dp += C_ * 27;
        s->p_sxy = dp;           dp += C_ * 2;
        s->p_exy = dp;           dp += C_ * 2;
        s->p_ixy = dp;
        s->h_x0_count = 0;
    }
    TP_CUDA_OK(cudaMallocHost(&s->h_active, s->slots * sizeof(int32_t)), { topay_solver_destroy(s); });
    TP_CUDA_OK(cudaMallocHost(&s->h_nodes, 2 * sizeof(unsigned long long)), { topay_solver_destroy(s); });
    s->ev.resize(5 * s->slots);
    for (auto& e : s->ev) cudaEventCreate(&e);
    cudaEventCreate(&s->ev_begin);
    cudaEventCreate(&s->ev_end);
    cudaEventCreateWithFlags(&s->ev_batch, cudaEventBlockingSync | cudaEventDisableTiming);
    // LU band + one right-hand-side matrix, or the two-loop's TMA ring + its two 256-entry tables + one
    // vector of scratch (single-warp variant)
    // at least six full-width stages; small solvers still get 64 KB so that short rows ride a deep ring
    s->smem_cand = std::max(std::max((size_t)6 * NP * (TP_BAND + 9), (size_t)6 * 2 * D.xs + 512 + D.xs) * sizeof(double),
                            (size_t)64 * 1024);
    D.smem_doubles = (int32_t)(s->smem_cand / sizeof(double));
    if (s->smem_cand > 227 * 1024) {
        tp_set_error("max_pieces too large for the per-candidate shared-memory working set");
        topay_solver_destroy(s);
        return TOPAY_ERR_TOO_LARGE;
    }
    // the attribute belongs to the kernel, not to this solver: solvers of different capacities coexist
    // (bench: 64-piece plans and 16-piece latency plans), so it only ever grows
    {
        static std::atomic<size_t> attr_set{0};
        size_t cur = attr_set.load();
        while (cur < s->smem_cand && !attr_set.compare_exchange_weak(cur, s->smem_cand)) {}
        TP_CUDA_OK(cudaFuncSetAttribute(k_cand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attr_set.load()),
                   { topay_solver_destroy(s); });
    }
    const size_t sm_int = (size_t)TP_WARPS_PER_BLOCK * D.ppw * 3 * (2 * D.K + 1) * sizeof(double);
    {
        static std::atomic<size_t> attr_int{0};
        size_t cur = attr_int.load();
        while (cur < sm_int && !attr_int.compare_exchange_weak(cur, sm_int)) {}
        TP_CUDA_OK(cudaFuncSetAttribute(k_integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attr_int.load()),
                   { topay_solver_destroy(s); });
    }
    {
        const int smp = (int)penalty_smem();
        TP_CUDA_OK(cudaFuncSetAttribute(k_penalty<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smp), { topay_solver_destroy(s); });
        TP_CUDA_OK(cudaFuncSetAttribute(k_penalty<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smp), { topay_solver_destroy(s); });
        TP_CUDA_OK(cudaFuncSetAttribute(k_penalty<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smp), { topay_solver_destroy(s); });
        TP_CUDA_OK(cudaFuncSetAttribute(k_penalty<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smp), { topay_solver_destroy(s); });
    }
    TP_CUDA_OK(cudaStreamSynchronize(s->stream), { topay_solver_destroy(s); });
    *out = s;
    return TOPAY_OK;
}

extern "C" void topay_solver_destroy(topay_solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    delete s->checker;
    for (void* p : s->allocs) cudaFree(p);
    if (s->dev.trace) cudaFree(s->dev.trace);
    if (s->dev.trace_len) cudaFree(s->dev.trace_len);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    if (s->h_active) cudaFreeHost(s->h_active);
    if (s->h_nodes) cudaFreeHost(s->h_nodes);
    for (auto& e : s->ev) cudaEventDestroy(e);
    if (s->ev_begin) cudaEventDestroy(s->ev_begin);
    if (s->ev_end) cudaEventDestroy(s->ev_end);
    if (s->ev_batch) cudaEventDestroy(s->ev_batch);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// Uploads problem data + x and (re)initialises the per-candidate state.
static int upload_problem(topay_solver* s, int n_cand, const int32_t* piece_num, const double* head,
                          const double* tail, const double* sxy, const double* exy, const double* inner_xy,
                          const double* x, int x_stride, const int32_t* s1_past, int phase, const double* lambda,
                          const double* rho) {
    TpSolverDev& D = s->dev;
    if (n_cand < 1 || n_cand > D.max_cand) {
        tp_set_error("n_cand above the solver capacity");
        return TOPAY_ERR_TOO_LARGE;
    }
    int maxN = 0;
    for (int c = 0; c < n_cand; c++) {
        if (piece_num[c] < 1 || piece_num[c] > D.max_pieces) {
            tp_set_error("piece_num above the solver capacity");
            return TOPAY_ERR_TOO_LARGE;
        }
        maxN = std::max(maxN, piece_num[c]);
    }
    cudaSetDevice(s->device);
    s->n_cand = n_cand;
    s->max_N = maxN;
    double* xs = s->h_x0;
    memset(xs, 0, (size_t)n_cand * D.xs * sizeof(double));
    s->h_x0_count = (size_t)n_cand * D.xs;
    for (int c = 0; c < n_cand; c++) {
        TpCandState& st = s->h_state[c];
        memset(&st, 0, sizeof(st));
        st.phase = phase;
        st.N = piece_num[c];
        st.n = topay_num_vars(piece_num[c]);
        st.ls_init = 1;
        st.s1_past = s1_past ? s1_past[c] : s->params.opt.s1_lbfgs.past;
        st.lambda[0] = lambda ? lambda[2 * c] : s->params.opt.alm_init_lambda[0];
        st.lambda[1] = lambda ? lambda[2 * c + 1] : s->params.opt.alm_init_lambda[1];
        st.rho[0] = rho ? rho[2 * c] : s->params.opt.alm_init_rho[0];
        st.rho[1] = rho ? rho[2 * c + 1] : s->params.opt.alm_init_rho[1];
        memcpy(&xs[(size_t)c * D.xs], x + (size_t)c * x_stride, st.n * sizeof(double));
    }
    // the packed problem goes through the pinned staging block
    memcpy(s->p_head, head, (size_t)n_cand * 27 * 8);
    memcpy(s->p_tail, tail, (size_t)n_cand * 27 * 8);
    memcpy(s->p_sxy, sxy, (size_t)n_cand * 2 * 8);
    memcpy(s->p_exy, exy, (size_t)n_cand * 2 * 8);
    memcpy(s->p_ixy, inner_xy, (size_t)n_cand * D.max_pieces * 2 * 8);
    cudaStream_t q = s->stream;
    TP_CUDA_OK(cudaMemcpyAsync(D.st, s->h_state, n_cand * sizeof(TpCandState), cudaMemcpyHostToDevice, q), {});
    TP_CUDA_OK(cudaMemcpyAsync(D.head_pva, s->p_head, (size_t)n_cand * 27 * 8, cudaMemcpyHostToDevice, q), {});
    TP_CUDA_OK(cudaMemcpyAsync(D.tail_pva, s->p_tail, (size_t)n_cand * 27 * 8, cudaMemcpyHostToDevice, q), {});
    TP_CUDA_OK(cudaMemcpyAsync(D.start_xy, s->p_sxy, (size_t)n_cand * 2 * 8, cudaMemcpyHostToDevice, q), {});
    TP_CUDA_OK(cudaMemcpyAsync(D.end_xy, s->p_exy, (size_t)n_cand * 2 * 8, cudaMemcpyHostToDevice, q), {});
    TP_CUDA_OK(cudaMemcpyAsync(D.init_inner_xy, s->p_ixy, (size_t)n_cand * D.max_pieces * 2 * 8,
                               cudaMemcpyHostToDevice, q), {});
    TP_CUDA_OK(cudaMemcpyAsync(D.x, xs, s->h_x0_count * 8, cudaMemcpyHostToDevice, q), {});
    TP_CUDA_OK(cudaStreamSynchronize(q), {});
    return TOPAY_OK;
}

extern "C" int topay_solver_eval(topay_solver* s, int stage, const topay_problem_batch* prob, const double* x,
                                 int x_stride, double* cost, double* grad, double* term_costs, double* coeff_out,
                                 double* final_xy_out) {
    if (!s || !prob || !x || !cost || !grad || (stage != 1 && stage != 2)) return TOPAY_ERR_INVALID_ARG;
    if (!solver_field_ready(s)) {
        tp_set_error("field not built: call topay_field_rebuild first");
        return TOPAY_ERR_NOT_READY;
    }
    int rc = upload_problem(s, prob->n_cand, prob->piece_num, prob->head_pva, prob->tail_pva, prob->start_xy,
                            prob->end_xy, prob->init_inner_xy, x, x_stride, nullptr, stage, prob->alm_lambda,
                            prob->alm_rho);
    if (rc != TOPAY_OK) return rc;
    const TpSolverDev& D = s->dev;
    const int n = prob->n_cand;
    cudaMemsetAsync(D.n_active, 0, s->slots * sizeof(int32_t), s->stream);
    launch_cand(s, TP_MODE_GEN, 0);
    launch_eval(s, false, 0);
    launch_cand(s, TP_MODE_ADJ, 1);
    TP_CUDA_OK(cudaStreamSynchronize(s->stream), {});
    TP_CUDA_OK(cudaGetLastError(), {});
    std::vector<double> g((size_t)n * D.xs);
    TP_CUDA_OK(cudaMemcpy(cost, D.f, n * 8, cudaMemcpyDeviceToHost), {});
    TP_CUDA_OK(cudaMemcpy(g.data(), D.g, g.size() * 8, cudaMemcpyDeviceToHost), {});
    for (int c = 0; c < n; c++)
        memcpy(grad + (size_t)c * x_stride, &g[(size_t)c * D.xs], topay_num_vars(prob->piece_num[c]) * 8);
    if (term_costs) TP_CUDA_OK(cudaMemcpy(term_costs, D.term_out, (size_t)n * TOPAY_NTERMS * 8, cudaMemcpyDeviceToHost), {});
    if (coeff_out)
        TP_CUDA_OK(cudaMemcpy(coeff_out, D.coeff, (size_t)n * 6 * D.max_pieces * 9 * 8, cudaMemcpyDeviceToHost), {});
    if (final_xy_out) {
        TP_CUDA_OK(cudaMemcpy(s->h_state, D.st, n * sizeof(TpCandState), cudaMemcpyDeviceToHost), {});
        for (int c = 0; c < n; c++) {
            final_xy_out[2 * c] = s->h_state[c].final_xy[0];
            final_xy_out[2 * c + 1] = s->h_state[c].final_xy[1];
        }
    }
    return TOPAY_OK;
}

extern "C" int topay_solver_upload(topay_solver* s, int n_cand, const int32_t* path_len, const double* init_paths,
                                   const double* bvel, const double* bacc) {
    if (!s || !path_len || !init_paths || !bvel || !bacc) return TOPAY_ERR_INVALID_ARG;
    const TpSolverDev& D = s->dev;
    if (n_cand < 1 || n_cand > D.max_cand) {
        tp_set_error("n_cand above the solver capacity");
        return TOPAY_ERR_TOO_LARGE;
    }
    const int NP = D.max_pieces, xstride = topay_num_vars(NP);
    std::vector<int32_t> pn(n_cand), past(n_cand);
    std::vector<double> head((size_t)n_cand * 27), tail((size_t)n_cand * 27), sxy((size_t)n_cand * 2),
        exy((size_t)n_cand * 2), ixy((size_t)n_cand * NP * 2), x((size_t)n_cand * xstride, 0.0);
    size_t off = 0;
    for (int c = 0; c < n_cand; c++) {
        int rc = topay_prepare_candidate(&s->params.opt, &s->params.robot, init_paths + off * 10, path_len[c],
                                         bvel + (size_t)c * 20, bacc + (size_t)c * 20, NP, &pn[c], &head[c * 27],
                                         &tail[c * 27], &sxy[c * 2], &exy[c * 2], &ixy[(size_t)c * NP * 2],
                                         &x[(size_t)c * xstride], &past[c]);
        if (rc != TOPAY_OK) {
            tp_set_error("candidate needs more pieces than max_pieces");
            return rc;
        }
        off += path_len[c];
    }
    return upload_problem(s, n_cand, pn.data(), head.data(), tail.data(), sxy.data(), exy.data(), ixy.data(),
                          x.data(), xstride, past.data(), 1, nullptr, nullptr);
}

extern "C" int topay_solver_run(topay_solver* s) {
    if (!s || s->n_cand < 1) return TOPAY_ERR_INVALID_ARG;
    if (!solver_field_ready(s)) {
        tp_set_error("field not built: call topay_field_rebuild first");
        return TOPAY_ERR_NOT_READY;
    }
    cudaSetDevice(s->device);
    TpSolverDev& D = s->dev;
    cudaStream_t q = s->stream;
    TpRunningGuard running;
    memset(&s->stats, 0, sizeof(s->stats));
    // reset to the uploaded initial state so that repeated runs do identical work
    TP_CUDA_OK(cudaMemcpyAsync(D.st, s->h_state, s->n_cand * sizeof(TpCandState), cudaMemcpyHostToDevice, q), {});
    TP_CUDA_OK(cudaMemcpyAsync(D.x, s->h_x0, s->h_x0_count * 8, cudaMemcpyHostToDevice, q), {});
    cudaMemsetAsync(D.node_count, 0, 2 * sizeof(unsigned long long), q);
    if (D.trace) cudaMemsetAsync(D.trace_len, 0, (size_t)D.max_cand * sizeof(int32_t), q);
    cudaEventRecord(s->ev_begin, q);
    cudaMemsetAsync(D.n_active, 0, s->slots * sizeof(int32_t), q);
    launch_cand(s, TP_MODE_GEN, 0);
    double ms_eval = 0.0, ms_k[3] = {0.0, 0.0, 0.0}, full_ms_cand = 0.0;
    unsigned long long hist_prev = 0, full_hist = 0;
    long long full_ticks = 0;
    // hard cap on ticks: every candidate does at most this many evaluations
    const long long max_ticks =
        (long long)(s->params.opt.alm_max_rounds + 1) * ((long long)s->params.opt.s2_lbfgs.max_iterations + 2) *
        (s->params.opt.s2_lbfgs.max_linesearch + 1);
    long long ticks = 0;
    bool done = false;
    const bool use_graph = !s->timed;
    if (use_graph && (!s->graph_exec || s->graph_n_cand != s->n_cand || s->graph_max_N != s->max_N)) {
        if (s->graph_exec) {
            cudaGraphExecDestroy(s->graph_exec);
            s->graph_exec = nullptr;
        }
        cudaGraph_t g = nullptr;
        const topay_solver_stats keep = s->stats;
        TP_CUDA_OK(cudaStreamBeginCapture(q, cudaStreamCaptureModeThreadLocal), {});
        cudaMemsetAsync(D.n_active, 0, s->slots * sizeof(int32_t), q);
        for (int t = 0; t < s->slots; t++) {
            launch_eval(s, false, t);
            launch_cand(s, TP_MODE_ADJ | TP_MODE_ADVANCE | TP_MODE_GEN, t);
        }
        cudaMemcpyAsync(s->h_active, D.n_active, s->slots * sizeof(int32_t), cudaMemcpyDeviceToHost, q);
        TP_CUDA_OK(cudaStreamEndCapture(q, &g), {});
        TP_CUDA_OK(cudaGraphInstantiate(&s->graph_exec, g, 0), { cudaGraphDestroy(g); });
        cudaGraphDestroy(g);
        s->graph_n_cand = s->n_cand;
        s->graph_max_N = s->max_N;
        s->stats = keep;   // the capture pass launched nothing
    }
    while (!done && ticks < max_ticks) {
        if (use_graph) {
            TP_CUDA_OK(cudaGraphLaunch(s->graph_exec, q), {});
            s->stats.kernel_launches += 4 * s->slots;
            s->stats.eval_launches += s->slots;
        } else {
            cudaMemsetAsync(D.n_active, 0, s->slots * sizeof(int32_t), q);
            for (int t = 0; t < s->slots; t++) {
                launch_eval(s, true, t);
                launch_cand(s, TP_MODE_ADJ | TP_MODE_ADVANCE | TP_MODE_GEN, t);
                cudaEventRecord(s->ev[5 * t + 4], q);
            }
            cudaMemcpyAsync(s->h_active, D.n_active, s->slots * sizeof(int32_t), cudaMemcpyDeviceToHost, q);
            cudaMemcpyAsync(s->h_nodes, D.node_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, q);
        }
        ticks += s->slots;
        if (s->spin_sync || g_solves_running.load(std::memory_order_relaxed) <= 2) {
            TP_CUDA_OK(cudaStreamSynchronize(q), {});
        } else {
            cudaEventRecord(s->ev_batch, q);
            TP_CUDA_OK(cudaEventSynchronize(s->ev_batch), {});
        }
        if (getenv("TOPAY_TICK_LOG")) {   // dev: wall time of each 16-tick batch vs candidates still active
            static thread_local double t_prev = 0.0;
            timespec ts;
            clock_gettime(CLOCK_MONOTONIC, &ts);
            const double now = ts.tv_sec + 1e-9 * ts.tv_nsec;
            if (ticks > s->slots) fprintf(stderr, "TICKLOG %lld %d %.1f\n", ticks, s->h_active[s->slots - 1], (now - t_prev) * 1e6 / s->slots);
            t_prev = now;
        }
        if (!use_graph) {
            // batches in which every candidate was still solving: the full-activity figures
            const bool full = s->h_active[s->slots - 1] == s->n_cand;
            const unsigned long long hist_now = s->h_nodes[1];
            double cand_batch = 0.0;
            for (int t = 0; t < s->slots; t++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, s->ev[5 * t + 1], s->ev[5 * t + 2]);
                ms_eval += ms;
                cudaEventElapsedTime(&ms, s->ev[5 * t], s->ev[5 * t + 1]);
                ms_k[0] += ms;
                cudaEventElapsedTime(&ms, s->ev[5 * t + 2], s->ev[5 * t + 3]);
                ms_k[1] += ms;
                cudaEventElapsedTime(&ms, s->ev[5 * t + 3], s->ev[5 * t + 4]);
                ms_k[2] += ms;
                cand_batch += ms;
            }
            if (full) {
                full_ms_cand += cand_batch;
                full_hist += hist_now - hist_prev;
                full_ticks += s->slots;
            }
            hist_prev = hist_now;
        }
        done = s->h_active[s->slots - 1] == 0;
    }
    cudaEventRecord(s->ev_end, q);
    cudaMemcpyAsync(s->h_nodes, D.node_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, q);
    TP_CUDA_OK(cudaStreamSynchronize(q), {});
    TP_CUDA_OK(cudaGetLastError(), {});
    cudaEventElapsedTime(&s->stats.ms_total, s->ev_begin, s->ev_end);
    s->stats.ms_eval = (float)ms_eval;
    s->stats.ms_integrate = (float)ms_k[0];
    s->stats.ms_chain = (float)ms_k[1];
    s->stats.ms_cand = (float)ms_k[2];
    s->stats.ticks = ticks;
    s->stats.eval_nodes = (int64_t)s->h_nodes[0];
    s->stats.hist_bytes = (int64_t)s->h_nodes[1] * 16;
    s->stats.full_ticks = full_ticks;
    s->stats.full_hist_bytes = (int64_t)full_hist * 16;
    s->stats.full_ms_cand = (float)full_ms_cand;   // one s_j and one y_j element per row element
    return TOPAY_OK;
}

extern "C" int topay_solver_download(topay_solver* s, topay_result_batch* out, int32_t* best_by_duration,
                                     int32_t* best_by_cost) {
    if (!s || !out || !out->status || s->n_cand < 1) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(s->device);
    const TpSolverDev& D = s->dev;
    const int n = s->n_cand, NP = D.max_pieces;
    std::vector<TpCandState> st(n);
    TP_CUDA_OK(cudaMemcpy(st.data(), D.st, n * sizeof(TpCandState), cudaMemcpyDeviceToHost), {});
    std::vector<double> T((size_t)n * NP);
    TP_CUDA_OK(cudaMemcpy(T.data(), D.T, T.size() * 8, cudaMemcpyDeviceToHost), {});
    int bd = -1, bc = -1;
    double bdv = 0, bcv = 0;
    int64_t evals = 0;
    for (int c = 0; c < n; c++) {
        double dur = 0.0;
        for (int i = 0; i < st[c].N; i++) dur += T[(size_t)c * NP + i];
        out->status[c] = st[c].status;
        if (out->lbfgs_code) out->lbfgs_code[c] = st[c].last_code;
        if (out->piece_num) out->piece_num[c] = st[c].N;
        if (out->iters) out->iters[c] = st[c].iters_total;
        if (out->evals) out->evals[c] = st[c].evals_total;
        if (out->alm_rounds) out->alm_rounds[c] = st[c].alm_round;
        if (out->cost) out->cost[c] = st[c].cost;
        if (out->duration) out->duration[c] = dur;
        if (out->final_xy_err) {
            out->final_xy_err[2 * c] = st[c].final_xy[0];
            out->final_xy_err[2 * c + 1] = st[c].final_xy[1];
        }
        evals += st[c].evals_total;
        if (st[c].status == 1) {
            // planner.cpp:999-1010: first success, replaced only by a strictly shorter duration
            if (bd < 0 || dur < bdv) {
                bd = c;
                bdv = dur;
            }
            if (bc < 0 || st[c].cost < bcv) {
                bc = c;
                bcv = st[c].cost;
            }
        }
    }
    s->stats.evals_total = evals;
    if (out->T) memcpy(out->T, T.data(), T.size() * 8);
    if (out->coeff) TP_CUDA_OK(cudaMemcpy(out->coeff, D.coeff, (size_t)n * 6 * NP * 9 * 8, cudaMemcpyDeviceToHost), {});
    if (out->x) {
        std::vector<double> xs((size_t)n * D.xs);
        TP_CUDA_OK(cudaMemcpy(xs.data(), D.x, xs.size() * 8, cudaMemcpyDeviceToHost), {});
        const int xstride = topay_num_vars(NP);
        for (int c = 0; c < n; c++) memcpy(out->x + (size_t)c * xstride, &xs[(size_t)c * D.xs], st[c].n * 8);
    }
    if (best_by_duration) *best_by_duration = bd;
    if (best_by_cost) *best_by_cost = bc;
    return TOPAY_OK;
}

extern "C" int topay_solver_solve_batch(topay_solver* s, int n_cand, const int32_t* path_len,
                                        const double* init_paths, const double* bvel, const double* bacc,
                                        topay_result_batch* out, int32_t* best_by_duration, int32_t* best_by_cost) {
    int rc = topay_solver_upload(s, n_cand, path_len, init_paths, bvel, bacc);
    if (rc != TOPAY_OK) return rc;
    rc = topay_solver_run(s);
    if (rc != TOPAY_OK) return rc;
    return topay_solver_download(s, out, best_by_duration, best_by_cost);
}

// Piece counts live in the per-candidate state and the start pose is split over start_xy and the
// yaw slot of head_pva; gather both into the layout the trajectory kernels read.
__global__ void k_solver_traj_view(const TpCandState* __restrict__ st, const double* __restrict__ start_xy,
                                   const double* __restrict__ head_pva, int n, int32_t* pn, double* start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    pn[c] = st[c].N;
    start[3 * c] = start_xy[2 * c];
    start[3 * c + 1] = start_xy[2 * c + 1];
    start[3 * c + 2] = head_pva[(size_t)c * 27];
}

extern "C" int topay_solver_check_feasible(topay_solver* s, topay_feasibility* out, int32_t* best_success) {
    if (!s || !out || !out->feasible || s->n_cand < 1) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(s->device);
    TpSolverDev& D = s->dev;
    const int n = s->n_cand, NP = D.max_pieces;
    int rc;
    if (!s->checker) {
        if ((rc = dev_alloc(s, &s->d_pn, (size_t)D.max_cand)) != TOPAY_OK) return rc;
        if ((rc = dev_alloc(s, &s->d_start, (size_t)D.max_cand * 3)) != TOPAY_OK) return rc;
        s->checker = new TpTrajChecker();
        if ((rc = s->checker->init(s->device, s->stream)) != TOPAY_OK) return rc;
    }
    k_solver_traj_view<<<(n + 127) / 128, 128, 0, s->stream>>>(D.st, D.start_xy, D.head_pva, n, s->d_pn, s->d_start);
    TpTrajView V{n, NP, s->d_pn, D.T, D.coeff, s->d_start};
    TpGrid G;
    solver_grid(s, &G);
    std::vector<int32_t> fp;
    int32_t* keep_fp = out->feasible_print;
    if (!keep_fp) {   // the gate below needs it
        fp.resize(n);
        out->feasible_print = fp.data();
    }
    rc = s->checker->check(V, s->params, G, out);
    if (rc == TOPAY_OK && best_success) {
        // planner.cpp:877-880 + :999-1010: optimizeTraj && printConstraintsSituations, shortest duration
        std::vector<TpCandState> st(n);
        TP_CUDA_OK(cudaMemcpy(st.data(), D.st, n * sizeof(TpCandState), cudaMemcpyDeviceToHost), { out->feasible_print = keep_fp; });
        std::vector<int32_t> ok(n);
        std::vector<double> dur(n);
        for (int c = 0; c < n; c++) {
            ok[c] = st[c].status == 1 && out->feasible_print[c];
            dur[c] = s->checker->h_meta[c].total;
        }
        *best_success = topay_select_shortest(ok.data(), dur.data(), n);
    }
    out->feasible_print = keep_fp;
    return rc;
}

// Developer aid: raw copy of one of the evaluation's intermediate device arrays after topay_solver_eval
// (bit-level A/B runs). which: 0 gnode, 1 gsum, 2 gdC, 3 gdT, 4 tot, 5 Ixy, 6 g. Returns the number of
// doubles written (<= cap) or a negative status.
extern "C" int64_t topay_solver_debug_download(topay_solver* s, int which, double* out, int64_t cap) {
    if (!s || !out) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(s->device);
    const TpSolverDev& D = s->dev;
    const size_t C_ = (size_t)D.max_cand, NP = (size_t)D.max_pieces, K = (size_t)D.K;
    const double* src = nullptr;
    size_t n = 0;
    switch (which) {
        case 0: src = D.gnode; n = C_ * NP * (K + 1) * 2; break;
        case 1: src = D.gsum; n = C_ * NP * 2; break;
        case 2: src = D.gdC; n = C_ * 6 * NP * 9; break;
        case 3: src = D.gdT; n = C_ * NP; break;
        case 4: src = D.tot; n = C_ * NP * 2; break;
        case 5: src = D.Ixy; n = C_ * NP * K * 2; break;
        case 6: src = D.g; n = C_ * (size_t)D.xs; break;
        case 7: src = D.gdC_end; n = C_ * NP * 54; break;
        default: return TOPAY_ERR_INVALID_ARG;
    }
    if ((int64_t)n > cap) n = (size_t)cap;
    TP_CUDA_OK(cudaMemcpy(out, src, n * sizeof(double), cudaMemcpyDeviceToHost), {});
    return (int64_t)n;
}

// Dev profiling: accumulated clock64() deltas of k_cand's phases for candidate 0 (16 slots).
extern "C" int topay_solver_phase_clocks(topay_solver* s, int enable, long long* out16) {
    if (!s) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(s->device);
    TpSolverDev& D = s->dev;
    s->graph_n_cand = -1;   // captured kernel arguments change
    if (enable && !D.prof) {
        TP_CUDA_OK(cudaMalloc(&D.prof, 16 * sizeof(long long)), {});
        cudaMemset(D.prof, 0, 16 * sizeof(long long));
    }
    if (out16 && D.prof) {
        TP_CUDA_OK(cudaMemcpy(out16, D.prof, 16 * sizeof(long long), cudaMemcpyDeviceToHost), {});
        cudaMemset(D.prof, 0, 16 * sizeof(long long));
    }
    if (!enable && D.prof) {
        cudaFree(D.prof);
        D.prof = nullptr;
    }
    return TOPAY_OK;
}

extern "C" int topay_solver_set_timed(topay_solver* s, int timed) {
    if (!s) return TOPAY_ERR_INVALID_ARG;
    s->timed = timed != 0;
    return TOPAY_OK;
}

extern "C" int topay_solver_set_trace(topay_solver* s, int cap) {
    if (!s || cap < 0) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(s->device);
    TpSolverDev& D = s->dev;
    s->graph_n_cand = -1;   // captured kernel arguments change
    if (D.trace) {
        cudaFree(D.trace);
        cudaFree(D.trace_len);
        D.trace = nullptr;
        D.trace_len = nullptr;
    }
    D.trace_cap = cap;
    if (cap > 0) {
        TP_CUDA_OK(cudaMalloc(&D.trace, (size_t)D.max_cand * cap * 4 * sizeof(double)), {});
        TP_CUDA_OK(cudaMalloc(&D.trace_len, (size_t)D.max_cand * sizeof(int32_t)), {});
        cudaMemset(D.trace_len, 0, (size_t)D.max_cand * sizeof(int32_t));
    }
    return TOPAY_OK;
}

extern "C" int topay_solver_download_trace(topay_solver* s, int cand, double* out, int cap, int32_t* len) {
    if (!s || !out || !len || cand < 0 || cand >= s->dev.max_cand) return TOPAY_ERR_INVALID_ARG;
    const TpSolverDev& D = s->dev;
    if (!D.trace) return TOPAY_ERR_NOT_READY;
    cudaSetDevice(s->device);
    int32_t n = 0;
    TP_CUDA_OK(cudaMemcpy(&n, D.trace_len + cand, sizeof(int32_t), cudaMemcpyDeviceToHost), {});
    n = std::min(n, cap);
    TP_CUDA_OK(cudaMemcpy(out, D.trace + (size_t)cand * D.trace_cap * 4, (size_t)n * 4 * sizeof(double),
                          cudaMemcpyDeviceToHost), {});
    *len = n;
    return TOPAY_OK;
}

extern "C" int topay_solver_last_stats(topay_solver* s, topay_solver_stats* out) {
    if (!s || !out) return TOPAY_ERR_INVALID_ARG;
    *out = s->stats;
    return TOPAY_OK;
}
