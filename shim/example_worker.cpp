// The reference's planner worker (src/planner/src/planner.cpp:847-918) against the shims: what a
// maintainer's call site looks like after switching the include. Compile check:
//   g++ -std=c++17 -I include -c shim/example_worker.cpp
#include <array>
#include <cstdio>

#include "topay_shim.hpp"

using namespace nmoma_planner;

struct Mat10x2 {   // stands in for Eigen::MatrixXd::Zero(10, 2)
    double a[20] = {0};
    double operator()(int r, int c) const { return a[r * 2 + c]; }
};

int plan_all_candidates(GridMap::Ptr grid_map, const std::vector<std::vector<std::array<double, 10>>>& front_paths) {
    // planner.cpp:59-66 creates 8 MomaTrajOpt instances and runs them on threads; one batched instance
    // replaces them.
    MomaTrajOpt opt(grid_map, /*max_cand=*/8, /*max_pieces=*/32);
    opt.init();
    std::vector<std::vector<double>> flat(front_paths.size());
    for (size_t c = 0; c < front_paths.size(); c++)
        for (const auto& wp : front_paths[c]) flat[c].insert(flat[c].end(), wp.begin(), wp.end());
    std::vector<double> bvel(front_paths.size() * 20, 0.0), bacc(front_paths.size() * 20, 0.0);
    std::vector<int> ok = opt.optimizeTrajBatch(flat, bvel.data(), bacc.data());
    // planner.cpp:877-880 + 999-1010: optimizeTraj && printConstraintsSituations per candidate, then the
    // shortest successful trajectory wins — one device pass over all candidates
    std::vector<int> success;
    const int winner = opt.selectFeasible(&success);
    if (winner >= 0) {
        MomaTraj t = opt.getTraj(winner);
        std::vector<double> state = t.getState(0.5 * t.getTotalDuration());     // moma_traj_opt.h:121
        std::printf("winner %d, duration %.3f s, mid-state x %.3f\n", winner, t.getTotalDuration(), state[0]);
    }
    // single-candidate call, the reference's own signature (planner.cpp:878)
    Mat10x2 boundary_vel, boundary_acc;
    MomaTrajOpt one(grid_map);
    one.init();
    bool succ = one.optimizeTraj(front_paths[0], boundary_vel, boundary_acc) &&
                one.printConstraintsSituations(one.getTraj()) && one.getTraj().is_init;
    return succ ? winner : -1;
}

// use_rog: true — the ROG-Map side (rog_map.cpp:118, prob_map.cpp:292-298, 352-354, 518-528)
void rog_update(rog_map::ESDFMap& esdf, const std::array<double, 3>& odom, const double* hit_points, int64_t n_hits) {
    esdf.mapSliding(odom);
    std::vector<uint8_t> from(n_hits, TOPAY_ROG_UNKNOWN), to(n_hits, TOPAY_ROG_OCCUPIED);
    esdf.updateGridCounterBatch(hit_points, from.data(), to.data(), n_hits);
    esdf.updateESDF3D(odom);
    double d;
    std::array<double, 3> g;
    esdf.getValueGrad(odom, d, g);
}
