// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C entry points over the REFERENCE'S OWN ROG-Map ESDF layer, compiled unmodified from where it lies under
// /root/reference (oracle/Makefile target `_ref`):
//   src/rog_map/src/rog_map/esdf_map.cpp     ESDFMap::updateESDF3D + ring fillESDF, the six lookups, isLineFree2d
//   src/rog_map/src/rog_map/counter_map.cpp  CounterMap::updateGridCounter, initCounterMap
//   src/rog_map/src/rog_map/sliding_map.cpp  SlidingMap index maths, mapSliding
//   src/rog_map/src/utils/raycaster.cpp      the DDA isLineFree2d walks
//   src/rog_map/src/rog_map/prob_map.cpp     ProbMap::updateProbMap, raycastProcess, hit / miss updates (+ inf_map.cpp,
//                                            which ProbMap notifies unconditionally)
// (-DORIGIN_AT_CORNER as src/rog_map/CMakeLists.txt:14) against the stand-ins of oracle/ref_stubs. The entry
// points mirror oracle_capi.cpp's oracle_rog_* so tests compare the two bit for bit. `#define private public`
// only opens the classes to this driver (buffer downloads); the reference's sources are compiled as they are.
#include <array>
#include <cstring>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <Eigen/Eigen>

#define private public
#define protected public
#include <rog_map/esdf_map.h>
#include <rog_map/prob_map.h>
#undef private
#undef protected

#include "../include/topay_b200.h"

namespace {
struct QuietRog {
    std::streambuf* old;
    QuietRog() : old(std::cout.rdbuf(nullptr)) {}
    ~QuietRog() {
        std::cout.rdbuf(old);
        std::cout.clear();
    }
};
rog_map::Vec3f v3(const double* p) { return rog_map::Vec3f(p[0], p[1], p[2]); }
}  // namespace

extern "C" {

void* ref_rog_create(const topay_rog_desc* d) {
    QuietRog q;
    rog_map::ESDFMap* m = new rog_map::ESDFMap();
    m->initESDFMap(rog_map::Vec3i(d->half_prob_map_size_i[0], d->half_prob_map_size_i[1], d->half_prob_map_size_i[2]),
                   d->prob_resolution, d->esdf_resolution, v3(d->local_update_box), d->map_sliding_en != 0,
                   0.0 /* sliding_thresh as ProbMap passes its own; the ESDF layer slides with every call */,
                   v3(d->fix_map_origin), d->unk_thresh);
    return m;
}
void ref_rog_geometry(void* h, int32_t* half, int32_t* size, double* res, int32_t* origin_i, int32_t* half_box) {
    rog_map::ESDFMap* m = (rog_map::ESDFMap*)h;
    for (int i = 0; i < 3; i++) {
        half[i] = m->sc_.half_map_size_i(i);
        size[i] = m->sc_.map_size_i(i);
        origin_i[i] = m->local_map_origin_i_(i);
        half_box[i] = m->half_local_update_box_i_(i);
    }
    *res = m->sc_.resolution;
}
void ref_rog_slide(void* h, const double* odom) {
    QuietRog q;
    ((rog_map::ESDFMap*)h)->mapSliding(v3(odom));
}
void ref_rog_update_counters(void* h, const double* pos, const uint8_t* from, const uint8_t* to, int64_t n) {
    rog_map::ESDFMap* m = (rog_map::ESDFMap*)h;
    for (int64_t i = 0; i < n; i++)
        m->updateGridCounter(v3(pos + 3 * i), (rog_map::GridType)from[i], (rog_map::GridType)to[i]);
}
void ref_rog_set_occupied_cnt(void* h, const int16_t* cnt) {
    rog_map::ESDFMap* m = (rog_map::ESDFMap*)h;
    std::memcpy(m->md_.occupied_cnt.data(), cnt, m->md_.occupied_cnt.size() * sizeof(int16_t));
}
void ref_rog_download_counters(void* h, int16_t* occ, int16_t* unk) {
    rog_map::ESDFMap* m = (rog_map::ESDFMap*)h;
    if (occ) std::memcpy(occ, m->md_.occupied_cnt.data(), m->md_.occupied_cnt.size() * sizeof(int16_t));
    if (unk) std::memcpy(unk, m->md_.unknown_cnt.data(), m->md_.unknown_cnt.size() * sizeof(int16_t));
}
void ref_rog_update_esdf(void* h, const double* odom) {
    QuietRog q;
    ((rog_map::ESDFMap*)h)->updateESDF3D(v3(odom));
}
void ref_rog_query(void* h, int kind, const double* pos, int64_t n, double* dist, double* grad) {
    rog_map::ESDFMap* m = (rog_map::ESDFMap*)h;
    for (int64_t i = 0; i < n; i++) {
        const Eigen::Vector3d p = v3(pos + 3 * i);
        double d = 0;
        Eigen::Vector3d g = Eigen::Vector3d::Zero();
        bool has_grad = true;
        switch (kind) {
            case TOPAY_ROG_Q_EDT: m->getValueGrad(p, d, g); break;
            case TOPAY_ROG_Q_FLAT: m->getValueGrad2d(p, d, g); break;
            case TOPAY_ROG_Q_CRITICAL: m->getCriticalValueGrad(p, d, g); break;
            case TOPAY_ROG_Q_CELL: d = m->getDistance(p); has_grad = false; break;
            case TOPAY_ROG_Q_CELL_FLAT: d = m->getDistance2d(p); has_grad = false; break;
            default: d = m->getCriticalDistance(p); has_grad = false; break;
        }
        dist[i] = d;
        if (grad && has_grad)
            for (int k = 0; k < 3; k++) grad[3 * i + k] = g[k];
    }
}
void ref_rog_evaluate_edt(void* h, const double* pos, int64_t n, double* dist, double* grad) {
    rog_map::ESDFMap* m = (rog_map::ESDFMap*)h;
    for (int64_t i = 0; i < n; i++) {
        Eigen::Vector3d g;
        m->evaluateEDT(v3(pos + 3 * i), dist[i]);
        if (grad) {
            m->evaluateFirstGrad(v3(pos + 3 * i), g);
            for (int k = 0; k < 3; k++) grad[3 * i + k] = g[k];
        }
    }
}
void ref_rog_is_line_free2d(void* h, const double* s, const double* e, int64_t n, double thr, int8_t* out) {
    rog_map::ESDFMap* m = (rog_map::ESDFMap*)h;
    for (int64_t i = 0; i < n; i++)
        out[i] = m->isLineFree2d(Eigen::Vector2d(s[2 * i], s[2 * i + 1]), Eigen::Vector2d(e[2 * i], e[2 * i + 1]), thr) ? 1 : 0;
}
void ref_rog_download(void* h, int which, double* out) {
    rog_map::ESDFMap* m = (rog_map::ESDFMap*)h;
    const std::vector<double>& v = which == TOPAY_ROG_BUF_DIST3 ? m->distance_buffer
                                 : which == TOPAY_ROG_BUF_NEG3 ? m->tmp_buffer1_
                                 : which == TOPAY_ROG_BUF_CRITICAL ? m->distance_buffer_2d : m->distance_buffer_flat;
    std::memcpy(out, v.data(), v.size() * sizeof(double));
}


// ---- rog_map::ProbMap itself. initProbMap() carries a function-level `static bool init_once` (one ProbMap per
// process), so the driver performs ITS steps (prob_map.cpp:25-88) on the public members instead of calling it;
// everything after construction is the reference's own code. Config's constructor reads rosparams: the fields it
// would set are filled here with config.hpp:160-262's arithmetic (resetMapSize() is the reference's).
struct RefProb {
    std::shared_ptr<rog_map::ProbMap> pm;
};
void* ref_prob_create(const topay_rog_desc* d, const topay_prob_desc* p) {
    QuietRog q;
    using namespace rog_map;
    RefProb* r = new RefProb();
    r->pm = std::make_shared<ProbMap>();
    ProbMap& m = *r->pm;
    Config& c = m.cfg_;
    c.resolution = d->prob_resolution;
    c.inflation_resolution = p->inflation_resolution;
    c.inflation_step = p->inflation_step;
    c.unk_inflation_en = false;
    c.esdf_en = true;
    c.esdf_resolution = d->esdf_resolution;
    c.esdf_local_update_box = v3(d->local_update_box);
    c.frontier_extraction_en = false;
    c.raycasting_en = p->raycasting_en != 0;
    c.map_sliding_en = d->map_sliding_en != 0;
    c.map_sliding_thresh = p->map_sliding_thresh;
    c.fix_map_origin = v3(d->fix_map_origin);
    c.intensity_thresh = p->intensity_thresh;
    c.point_filt_num = p->point_filt_num <= 0 ? 1 : p->point_filt_num;
    c.batch_update_size = p->batch_update_size <= 0 ? 1 : p->batch_update_size;
    c.unk_thresh = d->unk_thresh;
    c.p_hit = p->p_hit; c.p_miss = p->p_miss; c.p_min = p->p_min; c.p_max = p->p_max; c.p_occ = p->p_occ; c.p_free = p->p_free;
    c.raycast_range_min = p->raycast_range_min;
    c.raycast_range_max = p->raycast_range_max;
    c.sqr_raycast_range_max = c.raycast_range_max * c.raycast_range_max;
    c.sqr_raycast_range_min = c.raycast_range_min * c.raycast_range_min;
    c.local_update_box_d = v3(p->local_update_box);
    c.virtual_ground_height = p->virtual_ground_height;
    c.virtual_ceil_height = p->virtual_ceil_height;
    c.visualization_range = Vec3f(0, 0, 0);
    // (resetMapSize() derives the map sizes from map_size_d; they are overridden with the descriptor's below)
    c.map_size_d = Vec3f((2.0 * d->half_prob_map_size_i[0] + 1.0) * c.resolution, (2.0 * d->half_prob_map_size_i[1] + 1.0) * c.resolution,
                         (2.0 * d->half_prob_map_size_i[2] + 1.0) * c.resolution);
    c.resetMapSize();
    c.half_map_size_i = Vec3i(d->half_prob_map_size_i[0], d->half_prob_map_size_i[1], d->half_prob_map_size_i[2]);
    c.inf_half_map_size_i = c.half_map_size_i / (int)std::ceil(p->inflation_resolution / d->prob_resolution) + (p->inflation_step + 1) * Vec3i::Ones();
#define logit(x) (log((x) / (1 - (x))))
    c.l_hit = logit(c.p_hit); c.l_miss = logit(c.p_miss); c.l_min = logit(c.p_min);
    c.l_max = logit(c.p_max); c.l_occ = logit(c.p_occ); c.l_free = logit(c.p_free);
#undef logit
    c.spherical_neighbor.clear();
    for (int dx = -c.inflation_step; dx <= c.inflation_step; dx++)
        for (int dy = -c.inflation_step; dy <= c.inflation_step; dy++)
            for (int dz = -c.inflation_step; dz <= c.inflation_step; dz++)
                if (c.inflation_step == 1 || dx * dx + dy * dy + dz * dz <= c.inflation_step * c.inflation_step)
                    c.spherical_neighbor.emplace_back(dx, dy, dz);
    // prob_map.cpp:25-88 (initProbMap) without its once-per-process guard
    m.initSlidingMap(c.half_map_size_i, c.resolution, c.map_sliding_en, c.map_sliding_thresh, c.fix_map_origin);
    m.time_consuming_.resize(7);
    m.inf_map_ = std::make_shared<InfMap>(c);
    m.esdf_map_ = std::make_shared<ESDFMap>();
    m.esdf_map_->initESDFMap(c.half_map_size_i, c.resolution, c.esdf_resolution, c.esdf_local_update_box, c.map_sliding_en,
                             c.map_sliding_thresh, c.fix_map_origin, c.unk_thresh);
    m.posToGlobalIndex(c.visualization_range, m.sc_.visualization_range_i);
    m.posToGlobalIndex(c.virtual_ceil_height, m.sc_.virtual_ceil_height_id_g);
    m.posToGlobalIndex(c.virtual_ground_height, m.sc_.virtual_ground_height_id_g);
    c.virtual_ceil_height = m.sc_.virtual_ceil_height_id_g * c.resolution;
    c.virtual_ground_height = m.sc_.virtual_ground_height_id_g * c.resolution;
    if (!c.map_sliding_en) m.slideAllMap(c.fix_map_origin);
    const int map_size = m.sc_.map_size_i.prod();
    m.occupancy_buffer_.resize(map_size, 0);
    m.raycast_data_.raycaster.setResolution(c.resolution);
    m.raycast_data_.operation_cnt.resize(map_size, 0);
    m.raycast_data_.hit_cnt.resize(map_size, 0);
    m.resetLocalMap();
    return r;
}
void ref_prob_destroy(void* h) { delete (RefProb*)h; }
void* ref_prob_esdf(void* h) { return ((RefProb*)h)->pm->esdf_map_.get(); }   // for the ref_rog_* downloads
void ref_prob_update(void* h, const float* cloud, int64_t n, const double* pos) {
    QuietRog q;
    rog_map::PointCloud pc;
    for (int64_t i = 0; i < n; i++) {
        rog_map::PclPoint pt;
        pt.x = cloud[4 * i];
        pt.y = cloud[4 * i + 1];
        pt.z = cloud[4 * i + 2];
        pt.intensity = cloud[4 * i + 3];
        pc.push_back(pt);
    }
    rog_map::Pose pose;
    pose.first = v3(pos);
    pose.second = Eigen::Quaterniond(1, 0, 0, 0);
    ((RefProb*)h)->pm->updateProbMap(pc, pose);
}
void ref_prob_download(void* h, float* occ, int32_t* origin_i) {
    rog_map::ProbMap& m = *((RefProb*)h)->pm;
    std::memcpy(occ, m.occupancy_buffer_.data(), m.occupancy_buffer_.size() * sizeof(float));
    for (int i = 0; i < 3; i++) origin_i[i] = m.local_map_origin_i_(i);
}
void ref_prob_size(void* h, int32_t* size) {
    for (int i = 0; i < 3; i++) size[i] = ((RefProb*)h)->pm->sc_.map_size_i(i);
}
}  // extern "C"
extern "C" void ref_prob_debug(void* h, double* out) {
    rog_map::ProbMap& m = *((RefProb*)h)->pm;
    for (int i = 0; i < 3; i++) {
        out[i] = m.raycast_data_.local_update_box_min(i);
        out[3 + i] = m.raycast_data_.local_update_box_max(i);
        out[6 + i] = m.local_map_bound_min_d_(i);
        out[9 + i] = m.local_map_bound_max_d_(i);
        out[12 + i] = m.cfg_.half_local_update_box_i(i);
    }
    out[15] = m.cfg_.virtual_ceil_height; out[16] = m.cfg_.virtual_ground_height; out[17] = m.cfg_.sqr_raycast_range_max;
    out[18] = m.cfg_.raycast_range_min; out[19] = m.cfg_.l_hit;
}
