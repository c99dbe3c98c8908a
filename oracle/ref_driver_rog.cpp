// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// C entry points over the REFERENCE'S OWN ROG-Map ESDF layer, compiled unmodified from where it lies under
// /root/reference (oracle/Makefile target `_ref`):
//   src/rog_map/src/rog_map/esdf_map.cpp     ESDFMap::updateESDF3D + ring fillESDF, the six lookups, isLineFree2d
//   src/rog_map/src/rog_map/counter_map.cpp  CounterMap::updateGridCounter, initCounterMap
//   src/rog_map/src/rog_map/sliding_map.cpp  SlidingMap index maths, mapSliding
//   src/rog_map/src/utils/raycaster.cpp      the DDA isLineFree2d walks
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

}  // extern "C"
