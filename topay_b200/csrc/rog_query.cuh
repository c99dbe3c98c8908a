// ROG-Map ring-buffer lookups — the index arithmetic of rog_map::SlidingMap
// (src/rog_map/src/rog_map/sliding_map.cpp:168-218, ORIGIN_AT_CORNER) and the queries of
// rog_map::ESDFMap (src/rog_map/src/rog_map/esdf_map.cpp:78-152, 903-1097). No in-map test: a
// position outside the local map wraps onto the ring, as in the reference.
#pragma once
#include "field_query.cuh"

// products that feed a branch are kept un-contracted (no FMA) so that the walk below takes the
// same turns as the host arithmetic of the reference
#if defined(__CUDA_ARCH__)
#define TP_MUL_RN(a, b) __dmul_rn((a), (b))
#define TP_ADD_RN(a, b) __dadd_rn((a), (b))
#else
#define TP_MUL_RN(a, b) ((a) * (b))
#define TP_ADD_RN(a, b) ((a) + (b))
#endif

// posToGlobalIndex (sliding_map.cpp:175-184)
TP_HD int tp_rog_cell(const TpRog& r, double p) { return (int)floor(p * r.res_inv); }
// globalIndexToPos (sliding_map.cpp:195-203)
TP_HD double tp_rog_centre(const TpRog& r, int id) { return ((double)id + 0.5) * r.res; }
// globalIndexToLocalIndex + half (sliding_map.cpp:205-218, 168-173): ring coordinate in [0, size)
TP_HD int tp_rog_ring(const TpRog& r, int id_g, int a) {
    int v = id_g % r.size[a];
    if (v > r.half[a]) v -= r.size[a];
    else if (v < -r.half[a]) v += r.size[a];
    return v + r.half[a];
}
TP_HD size_t tp_rog_hash3(const TpRog& r, int gx, int gy, int gz) {
    return ((size_t)tp_rog_ring(r, gx, 0) * r.size[1] + tp_rog_ring(r, gy, 1)) * r.size[2] + tp_rog_ring(r, gz, 2);
}
TP_HD size_t tp_rog_hash2(const TpRog& r, int gx, int gy) {
    return (size_t)tp_rog_ring(r, gx, 0) * r.size[1] + tp_rog_ring(r, gy, 1);
}
// the reference looks a tap up through its cell-centre position (getDistance(pts[x][y][z]))
TP_HD int tp_rog_recell(const TpRog& r, int id) { return tp_rog_cell(r, tp_rog_centre(r, id)); }

// getSurroundPts anchor (esdf_map.cpp:903-925); shift = 1 on the interpolated axes, 0 on z for the 2-D kinds
TP_HD void tp_rog_anchor(const TpRog& r, double pos, double shift, int& idx, double& diff) {
    const double pm = pos - 0.5 * r.res * shift;
    idx = tp_rog_cell(r, pm);
    diff = (pos - tp_rog_centre(r, idx)) / r.res;
}

// evaluateEDT + evaluateFirstGrad = getValueGrad (esdf_map.cpp:951-1003). grad may be null.
TP_HD void tp_rog_value_grad(const TpRog& r, const double* pos, double& dist, double* grad) {
    int ix, iy, iz;
    double f0, f1, f2;
    tp_rog_anchor(r, pos[0], 1.0, ix, f0);
    tp_rog_anchor(r, pos[1], 1.0, iy, f1);
    tp_rog_anchor(r, pos[2], 1.0, iz, f2);
    double d[2][2][2];
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 2; y++)
#pragma unroll
            for (int z = 0; z < 2; z++)
                d[x][y][z] = TP_LDG(r.dist3 + tp_rog_hash3(r, tp_rog_recell(r, ix + x), tp_rog_recell(r, iy + y),
                                                            tp_rog_recell(r, iz + z)));
    const double v00 = (1 - f0) * d[0][0][0] + f0 * d[1][0][0];
    const double v01 = (1 - f0) * d[0][0][1] + f0 * d[1][0][1];
    const double v10 = (1 - f0) * d[0][1][0] + f0 * d[1][1][0];
    const double v11 = (1 - f0) * d[0][1][1] + f0 * d[1][1][1];
    const double v0 = (1 - f1) * v00 + f1 * v10;
    const double v1 = (1 - f1) * v01 + f1 * v11;
    dist = (1 - f2) * v0 + f2 * v1;
    if (grad) {
        grad[2] = (v1 - v0) * r.res_inv;
        grad[1] = ((1 - f2) * (v10 - v00) + f2 * (v11 - v01)) * r.res_inv;
        double gx = (1 - f2) * (1 - f1) * (d[1][0][0] - d[0][0][0]);
        gx += (1 - f2) * f1 * (d[1][1][0] - d[0][1][0]);
        gx += f2 * (1 - f1) * (d[1][0][1] - d[0][0][1]);
        gx += f2 * f1 * (d[1][1][1] - d[0][1][1]);
        grad[0] = gx * r.res_inv;
    }
}

// getCriticalValueGrad / getValueGrad2d (esdf_map.cpp:1005-1097) on `buf`; grad[2] = 0.
TP_HD void tp_rog_value_grad2d(const TpRog& r, const double* buf, const double* pos, double& dist, double* grad) {
    int ix, iy;
    double f0, f1;
    tp_rog_anchor(r, pos[0], 1.0, ix, f0);
    tp_rog_anchor(r, pos[1], 1.0, iy, f1);
    double d[2][2];
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 2; y++)
            d[x][y] = TP_LDG(buf + tp_rog_hash2(r, tp_rog_recell(r, ix + x), tp_rog_recell(r, iy + y)));
    const double fxy1 = f0 * d[1][0] + (1 - f0) * d[0][0];
    const double fxy2 = f0 * d[1][1] + (1 - f0) * d[0][1];
    dist = (1 - f1) * fxy1 + f1 * fxy2;
    if (grad) {
        grad[0] = ((1 - f1) * (d[1][0] - d[0][0]) + f1 * (d[1][1] - d[0][1])) * r.res_inv;
        grad[1] = (-fxy1 + fxy2) * r.res_inv;
        grad[2] = 0.0;
    }
}

// nearest-cell getters (esdf_map.cpp:78-120)
TP_HD double tp_rog_cell3(const TpRog& r, const double* pos) {
    return TP_LDG(r.dist3 + tp_rog_hash3(r, tp_rog_cell(r, pos[0]), tp_rog_cell(r, pos[1]), tp_rog_cell(r, pos[2])));
}
TP_HD double tp_rog_cell2(const TpRog& r, const double* buf, const double* pos) {
    return TP_LDG(buf + tp_rog_hash2(r, tp_rog_cell(r, pos[0]), tp_rog_cell(r, pos[1])));
}

// ESDFMap::isLineFree2d (esdf_map.cpp:122-152) stepping as rog_map::raycaster::RayCaster
// (include/utils/raycaster.cpp:66-192) with z = 0: visits the cells from the one holding `start`
// up to, not including, the one holding `end`. The walk is capped at max_steps cells (a ray
// that misses its end cell through rounding never terminates in the reference).
TP_HD bool tp_rog_line_free2d(const TpRog& r, const double* s, const double* e, double threshold, int max_steps) {
    const double BIG = 1.7976931348623157e308;
    int cur[2], ei[2], dir[2];
    double t_step[2] = {BIG, BIG}, t_bound[2] = {BIG, BIG};
    for (int i = 0; i < 2; i++) {
        cur[i] = (int)floor(s[i] / r.res);
        ei[i] = (int)floor(e[i] / r.res);
        const int dlt = ei[i] - cur[i];
        dir[i] = (0 < dlt) - (dlt < 0);
    }
    if (dir[0] != 0 || dir[1] != 0) {
        double dd[2] = {fabs(e[0] - s[0]), fabs(e[1] - s[1])};
        const double tmax = sqrt(TP_ADD_RN(TP_MUL_RN(dd[0], dd[0]), TP_MUL_RN(dd[1], dd[1])));
        for (int i = 0; i < 2; i++) {
            dd[i] /= tmax;
            t_step[i] = dir[i] == 0 ? BIG : fabs(r.res / dd[i]);
            const double nb = TP_ADD_RN(TP_MUL_RN((double)cur[i] + 0.5, r.res), TP_MUL_RN(TP_MUL_RN((double)dir[i], r.res), 0.5));
            t_bound[i] = dir[i] == 0 ? BIG : fabs(nb - s[i]) / dd[i];
        }
    }
    for (int it = 0; it < max_steps; it++) {
        if (cur[0] == ei[0] && cur[1] == ei[1]) return true;
        const int cx = tp_rog_recell(r, cur[0]), cy = tp_rog_recell(r, cur[1]);
        // the z axis never moves (t_bound_z = max): x steps only when strictly the nearest bound
        if (t_bound[0] < t_bound[1]) {
            cur[0] += dir[0];
            t_bound[0] += t_step[0];
        } else {
            cur[1] += dir[1];
            t_bound[1] += t_step[1];
        }
        if (TP_LDG(r.flat + tp_rog_hash2(r, cx, cy)) < threshold) return false;
    }
    return true;
}

// ---- GridMap's queries with the use_rog switch (src/map/include/map/grid_map.h), as the solver and the
// feasibility check call them: getDisWithGradI2d(pos, d, g) with the default flags (:364-392 ->
// getValueGrad2d, no inflation), getDisWithGradI3d (:443-461 -> evaluateEDT + evaluateFirstGrad),
// getDistance2d (:256-267) and getDistance3d (:307-322). No in-map test on the ROG side.
TP_HD void tp_field_query2d_flat(const TpGrid& g, const double* xy, double& d, double* grad) {
    if (g.kind == 1) {
        const double p[3] = {xy[0], xy[1], 0.0};
        double g3[3];
        tp_rog_value_grad2d(g.rog, g.rog.flat, p, d, g3);
        grad[0] = g3[0];
        grad[1] = g3[1];
    } else {
        tp_query2d(g, g.esdf2d, xy, d, grad);
    }
}
TP_HD void tp_field_query3d(const TpGrid& g, const double* p, double& d, double* grad) {
    if (g.kind == 1) tp_rog_value_grad(g.rog, p, d, grad);
    else tp_query3d(g, p, d, grad);
}
TP_HD double tp_field_distance2d(const TpGrid& g, const double* xy) {
    if (g.kind == 1) {
        const double p[3] = {xy[0], xy[1], 0.0};
        double d;
        tp_rog_value_grad2d(g.rog, g.rog.flat, p, d, nullptr);
        return d;
    }
    return tp_distance2d(g, xy);
}
TP_HD double tp_field_distance3d(const TpGrid& g, const double* p) {
    if (g.kind == 1) {
        double d;
        tp_rog_value_grad(g.rog, p, d, nullptr);
        return d;
    }
    return tp_distance3d(g, p);
}

#if defined(__CUDACC__)
// 16-bit counter += delta by CAS on the enclosing 32-bit word (CounterMap's int16 counters, counter_map.h:100-110)
__device__ __forceinline__ void tp_atomic_add16(int16_t* base, size_t idx, int delta) {
    unsigned int* w = reinterpret_cast<unsigned int*>(base) + (idx >> 1);
    const int sh = (idx & 1) ? 16 : 0;
    unsigned int old = *w, assumed;
    do {
        assumed = old;
        const unsigned int cur = (assumed >> sh) & 0xffffu;
        const unsigned int nxt = (cur + (unsigned int)delta) & 0xffffu;
        old = atomicCAS(w, assumed, (assumed & ~(0xffffu << sh)) | (nxt << sh));
    } while (old != assumed);
}
#endif
