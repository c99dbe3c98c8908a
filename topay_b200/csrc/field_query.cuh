// Dense-grid distance lookups — GridMap::getDisWithGradI2d/3d and getDistance2d/3d
// (src/map/include/map/grid_map.h:256-509) with the index helpers of :727-868.
// Addressing: 3-D x*Ny*Nz + y*Nz + z (grid_map.h:808-816); 2-D x*Ny + y (:798-806).
#pragma once
#include "hd.cuh"

#if defined(__CUDA_ARCH__)
#define TP_LDG(p) __ldg(p)
#else
#define TP_LDG(p) (*(p))
#endif

TP_HD bool tp_in_map2(const TpGrid& g, const double* p) {
    if (p[0] < g.min_boundary[0] + 1e-4 || p[1] < g.min_boundary[1] + 1e-4) return false;
    if (p[0] > g.max_boundary[0] - 1e-4 || p[1] > g.max_boundary[1] - 1e-4) return false;
    return true;
}
TP_HD bool tp_in_map3(const TpGrid& g, const double* p) {
    if (p[0] < g.min_boundary[0] + 1e-4 || p[1] < g.min_boundary[1] + 1e-4 || p[2] < g.min_boundary[2] + 1e-4)
        return false;
    if (p[0] > g.max_boundary[0] - 1e-4 || p[1] > g.max_boundary[1] - 1e-4 || p[2] > g.max_boundary[2] - 1e-4)
        return false;
    return true;
}
TP_HD int tp_clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// Anchor cell + fractional offsets of the interpolation (grid_map.h:472-480).
TP_HD void tp_anchor(const TpGrid& g, double pos, int axis, int& idx, double& diff) {
    const double pm = pos - 0.5 * g.resolution;
    idx = (int)floor((pm - g.origin[axis]) * g.resolution_inv);
    const double centre = (idx + 0.5) * g.resolution + g.origin[axis];
    diff = (pos - centre) * g.resolution_inv;
}

// getDisWithGradI3d (grid_map.h:443-509). grad may be null.
TP_HD void tp_query3d(const TpGrid& g, const double* pos, double& distance, double* grad) {
    if (!tp_in_map3(g, pos)) {
        distance = 0.0;
        if (grad) grad[0] = grad[1] = grad[2] = 0.0;
        return;
    }
    int ix, iy, iz;
    double dx, dy, dz;
    tp_anchor(g, pos[0], 0, ix, dx);
    tp_anchor(g, pos[1], 1, iy, dy);
    tp_anchor(g, pos[2], 2, iz, dz);
    const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
    const int x0 = tp_clampi(ix, nx - 1), x1 = tp_clampi(ix + 1, nx - 1);
    const int y0 = tp_clampi(iy, ny - 1), y1 = tp_clampi(iy + 1, ny - 1);
    const int z0 = tp_clampi(iz, nz - 1), z1 = tp_clampi(iz + 1, nz - 1);
    const double* b = g.esdf3d;
    const size_t sx = (size_t)ny * nz, sy = (size_t)nz;
    const double v000 = TP_LDG(b + x0 * sx + y0 * sy + z0), v001 = TP_LDG(b + x0 * sx + y0 * sy + z1);
    const double v010 = TP_LDG(b + x0 * sx + y1 * sy + z0), v011 = TP_LDG(b + x0 * sx + y1 * sy + z1);
    const double v100 = TP_LDG(b + x1 * sx + y0 * sy + z0), v101 = TP_LDG(b + x1 * sx + y0 * sy + z1);
    const double v110 = TP_LDG(b + x1 * sx + y1 * sy + z0), v111 = TP_LDG(b + x1 * sx + y1 * sy + z1);
    const double v00 = v000 * (1 - dx) + v100 * dx;
    const double v01 = v001 * (1 - dx) + v101 * dx;
    const double v10 = v010 * (1 - dx) + v110 * dx;
    const double v11 = v011 * (1 - dx) + v111 * dx;
    const double v0 = v00 * (1 - dy) + v10 * dy;
    const double v1 = v01 * (1 - dy) + v11 * dy;
    distance = v0 * (1.0 - dz) + v1 * dz;
    if (grad) {
        grad[2] = (v1 - v0) * g.resolution_inv;
        grad[1] = ((v10 - v00) * (1.0 - dz) + (v11 - v01) * dz) * g.resolution_inv;
        double gx = (1.0 - dz) * (1 - dy) * (v100 - v000);
        gx += (1.0 - dz) * dy * (v110 - v010);
        gx += dz * (1 - dy) * (v101 - v001);
        gx += dz * dy * (v111 - v011);
        grad[0] = gx * g.resolution_inv;
    }
}

// The same lookup in two steps for callers that need the gradient only when a hinge on the value is active (the
// sphere penalties of the stage-2 node: a sphere is almost always clear of its margin): tp_query3d_taps fetches the
// eight taps and returns the value, tp_query3d_grad finishes the gradient from them. Same operations in the same
// order as tp_query3d, so the two paths are bit-identical.
struct TpTaps3 {
    double v000, v001, v010, v011, v100, v101, v110, v111, dx, dy, dz;
    double v00, v01, v10, v11, v0, v1;
    bool in_map;
};
TP_HD double tp_query3d_taps(const TpGrid& g, const double* pos, TpTaps3& t) {
    t.in_map = tp_in_map3(g, pos);
    if (!t.in_map) return 0.0;
    int ix, iy, iz;
    tp_anchor(g, pos[0], 0, ix, t.dx);
    tp_anchor(g, pos[1], 1, iy, t.dy);
    tp_anchor(g, pos[2], 2, iz, t.dz);
    const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
    const int x0 = tp_clampi(ix, nx - 1), x1 = tp_clampi(ix + 1, nx - 1);
    const int y0 = tp_clampi(iy, ny - 1), y1 = tp_clampi(iy + 1, ny - 1);
    const int z0 = tp_clampi(iz, nz - 1), z1 = tp_clampi(iz + 1, nz - 1);
    const double* b = g.esdf3d;
    const size_t sx = (size_t)ny * nz, sy = (size_t)nz;
    t.v000 = TP_LDG(b + x0 * sx + y0 * sy + z0); t.v001 = TP_LDG(b + x0 * sx + y0 * sy + z1);
    t.v010 = TP_LDG(b + x0 * sx + y1 * sy + z0); t.v011 = TP_LDG(b + x0 * sx + y1 * sy + z1);
    t.v100 = TP_LDG(b + x1 * sx + y0 * sy + z0); t.v101 = TP_LDG(b + x1 * sx + y0 * sy + z1);
    t.v110 = TP_LDG(b + x1 * sx + y1 * sy + z0); t.v111 = TP_LDG(b + x1 * sx + y1 * sy + z1);
    t.v00 = t.v000 * (1 - t.dx) + t.v100 * t.dx;
    t.v01 = t.v001 * (1 - t.dx) + t.v101 * t.dx;
    t.v10 = t.v010 * (1 - t.dx) + t.v110 * t.dx;
    t.v11 = t.v011 * (1 - t.dx) + t.v111 * t.dx;
    t.v0 = t.v00 * (1 - t.dy) + t.v10 * t.dy;
    t.v1 = t.v01 * (1 - t.dy) + t.v11 * t.dy;
    return t.v0 * (1.0 - t.dz) + t.v1 * t.dz;
}
TP_HD void tp_query3d_grad(const TpGrid& g, const TpTaps3& t, double* grad) {
    if (!t.in_map) {
        grad[0] = grad[1] = grad[2] = 0.0;
        return;
    }
    grad[2] = (t.v1 - t.v0) * g.resolution_inv;
    grad[1] = ((t.v10 - t.v00) * (1.0 - t.dz) + (t.v11 - t.v01) * t.dz) * g.resolution_inv;
    double gx = (1.0 - t.dz) * (1 - t.dy) * (t.v100 - t.v000);
    gx += (1.0 - t.dz) * t.dy * (t.v110 - t.v010);
    gx += t.dz * (1 - t.dy) * (t.v101 - t.v001);
    gx += t.dz * t.dy * (t.v111 - t.v011);
    grad[0] = gx * g.resolution_inv;
}

// getDistance3d (grid_map.h:307-362): value only, 1e10 outside the map.
TP_HD double tp_distance3d(const TpGrid& g, const double* pos) {
    if (!tp_in_map3(g, pos)) return 1e+10;
    double d;
    tp_query3d(g, pos, d, nullptr);
    return d;
}

// getDisWithGradI2d (grid_map.h:364-441) on one of the three 2-D buffers.
TP_HD void tp_query2d(const TpGrid& g, const double* buf, const double* pos, double& distance, double* grad) {
    if (!tp_in_map2(g, pos)) {
        distance = 0.0;
        if (grad) grad[0] = grad[1] = 0.0;
        return;
    }
    int ix, iy;
    double dx, dy;
    tp_anchor(g, pos[0], 0, ix, dx);
    tp_anchor(g, pos[1], 1, iy, dy);
    const int nx = g.dims[0], ny = g.dims[1];
    const int x0 = tp_clampi(ix, nx - 1), x1 = tp_clampi(ix + 1, nx - 1);
    const int y0 = tp_clampi(iy, ny - 1), y1 = tp_clampi(iy + 1, ny - 1);
    const double v00 = TP_LDG(buf + (size_t)x0 * ny + y0), v01 = TP_LDG(buf + (size_t)x0 * ny + y1);
    const double v10 = TP_LDG(buf + (size_t)x1 * ny + y0), v11 = TP_LDG(buf + (size_t)x1 * ny + y1);
    const double v0 = v00 * (1 - dx) + v10 * dx;
    const double v1 = v01 * (1 - dx) + v11 * dx;
    distance = v0 * (1 - dy) + v1 * dy;
    if (grad) {
        grad[1] = (v1 - v0) * g.resolution_inv;
        double gx = (1 - dy) * (v10 - v00);
        gx += dy * (v11 - v01);
        grad[0] = gx * g.resolution_inv;
    }
}

// getDistance2d (grid_map.h:256-305): flat map only, 1e10 outside.
TP_HD double tp_distance2d(const TpGrid& g, const double* pos) {
    if (!tp_in_map2(g, pos)) return 1e+10;
    double d;
    tp_query2d(g, g.esdf2d, pos, d, nullptr);
    return d;
}

// isCollision2d / isCollision3d (grid_map.h:511-536, 695-724): outside the map counts as collision.
TP_HD bool tp_is_collision2d(const TpGrid& g, const double* pos, double threshold) {
    if (!tp_in_map2(g, pos)) return true;
    return tp_distance2d(g, pos) < threshold;
}
TP_HD bool tp_is_collision3d(const TpGrid& g, const double* pos, double threshold) {
    if (!tp_in_map3(g, pos)) return true;
    return tp_distance3d(g, pos) < threshold;
}
// posToIndex2d (grid_map.h:727-735)
TP_HD void tp_pos_to_index2(const TpGrid& g, const double* pos, int* id) {
    id[0] = (int)floor((pos[0] - g.origin[0]) * g.resolution_inv);
    id[1] = (int)floor((pos[1] - g.origin[1]) * g.resolution_inv);
}
// getDistCoarse2i / getDistCoarse2d (grid_map.h:887-940, dense branch): nearest cell, clamped
TP_HD double tp_dist_coarse2i(const TpGrid& g, int ix, int iy, bool critical) {
    const int x = tp_clampi(ix, g.dims[0] - 1), y = tp_clampi(iy, g.dims[1] - 1);
    const double* b = critical ? g.esdf2d_critical : g.esdf2d_inflate;
    return TP_LDG(b + (size_t)x * g.dims[1] + y);
}
// isLineCollisionGrid2d (grid_map.h:565-611): Bresenham over the flat map, both end cells tested.
// A cell outside the grid counts as occupied (the reference reads past its buffer there).
TP_HD bool tp_line_collision_grid2d(const TpGrid& g, const double* p1, const double* p2, double threshold) {
    int s[2], e[2];
    tp_pos_to_index2(g, p1, s);
    tp_pos_to_index2(g, p2, e);
    const int dx = abs(e[0] - s[0]), dy = abs(e[1] - s[1]);
    const int sx = s[0] < e[0] ? 1 : -1, sy = s[1] < e[1] ? 1 : -1;
    int err = dx - dy, x0 = s[0], y0 = s[1];
    for (;;) {
        if (x0 < 0 || y0 < 0 || x0 > g.dims[0] - 1 || y0 > g.dims[1] - 1) return true;
        if (TP_LDG(g.esdf2d + (size_t)x0 * g.dims[1] + y0) < threshold) return true;
        if (x0 == e[0] && y0 == e[1]) break;
        const int e2 = 2 * err;
        if (e2 > -dy) { err -= dy; x0 += sx; }
        if (e2 < dx) { err += dx; y0 += sy; }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------
// Front-end visibility ray (row N2): TopologyPRM::lineVisib (src/planner/src/topo_prm.cpp:278-315) over the planner's
// RayCaster (src/planner/src/utils/raycast.cpp:27-45 signum / mod / intbound, :253-299 setInput, :301-346 step) — an
// Amanatides & Woo walk in cell units from p1 / res to p2 / res; every cell but the end cell is looked up in the
// inflated (or critical) 2-D map through getDistCoarse2i; the first cell at or below `thresh` blocks the ray and
// pc = midpoint of that cell's centre and the previous cell's (z = 0). Divisions, fmod and floor are exact on both
// sides; the two multiply-adds of indexToPos3d are kept un-contracted, so verdict and pc are bit-identical.
// The walk is capped at the Manhattan cell distance + 2 steps (a ray that steps past its end cell loops for ever in
// the reference; stated deviation).
#if defined(__CUDA_ARCH__)
#define TP_FQ_MUL(a, b) __dmul_rn((a), (b))
#define TP_FQ_ADD(a, b) __dadd_rn((a), (b))
#else
#define TP_FQ_MUL(a, b) ((a) * (b))
#define TP_FQ_ADD(a, b) ((a) + (b))
#endif
TP_HD int tp_signum(int v) { return v == 0 ? 0 : (v < 0 ? -1 : 1); }
TP_HD double tp_intbound(double s, double ds) {
    if (ds < 0) {
        s = -s;
        ds = -ds;
    }
    s = fmod(fmod(s, 1.0) + 1.0, 1.0);
    return (1 - s) / ds;
}
TP_HD bool tp_line_visib(const TpGrid& g, const double* p1, const double* p2, double thresh, bool critical, double* pc) {
    const double res = g.resolution;
    double s[3], e[3], off[2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        s[i] = p1[i] / res;
        e[i] = p2[i] / res;
    }
    off[0] = 0.5 - g.origin[0] / res;
    off[1] = 0.5 - g.origin[1] / res;
    int c[3], en[3], st[3];
    double tm[3], td[3];
    long budget = 2;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        c[i] = (int)floor(s[i]);
        en[i] = (int)floor(e[i]);
        const double d = en[i] - c[i];
        st[i] = tp_signum((int)d);
        tm[i] = tp_intbound(s[i], d);
        td[i] = ((double)st[i]) / d;
        budget += en[i] > c[i] ? en[i] - c[i] : c[i] - en[i];
    }
    if (st[0] == 0 && st[1] == 0 && st[2] == 0) return true;
    int prev[2] = {c[0], c[1]};
    while (budget-- > 0) {
        if (c[0] == en[0] && c[1] == en[1] && c[2] == en[2]) return true;
        const int id[2] = {(int)(c[0] + off[0]), (int)(c[1] + off[1])};     // double -> int truncation, as the reference
        if (tp_dist_coarse2i(g, id[0], id[1], critical) <= thresh) {
            const double ax = TP_FQ_ADD(TP_FQ_MUL(id[0] + 0.5, res), g.origin[0]);
            const double ay = TP_FQ_ADD(TP_FQ_MUL(id[1] + 0.5, res), g.origin[1]);
            const double bx = TP_FQ_ADD(TP_FQ_MUL(prev[0] + 0.5, res), g.origin[0]);
            const double by = TP_FQ_ADD(TP_FQ_MUL(prev[1] + 0.5, res), g.origin[1]);
            pc[0] = TP_FQ_MUL(0.5, TP_FQ_ADD(ax, bx));
            pc[1] = TP_FQ_MUL(0.5, TP_FQ_ADD(ay, by));
            pc[2] = 0.0;
            return false;
        }
        prev[0] = id[0];
        prev[1] = id[1];
        if (tm[0] < tm[1]) {
            if (tm[0] < tm[2]) { c[0] += st[0]; tm[0] += td[0]; } else { c[2] += st[2]; tm[2] += td[2]; }
        } else {
            if (tm[1] < tm[2]) { c[1] += st[1]; tm[1] += td[1]; } else { c[2] += st[2]; tm[2] += td[2]; }
        }
    }
    return true;
}
