// topay_field: the dense occupancy / ESDF grid of the reference's GridMap in HBM.
//
//   rasterise   GridMap::regenerateMap ingest         src/map/src/grid_map.cpp:733-747
//   rebuild     GridMap::updateESDF + fillESDF        src/map/src/grid_map.cpp:89-123, 125-521
//   queries     getDisWithGradI2d/3d, getDistance2d/3d, isWholeBodyCollision
//                                                      src/map/include/map/grid_map.h:256-509, 613-650
//
// The distance transform is exact and integer: each separable pass computes
// min_v (q - v)^2 + f(v) over a line, which is what the reference's lower-envelope
// routine evaluates (its intermediates are exact integers in double), so
// the squared-distance grids are bit-identical; the final `res * sqrt(v)` and the
// sign combine are evaluated with explicit round-to-nearest intrinsics (no FMA
// contraction) so the fp64 grids are bit-identical too.
//
// Layout (as the reference): 3-D x*Ny*Nz + y*Nz + z, 2-D x*Ny + y.
// Pass 1 runs along the contiguous axis from the binary occupancy (warp scans, positive and
// negative transform together, one sign-packed int16 per cell); passes 2/3 run along a strided
// axis on tiles of 16 contiguous cells x the whole line staged in shared memory, each line
// answered by a divide and conquer over its queries (edt_line.cuh); every global access is a
// 64-128 B contiguous segment and a voxel moves 1 + 2 | 2 + 4 | 4 + 8 = 21 B through the passes.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "common_host.h"
#include "edt_line.cuh"
#include "field_query.cuh"
#include "robot.cuh"

static thread_local std::string g_last_error;
void tp_set_error(const std::string& msg) { g_last_error = msg; }
extern "C" const char* topay_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* topay_version(void) { return "topay_b200 0.1 (sm_100a)"; }
extern "C" const char* topay_strerror(int code) {
    switch (code) {
        case TOPAY_OK: return "ok";
        case TOPAY_ERR_INVALID_ARG: return "invalid argument";
        case TOPAY_ERR_NO_DEVICE: return "no usable CUDA device (the B200 path has no CPU fallback)";
        case TOPAY_ERR_CUDA: return "CUDA runtime error";
        case TOPAY_ERR_ALLOC: return "device allocation failed";
        case TOPAY_ERR_TOO_LARGE: return "problem larger than the configured capacity";
        case TOPAY_ERR_NOT_READY: return "field not built";
        default: return "unknown error";
    }
}

int tp_require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        tp_set_error("no usable CUDA device: libtopay_b200 runs on the GPU only");
        return TOPAY_ERR_NO_DEVICE;
    }
    return TOPAY_OK;
}

struct topay_field {
    topay_grid_desc desc;
    int device;
    cudaStream_t stream;
    TpGrid grid;            // device pointers + geometry
    int nx, ny, nz;
    size_t n2, n3;
    int8_t *occ3d, *occ2d, *occ2d_crit, *src2d;
    double *esdf3d, *esdf2d, *esdf2d_inflate, *esdf2d_crit;
    int16_t* packed16;      // pass-1 output, sign-packed 1-D distances (3-D sized)
    int32_t* packed32;      // pass-2 output, sign-packed squared distances (3-D sized)
    int32_t *sq_pos[4], *sq_neg[4];
    bool keep_sq;
    bool ready;
    cudaEvent_t ev0, ev1, ev2;
    float ms_total, ms_3d;
};

int tp_field_device(const topay_field* f) { return f->device; }
bool tp_field_ready(const topay_field* f) { return f->ready; }
void tp_field_grid(const topay_field* f, TpGrid* out) { *out = f->grid; }

// ------------------------------------------------------------------ kernels

// grid_map.cpp:733-747: one thread per float32 point.
__global__ void k_rasterize(const float* __restrict__ xyz, int64_t n, TpGrid g, double chassis_height,
                            int8_t* occ3d, int8_t* occ2d, int8_t* occ2d_crit) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
    const int ix = (int)floor(((double)px - g.origin[0]) * g.resolution_inv);
    const int iy = (int)floor(((double)py - g.origin[1]) * g.resolution_inv);
    const int iz = (int)floor(((double)pz - g.origin[2]) * g.resolution_inv);
    if (ix >= 0 && iy >= 0 && ix <= nx - 1 && iy <= ny - 1) {
        occ2d_crit[(size_t)ix * ny + iy] = 1;
        if ((double)pz < chassis_height) occ2d[(size_t)ix * ny + iy] = 1;
        if (iz >= 0 && iz <= nz - 1) occ3d[((size_t)ix * ny + iy) * nz + iz] = 1;
    }
}

// inflated maps: source = previous ESDF below the chassis radius (grid_map.cpp:288, 360)
__global__ void k_threshold(const double* __restrict__ esdf, double thr, int8_t* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = esdf[i] < thr ? 1 : 0;
}

// Pass 1, along the contiguous axis: 1-D distance to the nearest source (pos) and to the nearest
// non-source (neg) cell of each line, sign-packed into one int16 (edt_line.cuh; TP_INF16 = none on this line). One warp
// per line, four consecutive cells per lane (one 4-byte load, one 8-byte store when C % 4 == 0): the index of
// the last source at or before a cell is a running max inside the lane, a warp max-scan across
// lanes and a carry across the 128-cell chunks; the next source at or after it is the mirror image.
// Lines of up to 128 cells (the z axis of the 3-D grid) take a single chunk, everything in
// registers; longer lines (the 2-D maps) make a forward and a backward pass over their chunks.
#define TP_NEG_BIG (-(1 << 28))
#define TP_POS_BIG (1 << 28)

__device__ __forceinline__ int tp_scan_max_excl(int v, int lane) {
    // exclusive max-scan over lower lanes; identity TP_NEG_BIG
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x = max(x, y);
    }
    const int e = __shfl_up_sync(0xffffffffu, x, 1);
    return lane == 0 ? TP_NEG_BIG : e;
}
__device__ __forceinline__ int tp_scan_min_excl_rev(int v, int lane) {
    // exclusive min-scan over higher lanes; identity TP_POS_BIG
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_down_sync(0xffffffffu, x, d);
        if (lane + d < 32) x = min(x, y);
    }
    const int e = __shfl_down_sync(0xffffffffu, x, 1);
    return lane == 31 ? TP_POS_BIG : e;
}

template <bool VEC>
__global__ void __launch_bounds__(256)
k_edt_contig(const int8_t* __restrict__ src, int16_t* __restrict__ out, int64_t n_lines, int C) {
    const int lane = threadIdx.x & 31;
    const int64_t line = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (line >= n_lines) return;
    const int8_t* in = src + line * C;
    int16_t* o = out + line * C;
    const int nchunk = (C + 127) >> 7;
    auto load4 = [&](int c0, bool occ[4]) {
        if (VEC) {
            uchar4 v = make_uchar4(0, 0, 0, 0);
            if (c0 < C) v = *reinterpret_cast<const uchar4*>(in + c0);
            occ[0] = v.x == 1; occ[1] = v.y == 1; occ[2] = v.z == 1; occ[3] = v.w == 1;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) occ[k] = (c0 + k < C) && in[c0 + k] == 1;
        }
    };
    // forward: last source / last free cell at or before each cell
    int carry_p = TP_NEG_BIG, carry_n = TP_NEG_BIG;
    int fp[4], fn[4];           // kept for the single-chunk case
    for (int ch = 0; ch < nchunk; ch++) {
        const int c0 = (ch << 7) + lane * 4;
        bool occ[4];
        load4(c0, occ);
        int lp = TP_NEG_BIG, ln = TP_NEG_BIG, lastp[4], lastn[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool in_line = c0 + k < C;
            if (in_line && occ[k]) lp = c0 + k;
            if (in_line && !occ[k]) ln = c0 + k;
            lastp[k] = lp;
            lastn[k] = ln;
        }
        const int ep = max(tp_scan_max_excl(lp, lane), carry_p);
        const int en = max(tp_scan_max_excl(ln, lane), carry_n);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = c0 + k;
            const int a = max(lastp[k], ep), b = max(lastn[k], en);
            fp[k] = a < 0 ? TP_INF16 : min(c - a, TP_INF16);
            fn[k] = b < 0 ? TP_INF16 : min(c - b, TP_INF16);
        }
        carry_p = max(carry_p, __shfl_sync(0xffffffffu, max(lp, ep), 31));
        carry_n = max(carry_n, __shfl_sync(0xffffffffu, max(ln, en), 31));
        if (nchunk > 1) {   // the forward result of the transform the cell is not a source of, sign-packed
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (c0 + k < C) o[c0 + k] = (int16_t)(occ[k] ? -fn[k] : fp[k]);
        }
    }
    // backward: next source / next free cell at or after each cell, min with the forward result
    carry_p = TP_POS_BIG;
    carry_n = TP_POS_BIG;
    for (int ch = nchunk - 1; ch >= 0; ch--) {
        const int c0 = (ch << 7) + lane * 4;
        bool occ[4];
        load4(c0, occ);
        int np_ = TP_POS_BIG, nn_ = TP_POS_BIG, nextp[4], nextn[4];
#pragma unroll
        for (int k = 3; k >= 0; k--) {
            const bool in_line = c0 + k < C;
            if (in_line && occ[k]) np_ = c0 + k;
            if (in_line && !occ[k]) nn_ = c0 + k;
            nextp[k] = np_;
            nextn[k] = nn_;
        }
        const int ep = min(tp_scan_min_excl_rev(np_, lane), carry_p);
        const int en = min(tp_scan_min_excl_rev(nn_, lane), carry_n);
        int16_t res[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = c0 + k;
            const int a = min(nextp[k], ep), b = min(nextn[k], en);
            int bp = a >= TP_POS_BIG ? TP_INF16 : min(a - c, TP_INF16);
            int bn = b >= TP_POS_BIG ? TP_INF16 : min(b - c, TP_INF16);
            if (nchunk > 1) {
                if (c < C) {
                    const int f = o[c];
                    if (occ[k]) bn = min(bn, -f);
                    else bp = min(bp, f);
                }
            } else {
                bp = min(bp, fp[k]);
                bn = min(bn, fn[k]);
            }
            res[k] = (int16_t)(occ[k] ? -bn : bp);   // sign-packed (edt_line.cuh)
        }
        carry_p = min(carry_p, __shfl_sync(0xffffffffu, min(np_, ep), 0));
        carry_n = min(carry_n, __shfl_sync(0xffffffffu, min(nn_, en), 0));
        if (VEC) {
            if (c0 < C) *reinterpret_cast<uint2*>(o + c0) = make_uint2((uint16_t)res[0] | ((uint32_t)(uint16_t)res[1] << 16),
                                                                       (uint16_t)res[2] | ((uint32_t)(uint16_t)res[3] << 16));
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (c0 + k < C) o[c0 + k] = res[k];
        }
    }
}

__device__ __forceinline__ void tp_edt_store(bool FINAL, int bp, int bn, size_t idx, double res,
                                             int32_t* __restrict__ out_pos, int32_t* __restrict__ out_neg,
                                             double* __restrict__ esdf, const TpRogSink& rs, int c_line, int c_outer,
                                             int c_inner) {
    if (bp >= TP_INF32) bp = FINAL ? INT32_MAX : TP_INF32;
    if (bn >= TP_INF32) bn = FINAL ? INT32_MAX : TP_INF32;
    if (FINAL) {
        // grid_map.cpp:457, 503, 515-517 — round-to-nearest, no FMA contraction
        // exactly one of the two transforms is non-zero at a cell (it is a source of the other one): one square root
        const bool occ = bn > 0;
        const int x = occ ? bn : bp;
        const double r = __dmul_rn(res, __dsqrt_rn(x == INT32_MAX ? DBL_MAX : (double)x));
        const double dp = occ ? 0.0 : r;        // res * sqrt(0) = +0
        const double dn = occ ? r : 0.0;
        if (rs.enabled) {
            // ROG ring: box coordinates -> ring memory, q > mem_end ? q + id_l - S : q + id_l per axis
            size_t m;
            if (rs.dims3) {
                // last 3-D pass: line axis = x, outer = y, inner = z
                const int qx = c_line + rs.lo[0], qy = c_outer + rs.lo[1], qz = c_inner + rs.lo[2];
                const int mx = qx > rs.mem_end[0] ? qx + rs.idl[0] - rs.size[0] : qx + rs.idl[0];
                const int my = qy > rs.mem_end[1] ? qy + rs.idl[1] - rs.size[1] : qy + rs.idl[1];
                const int mz = qz > rs.mem_end[2] ? qz + rs.idl[2] - rs.size[2] : qz + rs.idl[2];
                m = ((size_t)mx * rs.size[1] + my) * rs.size[2] + mz;
            } else {
                // last 2-D pass: line axis = x, inner = y
                const int qx = c_line + rs.lo[0], qy = c_inner + rs.lo[1];
                const int mx = qx > rs.mem_end[0] ? qx + rs.idl[0] - rs.size[0] : qx + rs.idl[0];
                const int my = qy > rs.mem_end[1] ? qy + rs.idl[1] - rs.size[1] : qy + rs.idl[1];
                m = (size_t)mx * rs.size[1] + my;
            }
            double d = dp;
            if (rs.fuse && dn > 0.0) d = __dadd_rn(d, __dadd_rn(-dn, res));
            rs.dist[m] = d;
            rs.neg[m] = dn;
            return;
        }
        double d = dp;
        if (dn > 0.0) d = __dadd_rn(d, __dadd_rn(-dn, res));
        if (esdf) esdf[idx] = d;
        if (out_pos) {
            out_pos[idx] = bp;
            out_neg[idx] = bn;
        }
    } else {
        out_pos[idx] = bp;
        out_neg[idx] = bn;
    }
}

// Passes 2 and 3 (and the last pass of the 2-D maps), along a strided axis. Block = one outer index x TZ
// contiguous inner cells; the whole line of the tile is staged in shared memory as sign-packed squared distances
// (edt_line.cuh), TZ x 2-4 B contiguous per row on the way in, TZ x 4 B (intermediate) or TZ x 8 B (ESDF) on the
// way out. Thread (x, y): column x of the tile, chunk y (y + TY, ...) of TP_EDT_CHUNK consecutive cells.
//   1. runs of "no source" cells get their skip distances (two walks over the thread's own chunk, the carries of
//      the neighbouring chunks through first / last tables);
//   2. the chunk boundaries are answered against the whole line (window cut-off + skips);
//   3. every thread answers the inside of its chunk by halving between known minimisers — no barrier, the
//      minimisers (int16) of a chunk are private to its thread.
// The negative transform of each cell is the short outward search; both results leave the block the moment they
// are known. The 16 threads of a row write 64 / 128 contiguous bytes.
// IN16: input = pass 1's sign-packed int16 1-D distances, else pass 2's sign-packed int32 squared distances.
template <bool IN16, bool FINAL, int TZ>
__global__ void k_edt_line(const void* __restrict__ in_, int32_t* __restrict__ out32, int32_t* __restrict__ out_pos,
                           int32_t* __restrict__ out_neg, double* __restrict__ esdf, int n_line, size_t line_stride,
                           size_t outer_stride, int n_inner, int tiles, int CH, double res,
                           const __grid_constant__ TpRogSink sink) {
    extern __shared__ __align__(16) int sm_i[];
    const int TY = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
    const int nseg = (n_line + CH - 1) / CH, nb = tp_edt_boundaries(n_line, CH);
    int* s_f = sm_i;                                                        // [n_line][TZ]
    short* s_arg = reinterpret_cast<short*>(s_f + n_line * TZ);     // [n_line][TZ]
    short* s_first = s_arg + n_line * TZ;                           // [nseg][TZ]
    short* s_last = s_first + nseg * TZ;                            // [nseg][TZ]
    const int outer = blockIdx.x / tiles, tile = blockIdx.x % tiles;
    const int c = tile * TZ + tx;
    const bool cvalid = c < n_inner;
    const size_t base = (size_t)outer * outer_stride + c;
    for (int l = ty; l < n_line; l += TY) {
        int v = TP_INF32;
        if (cvalid) {
            const size_t idx = base + (size_t)l * line_stride;
            v = IN16 ? tp_sq16(static_cast<const int16_t*>(in_)[idx]) : static_cast<const int32_t*>(in_)[idx];
        }
        s_f[l * TZ + tx] = v;
    }
    __syncthreads();
    int* col = s_f + tx;
    // 1. skip distances of the "no source" runs
    uint32_t has_run = 0u;      // bit i: the i-th chunk of this thread holds a "no source" cell
    if (cvalid) {
        int it = 0;
        for (int sg = ty; sg < nseg; sg += TY, it++) {
            const int v0 = sg * CH, v1 = min(v0 + CH, n_line);
            int first = -1, last = -1, cnt = 0;
            for (int v = v0; v < v1; v++)
                if (col[v * TZ] < TP_INF32) {
                    if (first < 0) first = v;
                    last = v;
                    cnt++;
                }
            s_first[sg * TZ + tx] = (short)first;
            s_last[sg * TZ + tx] = (short)last;
            if (cnt != v1 - v0 || it >= 32) has_run |= 1u << (it & 31);
        }
    }
    __syncthreads();
    if (cvalid && has_run != 0u) {
        int it = 0;
        for (int sg = ty; sg < nseg; sg += TY, it++) {
            if (!((has_run >> (it & 31)) & 1u)) continue;
            const int v0 = sg * CH, v1 = min(v0 + CH, n_line);
            int lastc = -1, nextc = n_line;     // nearest cell with a value before / after the chunk
            for (int k = sg - 1; k >= 0 && lastc < 0; k--) lastc = s_last[k * TZ + tx];
            for (int k = sg + 1; k < nseg && nextc >= n_line; k++) {
                const int fk = s_first[k * TZ + tx];
                if (fk >= 0) nextc = fk;
            }
            for (int v = v1 - 1; v >= v0; v--) {
                if (col[v * TZ] < TP_INF32) nextc = v;
                else col[v * TZ] = tp_skip_make(0, nextc - v);
            }
            for (int v = v0; v < v1; v++) {
                const int fv = col[v * TZ];
                if (fv < TP_INF32) lastc = v;
                else col[v * TZ] = fv | ((v - lastc) << TP_SKIP_BITS);
            }
        }
    }
    __syncthreads();
    auto emit = [&](int l, int bp) {
        const int bn = tp_neg_search<TZ>(col, n_line, l);
        const size_t idx = base + (size_t)l * line_stride;
        if (FINAL) {
            tp_edt_store(true, bp, bn, idx, res, out_pos, out_neg, esdf, sink, l, outer, c);
        } else {
            // exactly one of the two is non-zero (the cell is a source of the other transform)
            out32[idx] = bn > 0 ? -(bn >= TP_INF32 ? TP_INF32 : bn) : (bp >= TP_INF32 ? TP_INF32 : bp);
        }
    };
    // 2. chunk boundaries, each against the whole line
    if (cvalid)
        for (int b = ty; b < nb; b += TY) {
            const int q = tp_edt_boundary(b, n_line, CH);
            int arg;
            const int bp = tp_dc_query<TZ>(col, q, 0, n_line - 1, arg);
            s_arg[q * TZ + tx] = (short)arg;
            emit(q, bp);
        }
    __syncthreads();
    // 3. the inside of the chunks, one thread per chunk
    if (cvalid)
        for (int ch = ty; ch < nb - 1; ch += TY) {
            const int lo = tp_edt_boundary(ch, n_line, CH), hi = tp_edt_boundary(ch + 1, n_line, CH);
            for (int h = tp_dc_top(hi - lo) >> 1; h >= 1; h >>= 1)
                for (int q = lo + h; q < hi; q += 2 * h) {
                    const int qr = q + h < hi ? q + h : hi;
                    const int a_lo = s_arg[(q - h) * TZ + tx], a_hi = s_arg[qr * TZ + tx];
                    int arg;
                    const int bp = tp_dc_query<TZ>(col, q, a_lo, a_hi, arg);
                    s_arg[q * TZ + tx] = (short)arg;
                    emit(q, bp);
                }
        }
}

// Pass 1 on large grids with short contiguous lines (the z axis of the 3-D grid: C = 16 NV <= 128 cells): one THREAD
// per line — 16-byte loads of the occupancy, folded into a 128-bit occupancy mask; the nearest cell of the other kind
// before a cell is a running index, the nearest one after it a find-first-set on the shifted mask; 16-byte stores of
// the sign-packed int16 distances. A warp per line (k_edt_contig) spends two 5-step warp scans on an 80-cell line.
template <int NV>
__global__ void __launch_bounds__(128)
k_edt_contig_thread(const int8_t* __restrict__ src, int16_t* __restrict__ out, int64_t n_lines) {
    constexpr int C = 16 * NV;
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines) return;
    const uint4* in = reinterpret_cast<const uint4*>(src + line * C);
    unsigned long long occ_lo = 0ull, occ_hi = 0ull;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const uint4 v = in[k];
        const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
        unsigned long long bits = 0ull;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // byte == 1 -> 0xff (per-byte compare), its top bits gathered into a nibble
            const uint32_t eq = __vcmpeq4(ww[j], 0x01010101u) & 0x80808080u;
            bits |= (unsigned long long)((eq * 0x00204081u) >> 28) << (4 * j);
        }
        if (k < 4) occ_lo |= bits << (16 * k);
        else occ_hi |= bits << (16 * (k - 4));
    }
    constexpr unsigned long long V_LO = C >= 64 ? ~0ull : (1ull << (C & 63)) - 1ull;
    constexpr unsigned long long V_HI = C <= 64 ? 0ull : (C >= 128 ? ~0ull : (1ull << ((C - 64) & 63)) - 1ull);
    const unsigned long long free_lo = ~occ_lo & V_LO, free_hi = ~occ_hi & V_HI;
    uint4* dst = reinterpret_cast<uint4*>(out + line * C);
    int lastp = -2 * TP_INF16, lastn = -2 * TP_INF16;
    uint32_t o[4];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const bool occ = c < 64 ? (occ_lo >> c) & 1ull : (occ_hi >> (c - 64)) & 1ull;
        if (occ) lastp = c; else lastn = c;
        // the other kind: free cells for an occupied cell, occupied ones for a free cell
        const unsigned long long lo = occ ? free_lo : occ_lo, hi = occ ? free_hi : occ_hi;
        int nxt = 2 * TP_INF16 + C;
        if (c < 64) {
            const unsigned long long x = lo >> c;
            if (x) nxt = c + __ffsll((long long)x) - 1;
            else if (C > 64 && hi) nxt = 64 + __ffsll((long long)hi) - 1;
        } else {
            const unsigned long long x = hi >> (c - 64);
            if (x) nxt = c + __ffsll((long long)x) - 1;
        }
        const int m = min(min(c - (occ ? lastn : lastp), nxt - c), TP_INF16);
        const uint32_t v = (uint32_t)(uint16_t)(int16_t)(occ ? -m : m);      // sign-packed (edt_line.cuh)
        if (c & 1) o[(c >> 1) & 3] |= v << 16; else o[(c >> 1) & 3] = v;
        if ((c & 7) == 7) dst[c >> 3] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// Passes 2 and 3 on large grids: the reference's own O(n) lower-envelope sweep in its integer two-scan form
// (edt_line.cuh, tp_env_*), NS threads per line. Block = 32 adjacent lines (lanes = consecutive inner cells, so every
// global access of a warp is one contiguous 64-256 B segment) x NS warps; warp j owns the sources of segment j of the
// line (L = ceil(n / NS) rounded up to 32 cells).
//   1. each thread builds the envelope of ITS segment's sources over the whole line (forward scan of L cells, input
//      fetched 16 cells ahead of the stack work; the stack is a per-thread local array of L entries, lane-interleaved by
//      the hardware) and records the segment's occupancy bits in shared memory;
//   2. the line is swept from its far end in chunks of 32 cells: every thread reads its envelope off at the chunk's
//      cells into shared memory, then the block's threads share the chunk's 32 x 32 outputs: minimum over the NS
//      envelopes = the positive transform of a free cell; an occupied cell takes the short outward search of the
//      negative transform instead (its sources, the free cells, are dense). Results leave the block at once.
// One thread per line would do with a fifth of the instructions per cell of the divide and conquer (k_edt_line), but an
// 800-cell line is then a 50 k-instruction dependency chain on 14 warps per SM; NS = 4 cuts the chain to a third and
// quadruples the warps.
__device__ __forceinline__ void tp_touch(const void* p) {
    unsigned a, b;
    asm volatile("ld.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
}

template <bool IN16, bool FINAL, int NS, int MAXSEG>
__global__ void __launch_bounds__(32 * NS)
k_edt_scan(const void* __restrict__ in_, int32_t* __restrict__ out32, int32_t* __restrict__ out_pos,
           int32_t* __restrict__ out_neg, double* __restrict__ esdf, int n_line, int L, size_t line_stride,
           size_t outer_stride, int n_inner, int64_t n_lines, double res, const __grid_constant__ TpRogSink sink) {
    extern __shared__ __align__(16) int sm_scan[];
    const int lane = threadIdx.x, j = threadIdx.y;
    const int nw = (n_line + 31) >> 5;
    int* s_val = sm_scan;                                              // [NS][32 cells][32 lanes]
    unsigned* s_occ = reinterpret_cast<unsigned*>(sm_scan + NS * 1024);   // [nw][32 lanes]
    int64_t t = (int64_t)blockIdx.x * 32 + lane;
    const bool valid = t < n_lines;
    if (!valid) t = n_lines - 1;        // a lane beyond the grid repeats the last line and writes nothing
    const int outer = (int)(t / n_inner), c = (int)(t % n_inner);
    const size_t base = (size_t)outer * outer_stride + c;
    TpEnvEntry stk[MAXSEG];
    // raw cell of the line, index clamped (no predicate on the load: the fetches of a chunk issue back to back and are
    // converted only when the scan reaches them)
    auto raw = [&](int u) -> int {
        const size_t idx = base + (size_t)min(max(u, 0), n_line - 1) * line_stride;
        return IN16 ? (int)static_cast<const int16_t*>(in_)[idx] : static_cast<const int32_t*>(in_)[idx];
    };
    auto cvt = [&](int r) -> int { return IN16 ? tp_sq16(r) : r; };
    constexpr int PF = 16;
    // ---- 1. envelope of the segment's sources
    const int s0 = j * L, s1 = min(s0 + L, n_line);
    TpEnv e{-1, 0, 0, 0};
    if (s0 < s1) {
        int nx[PF], nn[PF];
        int prev = s0 > 0 ? tp_env_val<false>(cvt(raw(s0 - 1))) : 1;
#pragma unroll
        for (int k = 0; k < PF; k++) nx[k] = raw(s0 + k);
        int cur_raw = cvt(nx[0]);
        unsigned bits = 0u;
        for (int u0 = s0; u0 < s1; u0 += PF) {
#pragma unroll
            for (int k = 0; k < PF; k++) nn[k] = raw(u0 + PF + k);
#pragma unroll
            for (int k = 0; k < PF; k++) {
                const int u = u0 + k;
                if (u < s1) {
                    bits |= (cur_raw < 0 ? 1u : 0u) << (u & 31);
                    const int next_raw = cvt(k + 1 < PF ? nx[k + 1] : nn[0]);
                    const int nxt = u + 1 < n_line ? tp_env_val<false>(next_raw) : 1;
                    const int cur = tp_env_val<false>(cur_raw);
                    tp_env_push(e, stk, n_line, u, cur, prev, nxt);
                    prev = cur;
                    cur_raw = next_raw;
                }
            }
            if (((u0 + PF) & 31) == 0 || u0 + PF >= s1) {
                s_occ[(u0 >> 5) * 32 + lane] = bits;
                bits = 0u;
            }
#pragma unroll
            for (int k = 0; k < PF; k++) nx[k] = nn[k];
        }
    }
#pragma unroll
    for (int k = 1; k <= 8; k++) tp_touch(&stk[max(e.q - k, 0)]);
    __syncthreads();
    // ---- 2. sweep from the far end, 32 cells at a time
    const double r_inf = __dmul_rn(res, __dsqrt_rn(DBL_MAX));      // "no obstacle anywhere" (fillESDF's DBL_MAX)
    if (e.q < 0) {          // no source in the segment: a sentinel entry that never pops and never wins
        e.sv = 0;
        e.gv = TP_INF32;
        e.tv = -1;
    }
    for (int ub = (nw - 1) << 5; ub >= 0; ub -= 32) {
        int* sv_out = s_val + (j * 32 + 31) * 32 + lane;
#pragma unroll 8
        for (int u = ub + 31; u >= ub; u--, sv_out -= 32) {
            const int d = u - e.sv;
            *sv_out = d * d + e.gv;         // < 2^28 + 2^29; cells past the line's end are never read
            if (u == e.tv && e.q > 0) {     // the bottom entry (t = 0) stays
                e.q--;
                tp_env_load(e, stk);
                tp_touch(&stk[max(e.q - 8, 0)]);
            }
        }
        __syncthreads();
        const unsigned occ = s_occ[(ub >> 5) * 32 + lane];
        for (int k = j; k < 32; k += NS) {
            const int u = ub + k;
            if (u >= n_line || !valid) continue;
            int bp = s_val[k * 32 + lane], bn = 0;
#pragma unroll
            for (int jj = 1; jj < NS; jj++) bp = min(bp, s_val[(jj * 32 + k) * 32 + lane]);
            bp = min(bp, TP_INF32);
            if ((occ >> k) & 1u) {
                // negative transform: outward search, nearest cells first, ended by the window cut-off d^2 >= best
                bp = 0;
                bn = tp_env_val<true>(cvt(raw(u)));
                for (int d = 1; d < n_line; d++) {
                    const int dd = d * d;
                    if (dd >= bn) break;
                    if (u - d >= 0) bn = min(bn, dd + tp_env_val<true>(cvt(raw(u - d))));
                    if (u + d < n_line) bn = min(bn, dd + tp_env_val<true>(cvt(raw(u + d))));
                }
            }
            const size_t idx = base + (size_t)u * line_stride;
            if (!FINAL) {
                out32[idx] = bn > 0 ? -bn : bp;
            } else if (sink.enabled) {
                tp_edt_store(true, bp, bn, idx, res, out_pos, out_neg, esdf, sink, u, outer, c);
            } else {
                // the dense field's store (tp_edt_store without the ring sink): one square root per cell, rounded
                // products and sums (grid_map.cpp:457, 503, 515-517)
                const bool oc = bn > 0;
                const int x = oc ? bn : bp;
                const double r = x >= TP_INF32 ? r_inf : __dmul_rn(res, __dsqrt_rn((double)x));
                if (esdf) esdf[idx] = oc ? __dadd_rn(-r, res) : r;
                if (out_pos) {
                    out_pos[idx] = bp >= TP_INF32 ? INT32_MAX : bp;
                    out_neg[idx] = bn >= TP_INF32 ? INT32_MAX : bn;
                }
            }
        }
        __syncthreads();
    }
}

__global__ void k_query3d(TpGrid g, const double* __restrict__ pos, int64_t n, double* dist, double* grad,
                          int value_only) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double p[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    if (value_only) {
        dist[i] = tp_distance3d(g, p);
        return;
    }
    double d, gr[3];
    tp_query3d(g, p, d, gr);
    dist[i] = d;
    if (grad) {
        grad[3 * i] = gr[0];
        grad[3 * i + 1] = gr[1];
        grad[3 * i + 2] = gr[2];
    }
}

__global__ void k_query2d(TpGrid g, const double* __restrict__ buf, const double* __restrict__ pos, int64_t n,
                          double* dist, double* grad, int value_only) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double p[2] = {pos[2 * i], pos[2 * i + 1]};
    if (value_only) {
        dist[i] = tp_distance2d(g, p);
        return;
    }
    double d, gr[2];
    tp_query2d(g, buf, p, d, gr);
    dist[i] = d;
    if (grad) {
        grad[2 * i] = gr[0];
        grad[2 * i + 1] = gr[1];
    }
}

// Batched predicates and nearest-cell lookups the front-end and the feasibility check use
// (grid_map.h:511-611, 695-724, 887-940). kind: 0 isCollision2d, 1 isCollision3d,
// 2 isLineCollisionGrid2d (a = p1, b = p2), 3 getDistCoarse2d, 4 getDistCoarse2i (ia = indices).
__global__ void k_field_misc(TpGrid g, int kind, const double* __restrict__ a, const double* __restrict__ b,
                             const int32_t* __restrict__ ia, int64_t n, double threshold, int critical, int8_t* flag,
                             double* val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (kind) {
        case 0: {
            const double p[2] = {a[2 * i], a[2 * i + 1]};
            flag[i] = tp_is_collision2d(g, p, threshold) ? 1 : 0;
            break;
        }
        case 1: {
            const double p[3] = {a[3 * i], a[3 * i + 1], a[3 * i + 2]};
            flag[i] = tp_is_collision3d(g, p, threshold) ? 1 : 0;
            break;
        }
        case 2: {
            const double p[2] = {a[2 * i], a[2 * i + 1]}, q[2] = {b[2 * i], b[2 * i + 1]};
            flag[i] = tp_line_collision_grid2d(g, p, q, threshold) ? 1 : 0;
            break;
        }
        case 3: {
            const double p[2] = {a[2 * i], a[2 * i + 1]};
            int id[2];
            tp_pos_to_index2(g, p, id);
            val[i] = tp_dist_coarse2i(g, id[0], id[1], critical != 0);
            break;
        }
        default: val[i] = tp_dist_coarse2i(g, ia[2 * i], ia[2 * i + 1], critical != 0); break;
    }
}

// TopologyPRM::lineVisib (topo_prm.cpp:278-315), one thread per segment.
__global__ void k_line_visib(TpGrid g, const double* __restrict__ p1, const double* __restrict__ p2, int64_t n,
                             double thresh, int critical, int8_t* visible, double* pc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a[3] = {p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]}, b[3] = {p2[3 * i], p2[3 * i + 1], p2[3 * i + 2]};
    double c[3] = {pc[3 * i], pc[3 * i + 1], pc[3 * i + 2]};
    visible[i] = tp_line_visib(g, a, b, thresh, critical != 0, c) ? 1 : 0;
    pc[3 * i] = c[0];
    pc[3 * i + 1] = c[1];
    pc[3 * i + 2] = c[2];
}

// GridMap::isWholeBodyCollision (grid_map.h:613-650), one thread per 10-D state.
__global__ void k_whole_body(TpGrid g, TpParams P, const double* __restrict__ states, int64_t n, int8_t* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* s = states + 10 * i;
    const topay_robot_params& rp = P.robot;
    bool hit = false;
    for (int q = 0; q < TOPAY_DOF; q++)
        if (s[3 + q] > rp.joint_pos_limit_max[q] || s[3 + q] < -rp.joint_pos_limit_max[q]) hit = true;
    if (!hit) {
        const double p2[2] = {s[0], s[1]};
        if (!tp_in_map2(g, p2) || tp_distance2d(g, p2) < rp.chassis_colli_radius) hit = true;
    }
    if (!hit) {
        double pos[10];
        for (int k = 0; k < 10; k++) pos[k] = s[k];
        TpFK fk;
        TpSphereStoreLocal pts;
        tp_fk(P, pos, fk, pts);
        for (int a = 0; a < P.n_sphere && !hit; a++) {
            const double r = P.sphere_r[a];
            const double pa[3] = {pts.at(a, 0), pts.at(a, 1), pts.at(a, 2)};
            if (!tp_in_map3(g, pa) || tp_distance3d(g, pa) < r) hit = true;
            if (a > 2 && pa[2] < rp.chassis_height + r) {
                const double dx = pa[0] - s[0], dy = pa[1] - s[1];
                if (sqrt(dx * dx + dy * dy) < rp.chassis_colli_radius + r) hit = true;
            }
            for (int b = a + 1; b < P.n_sphere; b++) {
                const double dx = pa[0] - pts.at(b, 0), dy = pa[1] - pts.at(b, 1), dz = pa[2] - pts.at(b, 2);
                if (sqrt(dx * dx + dy * dy + dz * dz) < r + P.sphere_r[b] && ((P.pair_mask[a] >> b) & 1u)) hit = true;
            }
        }
    }
    out[i] = hit ? 1 : 0;
}

// ------------------------------------------------------------------ host

namespace {

template <typename T>
int fmalloc(T** p, size_t count) {
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) {
        tp_set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        *p = nullptr;
        return TOPAY_ERR_ALLOC;
    }
    return TOPAY_OK;
}

// One signed transform over a [A][B][C] array (C contiguous): pass 1 along C, pass 2 along B
// and, when A > 1, pass 3 along A. The last pass writes `esdf` (+ the integer grids).
int signed_edt(topay_field* f, const int8_t* src, int A, int B, int C, double* esdf, int32_t* sqp, int32_t* sqn) {
    TpEdtScratch sc{f->stream, f->packed16, f->packed32, f->keep_sq, f->desc.resolution, TpRogSink{}};
    return tp_signed_edt(sc, src, A, B, C, esdf, sqp, sqn);
}

}  // namespace

// One signed transform over a [A][B][C] array; shared with rogfield.cu. esdf may be NULL (integer
// grids only, keep_sq must then be set).
int tp_signed_edt(const TpEdtScratch& f_, const int8_t* src, int A, int B, int C, double* esdf, int32_t* sqp,
                  int32_t* sqn) {
    const TpEdtScratch* f = &f_;
    cudaStream_t q = f->stream;
    const int64_t n_lines = (int64_t)A * B;
    // Large grids take the thread-per-line kernels (k_edt_contig_thread, k_edt_scan): their lines alone fill the
    // machine. Small ones (the 2-D maps, the default 200 x 200 x 16 field) keep the kernels that parallelise inside a
    // line. Both are exact, so the choice never shows in the result. TOPAY_EDT_SCAN = 0 / 1 forces it (tests).
    const char* env_scan = getenv("TOPAY_EDT_SCAN");
    const int scan_mode = env_scan ? atoi(env_scan) : -1;
    auto per_thread = [&](int64_t lines) { return scan_mode >= 0 ? scan_mode != 0 : lines >= 16384; };
    // pass 1
    if (C % 16 == 0 && C <= 128 && per_thread(n_lines)) {
        const unsigned blocks = (unsigned)((n_lines + 127) / 128);
        switch (C / 16) {
#define TP_P1(NV) case NV: k_edt_contig_thread<NV><<<blocks, 128, 0, q>>>(src, f->packed16, n_lines); break;
            TP_P1(1) TP_P1(2) TP_P1(3) TP_P1(4) TP_P1(5) TP_P1(6) TP_P1(7) TP_P1(8)
#undef TP_P1
        }
    } else {
        const int wpb = 8;
        const unsigned blocks = (unsigned)((n_lines + wpb - 1) / wpb);
        if (C % 4 == 0)
            k_edt_contig<true><<<blocks, wpb * 32, 0, q>>>(src, f->packed16, n_lines, C);
        else
            k_edt_contig<false><<<blocks, wpb * 32, 0, q>>>(src, f->packed16, n_lines, C);
    }
    auto strided = [&](bool in16, bool fin, int n_line, size_t line_stride, int n_outer, size_t outer_stride,
                       int n_inner) -> int {
        const int64_t nl = (int64_t)n_outer * n_inner;
        // Measured on 800 x 800 x 80 (B200): pass 3 (every cell of a line carries a value) 0.84 ms by the envelope scan
        // against 1.07 ms by the divide and conquer; pass 2 (values only under obstacles) 0.61 against 0.46 ms.
        if (n_line <= 1024 && (scan_mode >= 0 ? scan_mode != 0 : (!in16 && nl >= 16384))) {
            int32_t* op = fin && f->keep_sq ? sqp : nullptr;
            int32_t* on = fin && f->keep_sq ? sqn : nullptr;
            static const int env_ns = getenv("TOPAY_EDT_NS") ? atoi(getenv("TOPAY_EDT_NS")) : 0;   // dev
            const int NS = env_ns == 2 ? 2 : 4;
            const int L = (((n_line + NS - 1) / NS) + 31) & ~31;       // <= 256 (NS = 4) / 512 (NS = 2)
            const unsigned blocks = (unsigned)((nl + 31) / 32);
            const size_t smem = (size_t)NS * 1024 * 4 + (size_t)((n_line + 31) / 32) * 32 * 4;
            dim3 blk(32, NS);
#define TP_SCAN(IN, FIN, INP, OUT)                                                                                  \
            do {                                                                                                    \
                if (NS == 4)                                                                                        \
                    k_edt_scan<IN, FIN, 4, 256><<<blocks, blk, smem, q>>>(INP, OUT, op, on, fin ? esdf : nullptr, n_line,  \
                        L, line_stride, outer_stride, n_inner, nl, f->res, f->sink);                                \
                else                                                                                                \
                    k_edt_scan<IN, FIN, 2, 512><<<blocks, blk, smem, q>>>(INP, OUT, op, on, fin ? esdf : nullptr, n_line,  \
                        L, line_stride, outer_stride, n_inner, nl, f->res, f->sink);                                \
            } while (0)
            if (in16 && fin) TP_SCAN(true, true, f->packed16, nullptr);
            else if (in16) TP_SCAN(true, false, f->packed16, f->packed32);
            else TP_SCAN(false, true, f->packed32, nullptr);
#undef TP_SCAN
            return TOPAY_OK;
        }
        // Chunk length: 32 queries per thread on large grids; small grids (the 2-D maps, the default 200 x 200 x 16
        // field) take shorter chunks so that the chunk threads of the whole grid still fill the machine.
        static const int env_ch = getenv("TOPAY_EDT_CH") ? atoi(getenv("TOPAY_EDT_CH")) : 0;
        int CH = env_ch >= 4 ? env_ch : 32;
        while (CH > 4 && (long long)n_outer * n_inner * ((n_line + CH - 1) / CH) < 148 * 1024) CH >>= 1;
        // staged line (int32) + the minimisers (int16) + first / last tables of the chunks (int16)
        const int nseg = (n_line + CH - 1) / CH, nb = tp_edt_boundaries(n_line, CH);
        auto smem_of = [&](int tz) -> size_t {
            return (size_t)n_line * tz * 4 + (size_t)n_line * tz * 2 + (size_t)2 * nseg * tz * 2 + 16;
        };
        // Tile width: 16 cells (64 B rows in, 128 B rows of ESDF out) while at least two blocks fit an SM;
        // TOPAY_EDT_TZ / TOPAY_EDT_TY override (dev).
        static const int env_tz = getenv("TOPAY_EDT_TZ") ? atoi(getenv("TOPAY_EDT_TZ")) : 0;
        static const int env_ty = getenv("TOPAY_EDT_TY") ? atoi(getenv("TOPAY_EDT_TY")) : 0;
        int TZ = (env_tz == 4 || env_tz == 8 || env_tz == 16) ? env_tz : 16;
        while (TZ > 4 && smem_of(TZ) > (size_t)100 * 1024) TZ >>= 1;
        // small grids: narrower tiles so that the blocks cover all SMs
        while (env_tz <= 0 && TZ > 4 && (long long)n_outer * ((n_inner + TZ - 1) / TZ) < 296) TZ >>= 1;
        const size_t smem = smem_of(TZ);
        if (smem > 200 * 1024) {
            tp_set_error("grid line too long for the strided-pass staging buffer");
            return TOPAY_ERR_TOO_LARGE;
        }
        int TY = env_ty > 0 ? env_ty : std::max(1, std::min(1024 / TZ, nb));    // one thread per chunk boundary
        while ((TZ * TY) % 32 != 0 && TY < 1024 / TZ) TY++;     // whole warps
        const int tiles = (n_inner + TZ - 1) / TZ;
        dim3 blk(TZ, TY);
        const unsigned blocks = (unsigned)n_outer * tiles;
        int32_t* op = fin && f->keep_sq ? sqp : nullptr;
        int32_t* on = fin && f->keep_sq ? sqn : nullptr;
        const double res = f->res;
        // the attribute is per kernel and process-wide: it only ever grows (fields of different sizes coexist)
        static std::atomic<size_t> attr[9];
        auto launch = [&](auto kern, int slot, const void* in, int32_t* out32) -> int {
            size_t cur = attr[slot].load();
            while (cur < smem && !attr[slot].compare_exchange_weak(cur, smem)) {}
            TP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attr[slot].load()), {});
            kern<<<blocks, blk, smem, q>>>(in, out32, op, on, fin ? esdf : nullptr, n_line, line_stride, outer_stride,
                                          n_inner, tiles, CH, res, f->sink);
            return TOPAY_OK;
        };
        const int tzi = TZ == 16 ? 0 : (TZ == 8 ? 1 : 2);
#define TP_EDT_DISPATCH(IN, FIN, SLOT, INP, OUT)                                                        \
        (tzi == 0 ? launch(k_edt_line<IN, FIN, 16>, (SLOT) * 3 + 0, INP, OUT)                           \
                  : tzi == 1 ? launch(k_edt_line<IN, FIN, 8>, (SLOT) * 3 + 1, INP, OUT)                 \
                             : launch(k_edt_line<IN, FIN, 4>, (SLOT) * 3 + 2, INP, OUT))
        if (in16 && fin) return TP_EDT_DISPATCH(true, true, 0, f->packed16, nullptr);
        if (in16) return TP_EDT_DISPATCH(true, false, 1, f->packed16, f->packed32);
        return TP_EDT_DISPATCH(false, true, 2, f->packed32, nullptr);
#undef TP_EDT_DISPATCH
    };
    int rc;
    if (A == 1) {
        rc = strided(true, true, B, (size_t)C, 1, 0, C);
    } else {
        rc = strided(true, false, B, (size_t)C, A, (size_t)B * C, C);
        if (rc != TOPAY_OK) return rc;
        rc = strided(false, true, A, (size_t)B * C, B, (size_t)C, C);
    }
    return rc;
}

extern "C" int topay_field_create(const topay_grid_desc* desc, int device, topay_field** out) {
    if (!desc || !out || desc->resolution <= 0.0) return TOPAY_ERR_INVALID_ARG;
    int rc = tp_require_device(device);
    if (rc != TOPAY_OK) return rc;
    topay_field* f = new topay_field();
    memset(f, 0, sizeof(*f));
    f->desc = *desc;
    f->device = device;
    f->keep_sq = true;
    cudaSetDevice(device);
    tp_pool_keep(device);
    // grid_map.cpp:33-54
    TpGrid& g = f->grid;
    g.resolution = desc->resolution;
    g.resolution_inv = 1.0 / desc->resolution;
    for (int i = 0; i < 3; i++) {
        g.min_boundary[i] = -desc->map_size[i] / 2.0;
        g.max_boundary[i] = desc->map_size[i] / 2.0;
    }
    g.min_boundary[2] = 0.0;
    g.max_boundary[2] = desc->map_size[2];
    for (int i = 0; i < 3; i++) {
        g.origin[i] = g.min_boundary[i];
        g.dims[i] = (int)ceil(desc->map_size[i] / desc->resolution);
    }
    f->nx = g.dims[0];
    f->ny = g.dims[1];
    f->nz = g.dims[2];
    if (f->nx < 1 || f->ny < 1 || f->nz < 1 || f->nx >= TP_INF16 || f->ny >= TP_INF16 || f->nz >= TP_INF16) {
        delete f;
        tp_set_error("grid dimensions out of range");
        return TOPAY_ERR_INVALID_ARG;
    }
    f->n2 = (size_t)f->nx * f->ny;
    f->n3 = f->n2 * f->nz;
    TP_CUDA_OK(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking), { delete f; });
#define FA(p, n)                                  \
    if ((rc = fmalloc(&(p), (n))) != TOPAY_OK) {  \
        topay_field_destroy(f);                   \
        return rc;                                \
    }
    FA(f->occ3d, f->n3);
    FA(f->occ2d, f->n2);
    FA(f->occ2d_crit, f->n2);
    FA(f->src2d, f->n2);
    FA(f->esdf3d, f->n3);
    FA(f->esdf2d, f->n2);
    FA(f->esdf2d_inflate, f->n2);
    FA(f->esdf2d_crit, f->n2);
    FA(f->packed16, f->n3);
    FA(f->packed32, f->n3);
    for (int w = 0; w < 3; w++) {
        FA(f->sq_pos[w], f->n2);
        FA(f->sq_neg[w], f->n2);
    }
    FA(f->sq_pos[3], f->n3);
    FA(f->sq_neg[3], f->n3);
#undef FA
    cudaMemsetAsync(f->occ3d, 0, f->n3, f->stream);
    cudaMemsetAsync(f->occ2d, 0, f->n2, f->stream);
    cudaMemsetAsync(f->occ2d_crit, 0, f->n2, f->stream);
    cudaMemsetAsync(f->esdf3d, 0, f->n3 * 8, f->stream);
    cudaMemsetAsync(f->esdf2d, 0, f->n2 * 8, f->stream);
    cudaMemsetAsync(f->esdf2d_inflate, 0, f->n2 * 8, f->stream);
    cudaMemsetAsync(f->esdf2d_crit, 0, f->n2 * 8, f->stream);
    g.esdf3d = f->esdf3d;
    g.esdf2d = f->esdf2d;
    g.esdf2d_inflate = f->esdf2d_inflate;
    g.esdf2d_critical = f->esdf2d_crit;
    g.ready = 0;
    cudaEventCreate(&f->ev0);
    cudaEventCreate(&f->ev1);
    cudaEventCreate(&f->ev2);
    TP_CUDA_OK(cudaStreamSynchronize(f->stream), { topay_field_destroy(f); });
    *out = f;
    return TOPAY_OK;
}

extern "C" void topay_field_destroy(topay_field* f) {
    if (!f) return;
    cudaSetDevice(f->device);
    if (f->stream) cudaStreamSynchronize(f->stream);
    void* ptrs[] = {f->occ3d, f->occ2d, f->occ2d_crit, f->src2d, f->esdf3d, f->esdf2d, f->esdf2d_inflate,
                    f->esdf2d_crit, f->packed16, f->packed32, f->sq_pos[0], f->sq_pos[1], f->sq_pos[2],
                    f->sq_pos[3], f->sq_neg[0], f->sq_neg[1], f->sq_neg[2], f->sq_neg[3]};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (f->ev0) cudaEventDestroy(f->ev0);
    if (f->ev1) cudaEventDestroy(f->ev1);
    if (f->ev2) cudaEventDestroy(f->ev2);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
}

extern "C" int topay_field_dims(const topay_field* f, int32_t dims[3]) {
    if (!f || !dims) return TOPAY_ERR_INVALID_ARG;
    dims[0] = f->nx;
    dims[1] = f->ny;
    dims[2] = f->nz;
    return TOPAY_OK;
}

extern "C" int topay_field_set_keep_sqdist(topay_field* f, int keep) {
    if (!f) return TOPAY_ERR_INVALID_ARG;
    f->keep_sq = keep != 0;
    return TOPAY_OK;
}

extern "C" int topay_field_set_occupancy(topay_field* f, const int8_t* occ3d, const int8_t* occ2d,
                                         const int8_t* occ2d_critical) {
    if (!f) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(f->device);
    if (occ3d) TP_CUDA_OK(cudaMemcpyAsync(f->occ3d, occ3d, f->n3, cudaMemcpyHostToDevice, f->stream), {});
    if (occ2d) TP_CUDA_OK(cudaMemcpyAsync(f->occ2d, occ2d, f->n2, cudaMemcpyHostToDevice, f->stream), {});
    if (occ2d_critical)
        TP_CUDA_OK(cudaMemcpyAsync(f->occ2d_crit, occ2d_critical, f->n2, cudaMemcpyHostToDevice, f->stream), {});
    TP_CUDA_OK(cudaStreamSynchronize(f->stream), {});
    f->ready = false;
    f->grid.ready = 0;
    return TOPAY_OK;
}

extern "C" int topay_field_clear(topay_field* f, int clear_critical) {
    if (!f) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(f->device);
    cudaMemsetAsync(f->occ3d, 0, f->n3, f->stream);
    cudaMemsetAsync(f->occ2d, 0, f->n2, f->stream);
    if (clear_critical) cudaMemsetAsync(f->occ2d_crit, 0, f->n2, f->stream);
    f->ready = false;
    f->grid.ready = 0;
    return TOPAY_OK;
}

extern "C" int topay_field_rasterize_points(topay_field* f, const float* xyz, int64_t n) {
    if (!f || (!xyz && n > 0) || n < 0) return TOPAY_ERR_INVALID_ARG;
    if (n == 0) return TOPAY_OK;
    cudaSetDevice(f->device);
    float* d = nullptr;
    TP_CUDA_OK(cudaMallocAsync(&d, (size_t)n * 3 * sizeof(float), f->stream), {});
    TP_CUDA_OK(cudaMemcpyAsync(d, xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, f->stream),
               { cudaFreeAsync(d, f->stream); });
    k_rasterize<<<(unsigned)((n + 255) / 256), 256, 0, f->stream>>>(d, n, f->grid, f->desc.chassis_height, f->occ3d,
                                                                  f->occ2d, f->occ2d_crit);
    cudaFreeAsync(d, f->stream);
    TP_CUDA_OK(cudaStreamSynchronize(f->stream), {});
    f->ready = false;
    f->grid.ready = 0;
    return TOPAY_OK;
}

extern "C" int topay_field_rebuild(topay_field* f) {
    if (!f) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(f->device);
    cudaStream_t q = f->stream;
    const unsigned tb = (unsigned)((f->n2 + 255) / 256);
    int rc;
    cudaEventRecord(f->ev0, q);
    // the four 2-D maps, in the reference's order (grid_map.cpp:138-423)
    if ((rc = signed_edt(f, f->occ2d, 1, f->nx, f->ny, f->esdf2d, f->sq_pos[0], f->sq_neg[0])) != TOPAY_OK) return rc;
    if ((rc = signed_edt(f, f->occ2d_crit, 1, f->nx, f->ny, f->esdf2d_crit, f->sq_pos[2], f->sq_neg[2])) != TOPAY_OK)
        return rc;
    k_threshold<<<tb, 256, 0, q>>>(f->esdf2d_crit, f->desc.chassis_colli_radius, f->src2d, f->n2);
    if ((rc = signed_edt(f, f->src2d, 1, f->nx, f->ny, f->esdf2d_crit, f->sq_pos[2], f->sq_neg[2])) != TOPAY_OK)
        return rc;
    k_threshold<<<tb, 256, 0, q>>>(f->esdf2d, f->desc.chassis_colli_radius, f->src2d, f->n2);
    if ((rc = signed_edt(f, f->src2d, 1, f->nx, f->ny, f->esdf2d_inflate, f->sq_pos[1], f->sq_neg[1])) != TOPAY_OK)
        return rc;
    cudaEventRecord(f->ev1, q);
    // the 3-D map (grid_map.cpp:425-518): z, y, x
    if ((rc = signed_edt(f, f->occ3d, f->nx, f->ny, f->nz, f->esdf3d, f->sq_pos[3], f->sq_neg[3])) != TOPAY_OK)
        return rc;
    cudaEventRecord(f->ev2, q);
    TP_CUDA_OK(cudaStreamSynchronize(q), {});
    TP_CUDA_OK(cudaGetLastError(), {});
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, f->ev0, f->ev1);
    cudaEventElapsedTime(&b, f->ev1, f->ev2);
    f->ms_total = a + b;
    f->ms_3d = b;
    f->ready = true;
    f->grid.ready = 1;
    return TOPAY_OK;
}

extern "C" int topay_field_last_rebuild_ms(topay_field* f, float* ms_total, float* ms_3d) {
    if (!f) return TOPAY_ERR_INVALID_ARG;
    if (ms_total) *ms_total = f->ms_total;
    if (ms_3d) *ms_3d = f->ms_3d;
    return TOPAY_OK;
}

static int query_common(topay_field* f, const double* pos, int64_t n, int dim, int which, double* dist,
                        double* grad, int value_only) {
    if (!f || (!pos && n > 0) || !dist || n < 0) return TOPAY_ERR_INVALID_ARG;
    if (!f->ready) {
        tp_set_error("field not built: call topay_field_rebuild first");
        return TOPAY_ERR_NOT_READY;
    }
    if (n == 0) return TOPAY_OK;
    cudaSetDevice(f->device);
    double *dpos = nullptr, *dd = nullptr, *dg = nullptr;
    TP_CUDA_OK(cudaMallocAsync(&dpos, (size_t)n * dim * 8, f->stream), {});
    TP_CUDA_OK(cudaMallocAsync(&dd, (size_t)n * 8, f->stream), { cudaFreeAsync(dpos, f->stream); });
    if (grad)
        TP_CUDA_OK(cudaMallocAsync(&dg, (size_t)n * dim * 8, f->stream),
                   { cudaFreeAsync(dpos, f->stream); cudaFreeAsync(dd, f->stream); });
    cudaMemcpyAsync(dpos, pos, (size_t)n * dim * 8, cudaMemcpyHostToDevice, f->stream);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dim == 3)
        k_query3d<<<blocks, 256, 0, f->stream>>>(f->grid, dpos, n, dd, dg, value_only);
    else {
        const double* buf = which == TOPAY_MAP2D_CRITICAL ? f->esdf2d_crit
                            : which == TOPAY_MAP2D_INFLATE ? f->esdf2d_inflate : f->esdf2d;
        k_query2d<<<blocks, 256, 0, f->stream>>>(f->grid, buf, dpos, n, dd, dg, value_only);
    }
    cudaMemcpyAsync(dist, dd, (size_t)n * 8, cudaMemcpyDeviceToHost, f->stream);
    if (grad) cudaMemcpyAsync(grad, dg, (size_t)n * dim * 8, cudaMemcpyDeviceToHost, f->stream);
    cudaFreeAsync(dpos, f->stream);
    cudaFreeAsync(dd, f->stream);
    if (dg) cudaFreeAsync(dg, f->stream);
    cudaError_t e = cudaStreamSynchronize(f->stream);
    TP_CUDA_OK(e, {});
    TP_CUDA_OK(cudaGetLastError(), {});
    return TOPAY_OK;
}

extern "C" int topay_field_query3d(topay_field* f, const double* pos, int64_t n, double* dist, double* grad) {
    return query_common(f, pos, n, 3, TOPAY_MAP3D, dist, grad, 0);
}
extern "C" int topay_field_query2d(topay_field* f, const double* pos, int64_t n, int which, double* dist,
                                   double* grad) {
    if (which < 0 || which > 2) return TOPAY_ERR_INVALID_ARG;
    return query_common(f, pos, n, 2, which, dist, grad, 0);
}
extern "C" int topay_field_distance3d(topay_field* f, const double* pos, int64_t n, double* dist) {
    return query_common(f, pos, n, 3, TOPAY_MAP3D, dist, nullptr, 1);
}
extern "C" int topay_field_distance2d(topay_field* f, const double* pos, int64_t n, double* dist) {
    return query_common(f, pos, n, 2, TOPAY_MAP2D_FLAT, dist, nullptr, 1);
}

extern "C" int topay_field_query3d_dev(topay_field* f, const double* pos_dev, int64_t n, double* dist_dev,
                                       double* grad_dev) {
    if (!f || !pos_dev || !dist_dev || n < 0) return TOPAY_ERR_INVALID_ARG;
    if (!f->ready) return TOPAY_ERR_NOT_READY;
    if (n == 0) return TOPAY_OK;
    cudaSetDevice(f->device);
    k_query3d<<<(unsigned)((n + 255) / 256), 256, 0, f->stream>>>(f->grid, pos_dev, n, dist_dev, grad_dev, 0);
    TP_CUDA_OK(cudaGetLastError(), {});
    return TOPAY_OK;
}
extern "C" int topay_field_sync(topay_field* f) {
    if (!f) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(f->device);
    TP_CUDA_OK(cudaStreamSynchronize(f->stream), {});
    return TOPAY_OK;
}

extern "C" int topay_field_whole_body_collision(topay_field* f, const topay_robot_params* robot,
                                                const double* states, int64_t n, int8_t* out) {
    if (!f || !robot || (!states && n > 0) || !out || n < 0) return TOPAY_ERR_INVALID_ARG;
    if (!f->ready) return TOPAY_ERR_NOT_READY;
    if (n == 0) return TOPAY_OK;
    cudaSetDevice(f->device);
    TpParams P;
    memset(&P, 0, sizeof(P));
    P.robot = *robot;
    topay_opt_params_default(&P.opt);
    tp_derive_params(P);
    double* ds = nullptr;
    int8_t* dout = nullptr;
    TP_CUDA_OK(cudaMallocAsync(&ds, (size_t)n * 10 * 8, f->stream), {});
    TP_CUDA_OK(cudaMallocAsync(&dout, (size_t)n, f->stream), { cudaFreeAsync(ds, f->stream); });
    cudaMemcpyAsync(ds, states, (size_t)n * 10 * 8, cudaMemcpyHostToDevice, f->stream);
    k_whole_body<<<(unsigned)((n + 127) / 128), 128, 0, f->stream>>>(f->grid, P, ds, n, dout);
    cudaMemcpyAsync(out, dout, (size_t)n, cudaMemcpyDeviceToHost, f->stream);
    cudaFreeAsync(ds, f->stream);
    cudaFreeAsync(dout, f->stream);
    cudaError_t e = cudaStreamSynchronize(f->stream);
    TP_CUDA_OK(e, {});
    return TOPAY_OK;
}

// kind as k_field_misc; in_a / in_b: n x wa doubles (host), in_i: n x 2 int32 (host)
static int field_misc(topay_field* f, int kind, const double* in_a, int wa, const double* in_b, const int32_t* in_i,
                      int64_t n, double threshold, int critical, int8_t* flag, double* val) {
    if (!f || n < 0 || (n > 0 && ((kind <= 3 && !in_a) || (kind == 2 && !in_b) || (kind == 4 && !in_i) ||
                                  (kind <= 2 && !flag) || (kind >= 3 && !val))))
        return TOPAY_ERR_INVALID_ARG;
    if (!f->ready) {
        tp_set_error("field queried before topay_field_rebuild");
        return TOPAY_ERR_NOT_READY;
    }
    if (n == 0) return TOPAY_OK;
    cudaSetDevice(f->device);
    const size_t ba = in_a ? (size_t)n * wa * 8 : 0, bb = in_b ? (size_t)n * wa * 8 : 0, bi = in_i ? (size_t)n * 8 : 0;
    const size_t bv = (size_t)n * 8, bf = (size_t)n;
    char* d = nullptr;
    TP_CUDA_OK(cudaMallocAsync(&d, ba + bb + bi + bv + bf + 64, f->stream), {});
    double* da = (double*)d;
    double* db = (double*)(d + ba);
    int32_t* di = (int32_t*)(d + ba + bb);
    double* dv = (double*)(d + ba + bb + bi);
    int8_t* df = (int8_t*)(d + ba + bb + bi + bv);
    if (in_a) cudaMemcpyAsync(da, in_a, ba, cudaMemcpyHostToDevice, f->stream);
    if (in_b) cudaMemcpyAsync(db, in_b, bb, cudaMemcpyHostToDevice, f->stream);
    if (in_i) cudaMemcpyAsync(di, in_i, bi, cudaMemcpyHostToDevice, f->stream);
    k_field_misc<<<(unsigned)((n + 255) / 256), 256, 0, f->stream>>>(f->grid, kind, da, db, di, n, threshold, critical,
                                                                    df, dv);
    if (kind <= 2) cudaMemcpyAsync(flag, df, bf, cudaMemcpyDeviceToHost, f->stream);
    else cudaMemcpyAsync(val, dv, bv, cudaMemcpyDeviceToHost, f->stream);
    cudaFreeAsync(d, f->stream);
    cudaError_t e = cudaStreamSynchronize(f->stream);
    TP_CUDA_OK(e, {});
    TP_CUDA_OK(cudaGetLastError(), {});
    return TOPAY_OK;
}

extern "C" int topay_field_is_collision2d(topay_field* f, const double* pos, int64_t n, double threshold, int8_t* out) {
    return field_misc(f, 0, pos, 2, nullptr, nullptr, n, threshold, 0, out, nullptr);
}
extern "C" int topay_field_is_collision3d(topay_field* f, const double* pos, int64_t n, double threshold, int8_t* out) {
    return field_misc(f, 1, pos, 3, nullptr, nullptr, n, threshold, 0, out, nullptr);
}
extern "C" int topay_field_is_line_collision_grid2d(topay_field* f, const double* p1, const double* p2, int64_t n,
                                                    double threshold, int8_t* out) {
    return field_misc(f, 2, p1, 2, p2, nullptr, n, threshold, 0, out, nullptr);
}
extern "C" int topay_field_dist_coarse2d(topay_field* f, const double* pos, int64_t n, int critical, double* out) {
    return field_misc(f, 3, pos, 2, nullptr, nullptr, n, 0.0, critical, nullptr, out);
}
extern "C" int topay_field_dist_coarse2i(topay_field* f, const int32_t* idx, int64_t n, int critical, double* out) {
    return field_misc(f, 4, nullptr, 2, nullptr, idx, n, 0.0, critical, nullptr, out);
}

extern "C" int topay_field_line_visible(topay_field* f, const double* p1, const double* p2, int64_t n, double thresh,
                                        int use_critical, int8_t* visible, double* pc) {
    if (!f || n < 0 || (n > 0 && (!p1 || !p2 || !visible || !pc))) return TOPAY_ERR_INVALID_ARG;
    if (!f->ready) {
        tp_set_error("field queried before topay_field_rebuild");
        return TOPAY_ERR_NOT_READY;
    }
    if (n == 0) return TOPAY_OK;
    cudaSetDevice(f->device);
    const size_t bp = (size_t)n * 3 * 8;
    char* d = nullptr;
    TP_CUDA_OK(cudaMallocAsync(&d, 3 * bp + (size_t)n + 64, f->stream), {});
    double *d1 = (double*)d, *d2 = (double*)(d + bp), *dc = (double*)(d + 2 * bp);
    int8_t* dv = (int8_t*)(d + 3 * bp);
    cudaMemcpyAsync(d1, p1, bp, cudaMemcpyHostToDevice, f->stream);
    cudaMemcpyAsync(d2, p2, bp, cudaMemcpyHostToDevice, f->stream);
    cudaMemcpyAsync(dc, pc, bp, cudaMemcpyHostToDevice, f->stream);      // pc of a visible segment stays as passed in
    k_line_visib<<<(unsigned)((n + 127) / 128), 128, 0, f->stream>>>(f->grid, d1, d2, n, thresh, use_critical, dv, dc);
    cudaMemcpyAsync(visible, dv, (size_t)n, cudaMemcpyDeviceToHost, f->stream);
    cudaMemcpyAsync(pc, dc, bp, cudaMemcpyDeviceToHost, f->stream);
    cudaFreeAsync(d, f->stream);
    cudaError_t e = cudaStreamSynchronize(f->stream);
    TP_CUDA_OK(e, {});
    TP_CUDA_OK(cudaGetLastError(), {});
    return TOPAY_OK;
}

extern "C" int topay_discretize_path(const double* path, int n, int pt_num, double* out);
extern "C" double topay_path_length(const double* path, int n);

// TopologyPRM::sameTopoPath (topo_prm.cpp:424-448) for many pairs of paths at once: the equally spaced points of
// both paths of every pair (host arithmetic) become one batch of visibility rays (one launch), a pair is the same
// class when all of its rays are free.
extern "C" int topay_field_same_topo_paths(topay_field* f, const double* pts, const int32_t* offsets, int n_paths,
                                           const int32_t* pairs, int n_pairs, double thresh, int use_critical,
                                           int8_t* same) {
    if (!f || !pts || !offsets || !pairs || !same || n_paths < 1 || n_pairs < 0) return TOPAY_ERR_INVALID_ARG;
    std::vector<double> p1, p2;
    std::vector<int64_t> first(n_pairs + 1, 0);
    std::vector<double> a, b;
    for (int k = 0; k < n_pairs; k++) {
        const int i = pairs[2 * k], j = pairs[2 * k + 1];
        if (i < 0 || j < 0 || i >= n_paths || j >= n_paths) return TOPAY_ERR_INVALID_ARG;
        const double *pa = pts + 3 * (size_t)offsets[i], *pb = pts + 3 * (size_t)offsets[j];
        const int na = offsets[i + 1] - offsets[i], nb = offsets[j + 1] - offsets[j];
        if (na < 2 || nb < 2) return TOPAY_ERR_INVALID_ARG;
        const double max_len = std::max(topay_path_length(pa, na), topay_path_length(pb, nb));
        const int pt_num = (int)ceil(max_len / f->desc.resolution);
        if (pt_num < 2) {       // both paths inside one cell: the reference divides by pt_num - 1 = 0 here
            tp_set_error("same_topo_paths: paths shorter than one cell");
            return TOPAY_ERR_INVALID_ARG;
        }
        a.resize((size_t)pt_num * 3);
        b.resize((size_t)pt_num * 3);
        topay_discretize_path(pa, na, pt_num, a.data());
        topay_discretize_path(pb, nb, pt_num, b.data());
        p1.insert(p1.end(), a.begin(), a.end());
        p2.insert(p2.end(), b.begin(), b.end());
        first[k + 1] = first[k] + pt_num;
    }
    const int64_t n = first[n_pairs];
    std::vector<int8_t> vis((size_t)std::max<int64_t>(n, 1));
    std::vector<double> pc((size_t)std::max<int64_t>(n, 1) * 3, 0.0);
    const int rc = topay_field_line_visible(f, p1.data(), p2.data(), n, thresh, use_critical, vis.data(), pc.data());
    if (rc != TOPAY_OK) return rc;
    for (int k = 0; k < n_pairs; k++) {
        int8_t all = 1;
        for (int64_t r = first[k]; r < first[k + 1]; r++) all &= vis[r];
        same[k] = all;
    }
    return TOPAY_OK;
}

extern "C" int topay_field_download(topay_field* f, int which, double* esdf_out) {
    if (!f || !esdf_out || which < 0 || which > 3) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(f->device);
    const double* src = which == TOPAY_MAP3D ? f->esdf3d
                        : which == TOPAY_MAP2D_CRITICAL ? f->esdf2d_crit
                        : which == TOPAY_MAP2D_INFLATE ? f->esdf2d_inflate : f->esdf2d;
    const size_t n = which == TOPAY_MAP3D ? f->n3 : f->n2;
    TP_CUDA_OK(cudaMemcpy(esdf_out, src, n * 8, cudaMemcpyDeviceToHost), {});
    return TOPAY_OK;
}

extern "C" int topay_field_download_sqdist(topay_field* f, int which, int32_t* pos_sq, int32_t* neg_sq) {
    if (!f || which < 0 || which > 3) return TOPAY_ERR_INVALID_ARG;
    if (!f->keep_sq) {
        tp_set_error("integer grids were not kept (topay_field_set_keep_sqdist(f, 0))");
        return TOPAY_ERR_NOT_READY;
    }
    cudaSetDevice(f->device);
    const size_t n = which == TOPAY_MAP3D ? f->n3 : f->n2;
    if (pos_sq) TP_CUDA_OK(cudaMemcpy(pos_sq, f->sq_pos[which], n * 4, cudaMemcpyDeviceToHost), {});
    if (neg_sq) TP_CUDA_OK(cudaMemcpy(neg_sq, f->sq_neg[which], n * 4, cudaMemcpyDeviceToHost), {});
    return TOPAY_OK;
}

extern "C" int topay_field_download_occupancy(topay_field* f, int which, int8_t* occ_out) {
    if (!f || !occ_out || which < 0 || which > 3 || which == TOPAY_MAP2D_INFLATE) return TOPAY_ERR_INVALID_ARG;
    cudaSetDevice(f->device);
    const int8_t* src = which == TOPAY_MAP3D ? f->occ3d : which == TOPAY_MAP2D_CRITICAL ? f->occ2d_crit : f->occ2d;
    const size_t n = which == TOPAY_MAP3D ? f->n3 : f->n2;
    TP_CUDA_OK(cudaMemcpy(occ_out, src, n, cudaMemcpyDeviceToHost), {});
    return TOPAY_OK;
}
