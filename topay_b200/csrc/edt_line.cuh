// Per-line arithmetic of the separable exact squared-distance transform (fillESDF,
// src/map/src/grid_map.cpp:89-123; ROG-Map's ring version src/rog_map/src/rog_map/esdf_map.cpp:842-900).
//
// The reference computes, for every cell q of a line,  D(q) = min_v (q - v)^2 + f(v)  by sweeping a
// lower envelope of parabolas in doubles whose values are exact integers. D is an integer, so any exact
// evaluation of the minimum is bit-identical; this file evaluates it in int32 with a divide and conquer
// over the queries that parallelises inside a line:
//
//   for q1 < q2, every minimiser of q1 is <= every minimiser of q2          (*)
//
//   (adding  f(a) + (q1-a)^2 <= f(b) + (q1-b)^2  and  f(b) + (q2-b)^2 <= f(a) + (q2-a)^2  for b < a gives
//   q1 >= q2). So once the minimisers of q - h and q + h are known the search for q is confined to the
//   cells between them, and the ranges of all queries of a level telescope: a level costs O(n) candidate
//   evaluations instead of O(n x distance), log2(n) levels per line. Inside its range a query still walks
//   outwards from q and stops as soon as d^2 alone reaches the best value (the window cut-off).
//
// "No source" is the finite value TP_INF32 on both sides of (*), so lines with few or no sources need no
// special case; results >= TP_INF32 are reported as "none". (A line that has a source at all never selects a
// "no source" cell, so runs of them can be jumped over.)
//
// Intermediates are SIGN-PACKED: a cell is either a source of the positive transform (occupied: distance
// to the nearest obstacle is 0) or of the negative one (free: distance to the nearest free cell is 0), never
// both, so one signed integer carries both transforms:  v > 0: positive transform = v, negative = 0;
// v < 0: positive = 0, negative = -v.  (int16 1-D distances after the first pass, int32 squared distances after
// the second: 2 + 4 B per voxel instead of 4 + 8.)
#pragma once
#include "hd.cuh"

#define TP_INF16 16383
#define TP_INF32 (1 << 29)

TP_HD int tp_fpos(int v) { return v > 0 ? v : 0; }   // callers clamp values >= TP_INF32 (skip-carrying cells)
TP_HD int tp_fneg(int v) { return v < 0 ? -v : 0; }
// sign-packed 1-D distance (first pass) -> sign-packed squared distance
TP_HD int tp_sq16(int v) {
    const int a = v < 0 ? -v : v;
    const int s = a >= TP_INF16 ? TP_INF32 : a * a;
    return v < 0 ? -s : s;
}

// "No source" cells of the positive transform (free cells whose value is TP_INF32) come in long runs in the second
// pass (columns without any obstacle), so each of them carries two skip distances in its low bits: to the nearest
// cell that is not "no source" on its left (bits 14-27; a distance that leaves the line if there is none) and on
// its right (bits 0-13). A search that steps onto such a cell jumps over the whole run. Any value >= TP_INF32
// still reads as "no source".
#define TP_SKIP_BITS 14
#define TP_SKIP_MASK ((1 << TP_SKIP_BITS) - 1)
TP_HD int tp_skip_make(int left, int right) { return TP_INF32 | (left << TP_SKIP_BITS) | right; }

// Positive transform of query q, the minimiser known to lie in [a_lo, a_hi] (a_lo <= a_hi): returns the
// minimum and one minimiser. f is the sign-packed line, element v at f[v * STRIDE] (STRIDE = tile width, a
// compile-time constant so that the walks are pointer bumps). A result >= TP_INF32 means "no source on this line".
template <int STRIDE>
TP_HD int tp_dc_query(const int* __restrict__ f, int q, int a_lo, int a_hi, int& arg) {
    int d = q - a_lo;
    int fv = f[a_lo * STRIDE];
    int best = d * d + (fv >= TP_INF32 ? TP_INF32 : tp_fpos(fv));
    int ba = a_lo;
    if (a_hi != a_lo) {
        d = a_hi - q;
        fv = f[a_hi * STRIDE];
        const int c = d * d + (fv >= TP_INF32 ? TP_INF32 : tp_fpos(fv));
        if (c < best) {
            best = c;
            ba = a_hi;
        }
    }
    // interior cells, nearest first on each side, walked by their distance d from q; d^2 >= best ends a side
    // (window cut-off)
    {
        const int v0 = q < a_hi ? q : a_hi - 1;            // left side: cells v0, v0 - 1, ..., a_lo + 1
        const int dmax = q - a_lo;                         // d < dmax  <=>  v > a_lo
        const int* p = f + v0 * STRIDE;
        for (d = q - v0; d < dmax;) {
            const int dd = d * d;
            if (dd >= best) break;
            fv = *p;
            if (fv >= TP_INF32) {
                const int sk = (fv >> TP_SKIP_BITS) & TP_SKIP_MASK;
                d += sk;
                p -= sk * STRIDE;
                continue;
            }
            const int c = dd + tp_fpos(fv);
            if (c < best) {
                best = c;
                ba = q - d;
            }
            d++;
            p -= STRIDE;
        }
    }
    {
        const int v0 = q + 1 > a_lo ? q + 1 : a_lo + 1;    // right side: cells v0, v0 + 1, ..., a_hi - 1
        const int dmax = a_hi - q;
        const int* p = f + v0 * STRIDE;
        for (d = v0 - q; d < dmax;) {
            const int dd = d * d;
            if (dd >= best) break;
            fv = *p;
            if (fv >= TP_INF32) {
                const int sk = fv & TP_SKIP_MASK;
                d += sk;
                p += sk * STRIDE;
                continue;
            }
            const int c = dd + tp_fpos(fv);
            if (c < best) {
                best = c;
                ba = q + d;
            }
            d++;
            p += STRIDE;
        }
    }
    arg = ba;
    return best;
}

// Negative transform of cell l by the outward search: its sources (free cells) are dense, so the search ends
// after a step or two almost everywhere (at once on a free cell, within an obstacle's thickness inside one).
template <int STRIDE>
TP_HD int tp_neg_search(const int* __restrict__ f, int n, int l) {
    int bn = tp_fneg(f[l * STRIDE]);
    for (int d = 1; d < n; d++) {
        const int dd = d * d;
        if (dd >= bn) break;
        if (l - d >= 0) {
            const int c = dd + tp_fneg(f[(l - d) * STRIDE]);
            bn = c < bn ? c : bn;
        }
        if (l + d < n) {
            const int c = dd + tp_fneg(f[(l + d) * STRIDE]);
            bn = c < bn ? c : bn;
        }
    }
    return bn;
}

// Level schedule. The line is cut into chunks of CH queries (a power of two; 32 on large grids, less on small
// ones so that the chunks of a grid still fill the machine). The boundary queries q_b = min(b CH, n - 1),
// b = 0 .. nb - 1, are answered first, each against the whole line (window cut-off + skips keep that short); chunk c
// then owns the queries strictly between q_c and q_{c+1} and answers them by halving: level h = top/2, top/4, ..., 1
// takes q = q_c + (2j + 1) h < q_{c+1} between the minimisers of q - h and min(q + h, q_{c+1}). A chunk is walked by
// ONE thread, so no barrier separates its levels.
TP_HD int tp_edt_boundaries(int n, int ch) { return n <= 1 ? 1 : (n - 1 + ch - 1) / ch + 1; }
TP_HD int tp_edt_boundary(int b, int n, int ch) { return b * ch < n - 1 ? b * ch : n - 1; }
// smallest power of two >= len
TP_HD int tp_dc_top(int len) {
    int p = 1;
    while (p < len) p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------------------------------------
// Lower envelope by two scans, one thread per line (large grids: the lines of a pass outnumber the lanes of the
// machine, so a line needs no parallelism inside it and the O(n) stack algorithm of the reference itself wins).
// Integer form (Meijster et al.): the stack holds the parabolas of the envelope, entry k = (s_k, t_k, g_k):
// source cell, first cell where it beats entry k - 1, its value. A new source u pops every entry it beats at
// that entry's own start  ((t - s)^2 + g > (t - u)^2 + g_u),  then starts at
//     w = 1 + floor((u^2 - s^2 + g_u - g) / (2 (u - s)))       (numerator >= 2 t (u - s) >= 0 after the pops)
// if w is still inside the line. The backward scan reads the envelope off: cell u belongs to the top entry,
// which is popped when u reaches its start. Everything is int32 (n < 2^14, values < 2^29).
//
// A cell is a source (value 0) of exactly one of the two transforms, and its own result for that transform is 0
// without any search; so a scan only emits the cells that are NOT its sources, and a source strictly inside a
// run of sources is not pushed at all (the run's end cells are at least as near to every cell outside the run).
// With free space everywhere the negative transform's stack therefore stays a handful of entries deep.
//
// The top entry lives in registers (sv, tv, gv); stk[k] is written on every push.
struct TpEnv {
    int q, sv, tv, gv;      // q = index of the top entry, -1 = empty
};
struct alignas(8) TpEnvEntry {
    unsigned st;            // s | t << 16
    int g;
};
template <bool NEG>
TP_HD int tp_env_val(int f) {                   // sign-packed cell -> its value in this transform (0 = source)
    const int v = NEG ? -f : f;
    return v <= 0 ? 0 : (v >= TP_INF32 ? TP_INF32 : v);
}
TP_HD void tp_env_load(TpEnv& e, const TpEnvEntry* stk) {
    const TpEnvEntry w = stk[e.q];
    e.sv = (int)(w.st & 0xffffu);
    e.tv = (int)(w.st >> 16);
    e.gv = w.g;
}
// one cell of the forward scan: value cur at cell u, prev / nxt = values of its neighbours (anything non-zero at the
// line ends)
TP_HD void tp_env_push(TpEnv& e, TpEnvEntry* stk, int n, int u, int cur, int prev, int nxt) {
    if (cur >= TP_INF32 || (cur | prev | nxt) == 0) return;
    while (e.q >= 0) {
        const int a = e.tv - e.sv, b = e.tv - u;
        if (a * a + e.gv <= b * b + cur) break;
        if (--e.q >= 0) tp_env_load(e, stk);
    }
    if (e.q < 0) {
        e.q = 0;
        e.sv = u;
        e.tv = 0;
        e.gv = cur;
        stk[0] = TpEnvEntry{(unsigned)u, cur};
        return;
    }
    const unsigned num = (unsigned)(u * u - e.sv * e.sv + cur - e.gv), den = 2u * (unsigned)(u - e.sv);
#if defined(__CUDA_ARCH__)
    // floor(num / den) is only needed while it is below n < 2^14: a float estimate (relative error < 2^-21, i.e.
    // less than one unit there) corrected by the exact integer remainder; anything estimated at 2^15 or more starts
    // past the line's end. (num < 2^30, den < 2^15.)
    const float qf = __fdividef(__uint2float_rn(num), __uint2float_rn(den));     // MUFU.RCP + multiply, 2 ulp
    if (qf >= 32768.0f) return;
    int qi = (int)qf;
    const int rem = (int)num - qi * (int)den;
    qi += rem < 0 ? -1 : (rem >= (int)den ? 1 : 0);
    const int w = 1 + qi;
#else
    const int w = 1 + (int)(num / den);
#endif
    if (w < n) {
        e.q++;
        e.sv = u;
        e.tv = w;
        e.gv = cur;
        stk[e.q] = TpEnvEntry{(unsigned)u | ((unsigned)w << 16), cur};
    }
}
// one cell of the backward scan (u = n - 1 .. 0): the envelope's value at u (TP_INF32 = no source on the line)
TP_HD int tp_env_pop(TpEnv& e, const TpEnvEntry* stk, int u) {
    if (e.q < 0) return TP_INF32;
    const int d = u - e.sv;
    int val = d * d + e.gv;
    val = val > TP_INF32 ? TP_INF32 : val;
    if (u == e.tv && --e.q >= 0) tp_env_load(e, stk);
    return val;
}
