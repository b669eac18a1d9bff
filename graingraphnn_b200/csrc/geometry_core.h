// geometry_core.h — the per-grain arithmetic of gg_region_center, written once for the device kernel (geometry.cu) and for
// the host harness of the CPU test suite (tests/test_geometry_core.py compiles this header with g++: the arithmetic is
// IEEE add / divide / compare only, so host and device agree bit for bit).
//
// Follows graph_datastruct.py:686-708 (graph.update, periodic branch) with the numpy scalar types the reference ends up
// with: the first vertex of a region stays float32 (its "+1" shift rounds in float32), every later vertex is float64
// (`x += int64` in periodic_move :66-67), the first periodic_move compares a float32 difference, later ones a float64
// difference, and np.mean adds the first element to numpy's pairwise sum of the rest (8 running lanes from 8 elements on).
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define GG_HD inline
#include <cmath>
#else
#define GG_HD __host__ __device__ __forceinline__
#endif

// Global joint coordinate of test.py:472-474: (patch + offset) / domain_factor in fp32 (true division), or the patch
// coordinate itself when the domain is not scaled.
GG_HD float gg_global_coord(float patch, float offset, float factor) {
#ifdef __CUDA_ARCH__
    return factor > 1.0f ? __fdiv_rn(__fadd_rn(patch, offset), factor) : patch;
#else
    if (!(factor > 1.0f)) return patch;
    volatile float s = patch + offset;
    return s / factor;
#endif
}

// test.py:558-559: (fp32(centre) * factor) % 1 with torch's remainder (sign of the divisor).
GG_HD float gg_patch_coord(float c, float factor) {
    if (!(factor > 1.0f)) return c;
#ifdef __CUDA_ARCH__
    float v = __fmul_rn(c, factor);
#else
    volatile float vv = c * factor;
    float v = vv;
#endif
    float r = fmodf(v, 1.0f);
    if (r != 0.0f && r < 0.0f) r += 1.0f;
    return r;
}

// Vertex source: key[k] = rank of the k-th incident joint of this grain (position of the joint's first appearance as a
// target in the grain->joint edge list = its place in the reference's joint2vertex dict, graph_trajectory.py:1062-1080),
// fetch(k, &x, &y) = its global fp32 coordinates.  Vertices are visited in increasing key order by repeated minimum
// search (a grain has ~6 joints; no per-thread array, any degree works).
template <class Fetch>
struct GGRegionWalk {
    const int32_t* key;
    int32_t n;
    Fetch fetch;
    int32_t last;
    GG_HD void rewind() { last = -1; }
    GG_HD void next(float* x, float* y) {
        int32_t best = 0x7fffffff, at = 0;
        for (int32_t k = 0; k < n; ++k) {
            int32_t v = key[k];
            if (v > last && v < best) { best = v; at = k; }
        }
        last = best;
        fetch(at, x, y);
    }
};

// The same walk over a joint list that is ALREADY in dict order (gg_region_sort ran once for the topology): the per-step
// kernel then reads each joint id once per pass, no keys.
template <class Fetch>
struct GGOrderedWalk {
    int32_t n;
    Fetch fetch;
    int32_t at;
    GG_HD void rewind() { at = 0; }
    GG_HD void next(float* x, float* y) { fetch(at++, x, y); }
};

// One axis of the chain unwrap; state = previous moved vertex.
struct GGChain {
    float first;
    double prev;
    int32_t i;
    GG_HD void start(float v0) { first = v0; prev = (double)v0; i = 0; }
    GG_HD double step(float v) {                  // periodic_move(verts[i], verts[i-1]), graph_datastruct.py:55-72
        ++i;
        int shift;
        if (i == 1) {
#ifdef __CUDA_ARCH__
            float rel = __fsub_rn(v, first);
#else
            volatile float rr = v - first;
            float rel = rr;
#endif
            shift = (rel < -0.5f ? 1 : 0) - (rel > 0.5f ? 1 : 0);
        } else {
            double rel = (double)v - prev;
            shift = (rel < -0.5 ? 1 : 0) - (rel > 0.5 ? 1 : 0);
        }
        prev = (double)v + (double)shift;
        return prev;
    }
};

// numpy's add.reduce over float64 [a0, rest...]: a0 + pairwise_sum(rest) (loops_utils.h.src, blocks of <= 128).
struct GGNumpySum {
    double a0, res, r[8];
    int32_t m, i, full;                            // m = elements after the first, full = m - m % 8
    GG_HD void start(double first, int32_t n_total) {
        a0 = first; m = n_total - 1; i = 0; res = 0.0; full = m - (m & 7);
    }
    GG_HD void add(double v) {
        if (m < 8) res += v;
        else if (i < full) {
            // static indexing keeps r[] in registers; the first eight elements seed the lanes
            const bool seed = i < 8;
            switch (i & 7) {
                case 0: r[0] = seed ? v : r[0] + v; break; case 1: r[1] = seed ? v : r[1] + v; break;
                case 2: r[2] = seed ? v : r[2] + v; break; case 3: r[3] = seed ? v : r[3] + v; break;
                case 4: r[4] = seed ? v : r[4] + v; break; case 5: r[5] = seed ? v : r[5] + v; break;
                case 6: r[6] = seed ? v : r[6] + v; break; default: r[7] = seed ? v : r[7] + v; break;
            }
        } else {
            if (i == full) res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            res += v;
        }
        ++i;
    }
    GG_HD double total() {
        if (m >= 8 && full == m) res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        return a0 + res;
    }
};

// Centre of one region.  Returns false (centre untouched) for regions of <= 1 vertex (graph_datastruct.py:684).
template <class Walk>
GG_HD bool gg_region_center_one(Walk& w, double* cx, double* cy) {
    const int32_t n = w.n;
    if (n <= 1) return false;
    const float feps = -1e-12f;                    // `j > -eps` on a float32 scalar compares in float32 (NEP 50)
    const double deps = -1e-12;
    float x0, y0, x, y;
    // pass 1: is every moved vertex inside the lower bound, per axis (graph_datastruct.py:697-700)
    w.rewind();
    w.next(&x0, &y0);
    GGChain chx, chy;
    chx.start(x0); chy.start(y0);
    bool inx = x0 > feps, iny = y0 > feps;
    for (int32_t k = 1; k < n; ++k) {
        w.next(&x, &y);
        inx = (chx.step(x) > deps) && inx;
        iny = (chy.step(y) > deps) && iny;
    }
    // pass 2: shifted vertices (:701-703) summed the way np.mean sums them (:707)
#ifdef __CUDA_ARCH__
    const float fx0 = inx ? x0 : __fadd_rn(x0, 1.0f), fy0 = iny ? y0 : __fadd_rn(y0, 1.0f);
#else
    volatile float sx = x0 + 1.0f, sy = y0 + 1.0f;
    const float fx0 = inx ? x0 : (float)sx, fy0 = iny ? y0 : (float)sy;
#endif
    const double addx = inx ? 0.0 : 1.0, addy = iny ? 0.0 : 1.0;
    GGNumpySum sumx, sumy;
    sumx.start((double)fx0, n); sumy.start((double)fy0, n);
    w.rewind();
    w.next(&x0, &y0);
    chx.start(x0); chy.start(y0);
    for (int32_t k = 1; k < n; ++k) {
        w.next(&x, &y);
        sumx.add(chx.step(x) + addx);
        sumy.add(chy.step(y) + addy);
    }
    *cx = sumx.total() / (double)n;
    *cy = sumy.total() / (double)n;
    return true;
}
