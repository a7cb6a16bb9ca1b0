// Device-side building blocks of the RadSearch step / reset kernels (sm_100a).
// R: = /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py
//
// Everything boolean is exact int32 arithmetic on the integer lattice the reference lives on;
// valid for |coordinate| <= 16383 (RS_ST_COORD_RANGE is raised otherwise).
#pragma once
#ifdef RS_HOST_EMU
#include "cuda_host_shim.h"   // tests/emu: compiles this header as host C++ to debug the kernel logic without a GPU
#else
#include <cuda_runtime.h>
#endif
#ifdef RS_HOST_EMU
#define RS_FAST_LOGF(x) logf(x)
#else
#define RS_FAST_LOGF(x) __logf(x)
#endif
#include <stdint.h>

#include "../../include/radsearch_b200.h"

namespace rs {

// get_step R:205-224: 0 left, 1 up-left, 2 up, 3 up-right, 4 right, 5 down-right, 6 down, 7 down-left, 8 idle
__device__ __forceinline__ int step_dx(int a) {
    // packed LUT, 8 bits each biased by 128: {-100,-71,0,71,100,71,0,-71,0}
    const int cx = (a == 0 || a == 1 || a == 7) ? -1 : ((a >= 3 && a <= 5) ? 1 : 0);
    const int mag = (a & 1) ? 71 : 100;
    return a == 8 ? 0 : cx * mag;
}
__device__ __forceinline__ int step_dy(int a) {
    const int cy = (a >= 1 && a <= 3) ? 1 : ((a >= 5 && a <= 7) ? -1 : 0);
    const int mag = (a & 1) ? 71 : 100;
    return a == 8 ? 0 : cy * mag;
}
__device__ __forceinline__ int coef_x(int d) { return (d == 0 || d == 1 || d == 7) ? -1 : ((d >= 3 && d <= 5) ? 1 : 0); }
__device__ __forceinline__ int coef_y(int d) { return (d >= 1 && d <= 3) ? 1 : ((d >= 5 && d <= 7) ? -1 : 0); }

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based stream: key = seed, counter = (env_id, domain<<24 | agent<<16 | block, step_ctr)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    const double *u;      // injected next_double stream (or nullptr)
    int n_u, pos;
    uint32_t k0, k1, c0, c1, c2, c3;
    uint32_t buf[4];
    int have;
    uint32_t status;

    __device__ __forceinline__ void init_philox(uint64_t seed, uint32_t env_id, uint32_t domain, uint32_t agent,
                                                uint64_t step_ctr) {
        u = nullptr; n_u = 0; pos = 0;
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        c0 = env_id; c1 = (domain << 24) | (agent << 16);
        c2 = (uint32_t)step_ctr; c3 = (uint32_t)(step_ctr >> 32);
        have = 0; status = 0;
    }
    __device__ __forceinline__ void init_inject(const double *p, int n) {
        u = p; n_u = n; pos = 0; have = 0; status = 0;
        k0 = k1 = c0 = c1 = c2 = c3 = 0;
    }
    __device__ __forceinline__ uint32_t u32() {
        if (have == 0) {
            philox4x32_10(c0, c1, c2, c3, k0, k1, buf);
            c1 += 1;
            have = 4;
        }
        const int i = 4 - have;
        have -= 1;
        return i == 0 ? buf[0] : (i == 1 ? buf[1] : (i == 2 ? buf[2] : buf[3]));
    }
    // numpy next_double: (next_uint64 >> 11) * 2^-53
    __device__ __forceinline__ double next_double() {
        if (u) {
            if (pos >= n_u) { status |= RS_ST_UNIFORMS_OUT; return 0.5; }
            return u[pos++];
        }
        const uint32_t lo = u32();
        const uint32_t hi = u32();
        const uint64_t w = ((uint64_t)hi << 32) | lo;
        return (double)(w >> 11) * (1.0 / 9007199254740992.0);
    }
    // unbiased integer in [0, range) (Lemire)
    __device__ __forceinline__ uint32_t below(uint32_t range) {
        uint64_t m = (uint64_t)u32() * range;
        uint32_t l = (uint32_t)m;
        if (l < range) {
            const uint32_t t = (0u - range) % range;
            while (l < t) {
                m = (uint64_t)u32() * range;
                l = (uint32_t)m;
            }
        }
        return (uint32_t)(m >> 32);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// numpy Generator.poisson (distributions.c: random_poisson / _ptrs / _mult / random_loggam), call site R:498
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ double loggam(double x) {
    const double a[10] = {8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04,
                          -5.952380952380952e-04, 8.417508417508418e-04, -1.917526917526918e-03,
                          6.410256410256410e-03, -2.955065359477124e-02, 1.796443723688307e-01,
                          -1.39243221690590e+00};
    if (x == 1.0 || x == 2.0) return 0.0;
    const long long n = (x < 7.0) ? (long long)(7 - x) : 0;
    double x0 = x + (double)n;
    const double x2 = (1.0 / x0) * (1.0 / x0);
    double gl0 = a[9];
#pragma unroll
    for (int k = 8; k >= 0; k--) {
        gl0 *= x2;
        gl0 += a[k];
    }
    double gl = gl0 / x0 + 0.5 * 1.8378770664093453e+00 + (x0 - 0.5) * log(x0) - x0;
    if (x < 7.0) {
        for (long long k = 1; k <= n; k++) {
            gl -= log(x0 - 1.0);
            x0 -= 1.0;
        }
    }
    return gl;
}

// fp32 Stirling log-gamma(k+1) for the KS-equivalent fast acceptance test (k >= 0)
__device__ __forceinline__ float lgamma1p_fast(float k) {
    if (k < 8.0f) return lgammaf(k + 1.0f);
    const float x = k + 1.0f;
    const float inv = 1.0f / x;
    return (x - 0.5f) * RS_FAST_LOGF(x) - x + 0.918938533f + inv * (0.0833333333f - inv * inv * 0.00277777778f);
}

template <bool kFast>
__device__ __forceinline__ long long poisson(Rng &g, double lam) {
    if (lam >= 10) {
        const double slam = sqrt(lam);
        const double b = 0.931 + 2.53 * slam;
        const double a = -0.059 + 0.02483 * b;
        const double vr = 0.9277 - 3.6224 / (b - 2);
        for (int it = 0; it < 1000; it++) {
            const double U = g.next_double() - 0.5;
            const double V = g.next_double();
            const double us = 0.5 - fabs(U);
            const long long k = (long long)floor((2 * a / us + b) * U + lam + 0.43);
            if (us >= 0.07 && V <= vr) return k;
            if (k < 0 || (us < 0.013 && V > us)) continue;
            if (kFast) {
                const float invalpha = 1.1239f + 1.1328f / ((float)b - 3.4f);
                const float fus = (float)us;
                const float lhs = RS_FAST_LOGF((float)V) + RS_FAST_LOGF(invalpha) - RS_FAST_LOGF((float)a / (fus * fus) + (float)b);
                const float fk = (float)k;
                // -lam + k*log(lam) - lgamma(k+1) evaluated around k ~ lam without cancellation (Stirling):
                // k*log1p((lam-k)/k) - (lam-k) - 0.5*log(2*pi*k) - 1/(12k);  exact lgamma for small k
                float rhs;
                if (fk >= 8.0f) {
                    const float fd = (float)(lam - (double)k);
                    rhs = fk * log1pf(fd / fk) - fd - 0.5f * RS_FAST_LOGF(6.28318531f * fk) - 0.0833333333f / fk;
                } else {
                    const float flam = (float)lam;
                    rhs = -flam + fk * RS_FAST_LOGF(flam) - lgamma1p_fast(fk);
                }
                if (lhs <= rhs) return k;
            } else {
                const double loglam = log(lam);
                const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
                if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + (double)k * loglam - loggam((double)(k + 1))))
                    return k;
            }
            if (g.u && g.pos >= g.n_u) break;
        }
        g.status |= RS_ST_UNIFORMS_OUT;
        return (long long)floor(lam);
    } else if (lam == 0) {
        return 0;
    } else {
        const double enlam = exp(-lam);
        double prod = 1.0;
        long long X = 0;
        for (;;) {
            prod *= g.next_double();
            if (prod > enlam) X += 1; else return X;
            if (g.u && g.pos >= g.n_u) { g.status |= RS_ST_UNIFORMS_OUT; return X; }
        }
    }
}

// Single-precision PTRS for RS_F_FAST_POISSON (Philox stream only; KS-equivalent to numpy, not bit-identical): the same
// transformed-rejection algorithm with the hat/squeeze constants, the proposal offset (2a/us + b) * U and the acceptance
// test in fp32 (the offset is below ~1.3e3, so its rounding error is ~1e-4 of a count), 24-bit uniforms placed at cell
// centres (never 0 or 1), and lambda added in fp64.  One Philox4x32 block feeds two proposals.
struct PtrsF32 {
    float b, a, vr;
    double lam;
    __device__ __forceinline__ void init(double lam_) {
        lam = lam_;
        const float slam = sqrtf((float)lam_);
        b = 0.931f + 2.53f * slam;
        a = -0.059f + 0.02483f * b;
        vr = 0.9277f - 3.6224f / (b - 2.0f);
    }
    // 0 = rejected, 1 = accepted by the squeeze, 2 = accepted by the full test (only tried when `full`)
    __device__ __forceinline__ int propose(uint32_t xu, uint32_t xv, bool full, long long &k_out) const {
        const float U = ((float)(xu >> 8) + 0.5f) * 5.9604644775390625e-08f - 0.5f;
        const float V = ((float)(xv >> 8) + 0.5f) * 5.9604644775390625e-08f;
        const float us = 0.5f - fabsf(U);
        const float off = (2.0f * a / us + b) * U;
        const long long k = (long long)floor((double)off + lam + 0.43);
        k_out = k;
        if (us >= 0.07f && V <= vr) return 1;
        if (!full || k < 0 || (us < 0.013f && V > us)) return 0;
        const float invalpha = 1.1239f + 1.1328f / (b - 3.4f);
        const float lhs = RS_FAST_LOGF(V) + RS_FAST_LOGF(invalpha) - RS_FAST_LOGF(a / (us * us) + b);
        const float fk = (float)k;
        // -lam + k*log(lam) - lgamma(k+1) evaluated around k ~ lam without cancellation (Stirling):
        // k*log1p((lam-k)/k) - (lam-k) - 0.5*log(2*pi*k) - 1/(12k);  exact lgamma for small k
        float rhs;
        if (fk >= 8.0f) {
            const float fd = (float)(lam - (double)k);
            rhs = fk * log1pf(fd / fk) - fd - 0.5f * RS_FAST_LOGF(6.28318531f * fk) - 0.0833333333f / fk;
        } else {
            const float flam = (float)lam;
            rhs = -flam + fk * RS_FAST_LOGF(flam) - lgamma1p_fast(fk);
        }
        return lhs <= rhs ? 2 : 0;
    }
};

// lam >= 10.  Counter block j of the stream (seed; env_id, domain|agent|j, step_ctr) serves proposals 2j and 2j+1.
__device__ __noinline__ long long poisson_f32(uint64_t seed, uint32_t env_id, uint32_t domain, uint32_t agent,
                                              uint64_t ctr, double lam) {
    PtrsF32 s;
    s.init(lam);
    const uint32_t c1 = (domain << 24) | (agent << 16);
    for (uint32_t j = 0; j < 500u; j++) {
        uint32_t x[4];
        philox4x32_10(env_id, c1 + j, (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), x);
        long long k;
        if (s.propose(x[0], x[1], true, k)) return k;
        if (s.propose(x[2], x[3], true, k)) return k;
    }
    return (long long)floor(lam);
}

// Python round(x, 2) R:613: decimal rounding of the exact binary value, ties to even.
__device__ __forceinline__ double round2(double x) {
    const double p = __dmul_rn(x, 100.0);
    const double e = __fma_rn(x, 100.0, -p);   // exact residual: x*100 = p + e
    double n = rint(p);
    const double diff = p - n;
    if (diff == 0.5 || diff == -0.5) {
        if (e > 0) n = floor(p) + 1.0;
        else if (e < 0) n = floor(p);
    }
    return n / 100.0;
}

// ---------------------------------------------------------------------------------------------------------------
// exact lattice geometry
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_rect_closed(int px, int py, int4 r) {
    return r.x <= px && px <= r.z && r.y <= py && py <= r.w;
}
__device__ __forceinline__ bool in_rect_open(int px, int py, int4 r) {
    return r.x < px && px < r.z && r.y < py && py < r.w;
}

// Segment pq against the axis-aligned rectangle r by separating axes.  The four corner cross products
//   cr[i] = (corner_i - p) x (q - p),  corners in the order p0 (x0,y0), p1 (x0,y1), p2 (x1,y1), p3 (x1,y0)
// share four multiplications.  The open segment meets the OPEN rectangle (bit0: blocks visibility for the shortest
// path) iff the open coordinate intervals overlap on both axes and the line strictly separates two corners: on the line
// the x-slab, the y-slab and the segment are three open parameter intervals, pairwise intersecting by those three tests,
// hence with a common point (Helly in one dimension).  The closed forms of the same tests give bit1: the closed
// segment meets the CLOSED rectangle.  (p == q: bit0 is 0; never asked for a point strictly inside an obstruction.)
// |coordinates| <= 16383 keeps every product below 2^30 and every difference inside int32.
__device__ __forceinline__ int seg_rect(int px, int py, int qx, int qy, int4 r, int cr[4]) {
    const int dx = qx - px, dy = qy - py;
    const int a = (r.x - px) * dy, b = (r.z - px) * dy, c = (r.y - py) * dx, d = (r.w - py) * dx;
    cr[0] = a - c; cr[1] = a - d; cr[2] = b - d; cr[3] = b - c;
    const int mn = min(min(cr[0], cr[1]), min(cr[2], cr[3])), mx = max(max(cr[0], cr[1]), max(cr[2], cr[3]));
    const int xlo = min(px, qx), xhi = max(px, qx), ylo = min(py, qy), yhi = max(py, qy);
    const bool open_box = r.x < xhi && xlo < r.z && r.y < yhi && ylo < r.w;
    const bool closed_box = r.x <= xhi && xlo <= r.z && r.y <= yhi && ylo <= r.w;
    return (int)(open_box && mn < 0 && mx > 0) | ((int)(closed_box && mn <= 0 && mx >= 0) << 1);
}
__device__ __forceinline__ int seg_rect(int px, int py, int qx, int qy, int4 r) {
    int cr[4];
    return seg_rect(px, py, qx, qy, r, cr);
}

// A segment p -> q prepared for tests against many rectangles: everything that does not depend on the rectangle.
// The cross product of corner (x, y) is (x - px) * dy - (y - py) * dx = x * dy - y * dx + t with t = py * dx - px * dy:
// two multiply-adds per corner coordinate, no subtraction per rectangle (all values fit int32 for |coord| <= 16383).
struct Seg1 {
    int px, py, dx, dy, ndx, t, xlo, xhi, ylo, yhi;
};
__device__ __forceinline__ Seg1 make_seg1(int px, int py, int qx, int qy) {
    Seg1 s;
    s.px = px; s.py = py; s.dx = qx - px; s.dy = qy - py; s.ndx = -s.dx;
    s.t = py * s.dx - px * s.dy;
    s.xlo = min(px, qx); s.xhi = max(px, qx); s.ylo = min(py, qy); s.yhi = max(py, qy);
    return s;
}
// seg_rect's bit 0 (the segment meets the OPEN rectangle) as a predicate: min / max of the four corner cross products
// from the products of the two x and the two y coordinates, compared without forming the corner values
__device__ __forceinline__ bool seg_open1(const Seg1 &s, int4 r) {
    const int a = r.x * s.dy + s.t, b = r.z * s.dy + s.t;               // x part (+ t)
    const int c = r.y * s.dx, d = r.w * s.dx;                          // y part
    const bool mn_neg = min(a, b) < max(c, d), mx_pos = max(a, b) > min(c, d);
    return (r.x < s.xhi) & (s.xlo < r.z) & (r.y < s.yhi) & (s.ylo < r.w) & mn_neg & mx_pos;
}
// the same without the box comparison, for callers that have already filtered the rectangles by the segment's box
__device__ __forceinline__ bool seg_cross_open1(const Seg1 &s, int4 r) {
    const int a = r.x * s.dy + s.t, b = r.z * s.dy + s.t;
    const int c = r.y * s.dx, d = r.w * s.dx;
    return (min(a, b) < max(c, d)) & (max(a, b) > min(c, d));
}
// seg_rect with both bits as predicates and the four cross products (order p0, p1, p2, p3 as seg_rect)
__device__ __forceinline__ void seg_both1(const Seg1 &s, int4 r, bool &open, bool &closed, int cr[4]) {
    const int a = r.x * s.dy + s.t, b = r.z * s.dy + s.t;
    cr[0] = r.y * s.ndx + a; cr[1] = r.w * s.ndx + a; cr[2] = r.w * s.ndx + b; cr[3] = r.y * s.ndx + b;
    const int mn = min(min(cr[0], cr[1]), min(cr[2], cr[3])), mx = max(max(cr[0], cr[1]), max(cr[2], cr[3]));
    open = (r.x < s.xhi) & (s.xlo < r.z) & (r.y < s.yhi) & (s.ylo < r.w) & (mn < 0) & (mx > 0);
    closed = (r.x <= s.xhi) & (s.xlo <= r.z) & (r.y <= s.yhi) & (s.ylo <= r.w) & (mn <= 0) & (mx >= 0);
}

// a rectangle corner whose projection lies inside pq within 0.001 of it: cross^2 * 1e6 < |pq|^2 (so |cross| <= 3 on
// this lattice, and |pq| > 1000); cr[] from seg_rect
__device__ __forceinline__ bool corner_grazes(int px, int py, int dx, int dy, int l2, int4 r, const int cr[4]) {
    bool hit = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int ac = cr[i] < 0 ? -cr[i] : cr[i];
        if (ac <= 3 && ac * ac * 1000000 < l2) {
            const int wx = ((i < 2) ? r.x : r.z) - px, wy = ((i == 0 || i == 3) ? r.y : r.w) - py;
            const int t = wx * dx + wy * dy;
            hit = hit || (t >= 0 && t <= l2);
        }
    }
    return hit;
}

// vis.boundary_distance(Line_Segment(p,q), rect) < 0.001  (R:1110, 1141): touches/crosses the boundary, or a corner
// whose projection lies inside the segment is within 0.001 of it (cross^2 * 1e6 < |pq|^2, |cross| <= 3).
__device__ __forceinline__ bool los_blocked_rect(int px, int py, int qx, int qy, int4 r) {
    int cr[4];
    const int h = seg_rect(px, py, qx, qy, r, cr);
    if ((h & 2) && !(in_rect_open(px, py, r) && in_rect_open(qx, qy, r))) return true;
    const int dx = qx - px, dy = qy - py;
    const int l2 = dx * dx + dy * dy;
    if (l2 <= 1000000) return false;          // a corner with cross != 0 needs |pq| > 1000
    return corner_grazes(px, py, dx, dy, l2, r, cr);
}

__device__ __forceinline__ bool rects_touch(int4 a, int4 b) {
    const bool closed = a.x <= b.z && b.x <= a.z && a.y <= b.w && b.y <= a.w;
    const bool a_in_b = b.x < a.x && a.z < b.z && b.y < a.y && a.w < b.w;
    const bool b_in_a = a.x < b.x && b.z < a.z && a.y < b.y && b.w < a.w;
    return closed && !a_in_b && !b_in_a;
}
__device__ __forceinline__ bool rects_nested(int4 a, int4 b) {
    const bool a_in_b = b.x <= a.x && a.z <= b.z && b.y <= a.y && a.w <= b.w;
    const bool b_in_a = a.x <= b.x && b.z <= a.z && a.y <= b.y && b.w <= a.w;
    return a_in_b || b_in_a;
}

__device__ __forceinline__ int corner_x(int4 r, int i) { return (i < 2) ? r.x : r.z; }        // p0,p1 | p2,p3
__device__ __forceinline__ int corner_y(int4 r, int i) { return (i == 0 || i == 3) ? r.y : r.w; }

}  // namespace rs
