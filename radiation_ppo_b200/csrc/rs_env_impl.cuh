// Per-environment logic of the RadSearch step / reset kernels.  One thread owns one environment; its obstruction
// rectangles, source-distance table and corner-visibility masks live in a private shared-memory column (Col<T>).
// R: = /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py, T: = algos/multiagent/train.py
#pragma once
#include "rs_device.cuh"

namespace rs {

template <typename T>
struct Col {
    T *p;
    int stride;
    __device__ __forceinline__ T &operator[](int i) const { return p[i * stride]; }
};

struct Params {
    int bx0, by0, bx1, by1;       // bbox
    int sx0, sy0, sx1, sy1;       // search area R:393-420
    int lo, hi;                   // observation_area
    int enforce, n_agents, obstruction_count, count_law, max_ep_len, k_max, standardize;
    double max_dist;              // R:423-425
    double inv_max_dist;          // 1 / max_dist rounded to nearest (div_const in rs_step1.cuh)
    double inv_scale;             // 1 / search_area[2][1]  R:435
};

__host__ __device__ inline Params make_params(const RsConfig &c) {
    Params p;
    p.bx0 = c.bbox[0]; p.by0 = c.bbox[1]; p.bx1 = c.bbox[2]; p.by1 = c.bbox[3];
    p.lo = c.obs_area[0]; p.hi = c.obs_area[1];
    p.sx0 = p.bx0 + p.lo; p.sy0 = p.by0 + p.lo; p.sx1 = p.bx1 - p.hi; p.sy1 = p.by1 - p.hi;
    p.enforce = c.enforce; p.n_agents = c.n_agents; p.obstruction_count = c.obstruction_count;
    p.count_law = c.count_law; p.max_ep_len = c.max_ep_len; p.k_max = c.k_max;
    p.standardize = c.standardize;
    const double dy = (double)(p.sy1 - p.sy0);
    p.max_dist = sqrt(dy * dy);
    p.inv_max_dist = 1.0 / p.max_dist;
    p.inv_scale = 1.0 / (double)p.sy1;
    return p;
}

struct EnvView {
    Col<int4> rects;
    Col<double> dsrc;
    int num_obs;
    int sx, sy;          // source
    int intensity, bkg;
};

// ---------------------------------------------------------------------------------------------------------------
// shortest path source -> p around the rectangles (R:491-493) from the per-episode table dsrc[corner]
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool visible(const EnvView &e, int px, int py, int qx, int qy) {
    // the open segment can only meet an open rectangle whose box overlaps the segment's box; each lane walks its own
    // (short) list of such rectangles so that the warp stays converged
    const Seg1 s = make_seg1(px, py, qx, qy);
    uint32_t m = 0u;
    for (int k = 0; k < e.num_obs; k++) {
        const int4 r = e.rects[k];
        if (r.x < s.xhi && s.xlo < r.z && r.y < s.yhi && s.ylo < r.w) m |= 1u << k;
    }
    bool hit = false;
    while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        hit = hit || seg_cross_open1(s, e.rects[k]);
    }
    return !hit;
}

__device__ __forceinline__ double dist_int(int dx, int dy) { return sqrt((double)(dx * dx + dy * dy)); }

// Corner i of rectangle r (order p0 (x0,y0), p1 (x0,y1), p2 (x1,y1), p3 (x1,y0)) is hidden from p by its OWN rectangle iff p
// lies strictly on the inner side of both edges that meet there: the segment then runs through the open rectangle just
// before it reaches the corner.  visible() returns false for exactly these, so callers skip the test.
__device__ __forceinline__ bool corner_hidden_by_own_rect(int4 r, int i, int px, int py) {
    const bool in_x = (i < 2) ? px > r.x : px < r.z;
    const bool in_y = (i == 0 || i == 3) ? py > r.y : py < r.w;
    return in_x && in_y;
}

// the same test as a real function: the step kernel calls it from three places (hint, pair, fall-back walk) and its
// instruction footprint, not its call overhead, is what costs there (the kernel is instruction-fetch limited)
__device__ __noinline__ bool visible_call(const int4 *rects, int stride, int num_obs, int px, int py, int qx, int qy) {
    EnvView e;
    e.rects = Col<int4>{const_cast<int4 *>(rects), stride};
    e.num_obs = num_obs;
    return visible(e, px, py, qx, qy);
}

// Straightforward form (used by the reset path): direct segment if visible, else min over corners.
__device__ __forceinline__ double shortest_path(const EnvView &e, int px, int py) {
    if (visible(e, px, py, e.sx, e.sy)) return dist_int(px - e.sx, py - e.sy);
    double best = __longlong_as_double(0x7ff0000000000000LL);
    const int nc = 4 * e.num_obs;
    for (int c = 0; c < nc; c++) {
        const double ds = e.dsrc[c];
        if (!(ds < best)) continue;                     // cannot improve (also skips unreachable corners)
        const int4 r = e.rects[c >> 2];
        const int cx = corner_x(r, c & 3), cy = corner_y(r, c & 3);
        const double cand = ds + dist_int(px - cx, py - cy);
        if (cand < best && visible(e, px, py, cx, cy)) best = cand;
    }
    return best;
}

// One pass over the rectangles for the segment detector -> source: `direct` = mutually visible (shortest path is the
// segment), `blocked` = boundary_distance < 0.001 for some rectangle (R:1139-1141, without the isclose clause).
__device__ __forceinline__ void source_segment(const EnvView &e, int px, int py, bool &direct, bool &blocked) {
    const Seg1 s = make_seg1(px, py, e.sx, e.sy);
    const int dx = s.dx, dy = s.dy;
    const int l2 = dx * dx + dy * dy;
    const int xlo = s.xlo - 1, xhi = s.xhi + 1, ylo = s.ylo - 1, yhi = s.yhi + 1;
    bool vis_ok = true, blk = false;
    uint32_t m = 0u;
    for (int k = 0; k < e.num_obs; k++) {
        const int4 r = e.rects[k];
        if (r.x <= xhi && xlo <= r.z && r.y <= yhi && ylo <= r.w) m |= 1u << k;
    }
    while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        const int4 r = e.rects[k];
        int cr[4];
        bool open, closed;
        seg_both1(s, r, open, closed, cr);
        vis_ok = vis_ok && !open;
        bool b = closed && !(in_rect_open(px, py, r) && in_rect_open(e.sx, e.sy, r));
        // near-corner clause: only for |pq| > 1000
        if (!b && l2 > 1000000) b = corner_grazes(px, py, dx, dy, l2, r, cr);
        blk = blk || b;
    }
    direct = vis_ok;
    blocked = blk;
}

// Hot-path form: the same minimum, found with few visibility tests.  drow = the env's source-distance row (4 doubles
// per rectangle, 16-byte aligned).  `hint` is the corner that was optimal at the previous step (any value is allowed:
// it only seeds the upper bound `best`).  A corner c can improve on `best` only if
//   (1) the path can bend tautly around its rectangle there: seen from p, both edges of the rectangle at that corner lie
//       on one (closed) side of the line p -> corner, i.e. u.x*u.y <= 0 at p0/p2 and >= 0 at p1/p3 (u = corner - p).
//       A bend at any other visible corner can be cut short by >= 1e-8 (lattice geometry), far above the fp64 rounding
//       of the sums, so dropping those corners cannot change the minimum;
//   (2) dsrc[c] + max(|u.x|, |u.y|) < best: an exact (integer -> double) lower bound of its candidate dsrc[c] + |u|.
// Pass A marks such corners in a per-thread bit mask with converged, branch-free code; pass B walks the thread's own
// few marked corners (exact candidate, then the visibility test only if it would improve).  Exactly the value of
// shortest_path().
// Pass A (+ the hint): returns the bit mask of the corners that may still improve on `best`; best / besti = the value
// through the hint corner (inf / -1 when the hint is unusable).
__device__ __forceinline__ uint32_t sp_seed_and_mask(const EnvView &e, const double *drow, int px, int py, int hint,
                                                     double &best, int &besti) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int nc = 4 * e.num_obs;
    best = inf;
    besti = -1;
    if (hint < nc) {
        const int4 r = e.rects[hint >> 2];
        const int cx = corner_x(r, hint & 3), cy = corner_y(r, hint & 3);
        const double ds = drow[hint];
        if (ds < inf && visible_call(e.rects.p, e.rects.stride, e.num_obs, px, py, cx, cy)) {
            best = ds + dist_int(px - cx, py - cy);
            besti = hint;
        }
    }
    uint32_t mask = 0u;
    for (int k = 0; k < e.num_obs; k++) {
        const int4 r = e.rects[k];
        const double2 d01 = reinterpret_cast<const double2 *>(drow)[2 * k];
        const double2 d23 = reinterpret_cast<const double2 *>(drow)[2 * k + 1];
        const int ux0 = r.x - px, ux1 = r.z - px, uy0 = r.y - py, uy1 = r.w - py;
        const int ax0 = abs(ux0), ax1 = abs(ux1), ay0 = abs(uy0), ay1 = abs(uy1);
        uint32_t m4 = 0u;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int ux = (i < 2) ? ux0 : ux1, uy = (i == 0 || i == 3) ? uy0 : uy1;
            const int linf = max((i < 2) ? ax0 : ax1, (i == 0 || i == 3) ? ay0 : ay1);
            const double ds = i == 0 ? d01.x : (i == 1 ? d01.y : (i == 2 ? d23.x : d23.y));
            const int pr = ux * uy;
            const bool tangent = (i & 1) ? (pr >= 0) : (pr <= 0);
            if (tangent && ds + (double)linf < best) m4 |= 1u << i;
        }
        mask |= m4 << (4 * k);
    }
    if (besti >= 0) mask &= ~(1u << besti);
    return mask;
}

// Pass B for one marked corner: its exact candidate if it beats `best` and is visible from p, else inf.
__device__ __forceinline__ double sp_corner_candidate(const EnvView &e, const double *drow, int px, int py, int c,
                                                      double best) {
    const int4 r = e.rects[c >> 2];
    const int cx = corner_x(r, c & 3), cy = corner_y(r, c & 3);
    const double cand = drow[c] + dist_int(px - cx, py - cy);
    return (cand < best && visible_call(e.rects.p, e.rects.stride, e.num_obs, px, py, cx, cy))
               ? cand : __longlong_as_double(0x7ff0000000000000LL);
}

__device__ __forceinline__ double shortest_path_pruned(const EnvView &e, const double *drow, int px, int py,
                                                       int &hint) {
    double best;
    int besti;
    uint32_t mask = sp_seed_and_mask(e, drow, px, py, hint, best, besti);
    while (mask) {
        const int c = __ffs(mask) - 1;
        mask &= mask - 1;
        const double cand = sp_corner_candidate(e, drow, px, py, c, best);
        if (cand < best) { best = cand; besti = c; }
    }
    if (besti >= 0) hint = besti;
    return best;
}

// the leftover `not isclose(sqrt(euc_dist), sp_dist, abs_tol=0.1)` clause of is_intersect (R:1141-1143): it can only
// hold for euc_dist <= 2 because sp_dist >= euc_dist
__device__ __forceinline__ bool isclose_quirk(double euc, double sp) {
    if (euc > 2.0 && sp >= euc) return false;
    const double a = sqrt(euc), b = sp;
    const double diff = fabs(a - b), big = fmax(fabs(a), fabs(b));
    const double tol = fmax(1e-09 * big, 0.1);
    return a == b || (isfinite(a) && isfinite(b) && diff <= tol);
}

// is_intersect R:1133-1146
__device__ __forceinline__ bool los_blocked(const EnvView &e, int px, int py, double euc, double sp) {
    if (isclose_quirk(euc, sp)) return false;
    bool direct, blocked;
    source_segment(e, px, py, direct, blocked);
    return blocked;
}

// ---------------------------------------------------------------------------------------------------------------
// obstruction_sensors R:1172-1261 and correct_coords R:1263-1306
// ---------------------------------------------------------------------------------------------------------------
// closed ray [p, p+step(d)] touches the axis-aligned lattice edge; d2 = squared distance p -> edge (clamped projection)
__device__ __forceinline__ bool ray_hits_vedge(int px, int py, int sx, int sy, int c, int ya, int yb) {
    // edge x = c, y in [ya, yb]
    if (sx == 0) {
        if (px != c) return false;
        const int lo = min(py, py + sy), hi = max(py, py + sy);
        return lo <= yb && ya <= hi;
    }
    const int t = c - px;                           // need t/sx in [0,1]
    if (sx > 0 ? (t < 0 || t > sx) : (t > 0 || t < sx)) return false;
    // y at the crossing: sy is 0 or +-|sx|
    const int yat = py + (sy == 0 ? 0 : ((sy > 0) == (sx > 0) ? t : -t));
    return ya <= yat && yat <= yb;
}
__device__ __forceinline__ int clampdist(int v, int a, int b) { return v < a ? a - v : (v > b ? v - b : 0); }

__device__ __noinline__ void correct_coords(int px, int py, int4 r, float out[8], uint32_t &status) {
    int lo_d[8];
    int nstar = 0x7fffffff;
#pragma unroll
    for (int d = 0; d < 8; d++) {
        const int cx = coef_x(d), cy = coef_y(d);
        int lo = 1, hi = 0x7fffffff;
        bool ok = true;
        const int X = 10 * px, Y = 10 * py;
        if (cx == 0) ok = ok && (10 * r.x <= X && X <= 10 * r.z);
        else if (cx > 0) { lo = max(lo, 10 * r.x - X); hi = min(hi, 10 * r.z - X); }
        else { lo = max(lo, X - 10 * r.z); hi = min(hi, X - 10 * r.x); }
        if (cy == 0) ok = ok && (10 * r.y <= Y && Y <= 10 * r.w);
        else if (cy > 0) { lo = max(lo, 10 * r.y - Y); hi = min(hi, 10 * r.w - Y); }
        else { lo = max(lo, Y - 10 * r.w); hi = min(hi, Y - 10 * r.y); }
        ok = ok && lo <= hi;
        lo_d[d] = ok ? lo : 0x7fffffff;
        nstar = min(nstar, lo_d[d]);
    }
#pragma unroll
    for (int d = 0; d < 8; d++) out[d] = 0.0f;
    if (nstar == 0x7fffffff) { status |= RS_ST_CORRECT_MISS; return; }
    int xc = 0, cnt = 0;
#pragma unroll
    for (int d = 0; d < 8; d++)
        if (lo_d[d] == nstar) { xc |= 1 << d; cnt++; }
    if (cnt >= 4) {
#pragma unroll
        for (int ii = 0; ii <= 6; ii += 2) {
            const int lo = (ii + 7) & 7, hi = ii + 1;
            if (((xc >> lo) & 1) && ((xc >> hi) & 1)) { out[ii] = 1.0f; out[lo] = 1.0f; out[hi] = 1.0f; }
        }
    }
}

// candidate rectangles of the sensors: a ray is at most 100 (71 per axis) long
__device__ __forceinline__ int sensor_candidates(const EnvView &e, int px, int py) {
    int cand = 0;
    for (int k = 0; k < e.num_obs; k++) {
        const int4 r = e.rects[k];
        if (r.x - 100 <= px && px <= r.z + 100 && r.y - 100 <= py && py <= r.w + 100) cand |= 1 << k;
    }
    return cand;
}

// the obstruction part of obstruction_sensors (R:1186-1226) over the candidate rectangles `cand`
__device__ __forceinline__ void sensors_rects(const EnvView &e, int px, int py, int cand, float out[8],
                                              uint32_t &status) {
    // squared distance of the best scored edge per direction; -1 = no hit
    int best_d2[8];
#pragma unroll
    for (int d = 0; d < 8; d++) best_d2[d] = -1;
    if (cand) {
        unsigned long long hits = 0ull;              // 8 bits per rectangle (obs_idx_ls R:1190)
        // per-direction running state across this lane's candidate rectangles (index order, R:1186-1217)
        int inter_d[8], dmin_d[8];
#pragma unroll
        for (int d = 0; d < 8; d++) { inter_d[d] = 0; dmin_d[d] = -1; }
        int todo = cand;
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const int4 r = e.rects[k];
            int hk = 0;
#pragma unroll
            for (int d = 0; d < 8; d++) {
                const int sx = step_dx(d), sy = step_dy(d);
                int inter = inter_d[d], dmin = dmin_d[d];
                // edge order R:1000-1006: (p0,p1) left, (p0,p3) bottom, (p2,p1) top, (p2,p3) right
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    bool hit;
                    int d2;
                    if (s == 0 || s == 3) {
                        const int c = (s == 0) ? r.x : r.z;
                        hit = ray_hits_vedge(px, py, sx, sy, c, r.y, r.w);
                        const int ddx = px - c, ddy = clampdist(py, r.y, r.w);
                        d2 = ddx * ddx + ddy * ddy;
                    } else {
                        const int c = (s == 1) ? r.y : r.w;
                        hit = ray_hits_vedge(py, px, sy, sx, c, r.x, r.z);      // transposed
                        const int ddy = py - c, ddx = clampdist(px, r.x, r.z);
                        d2 = ddx * ddx + ddy * ddy;
                    }
                    if (inter < 2 && hit) {
                        dmin = (dmin < 0 || d2 < dmin) ? d2 : dmin;
                        inter++;
                        hk++;
                    }
                }
                inter_d[d] = inter; dmin_d[d] = dmin;
            }
            hits |= (unsigned long long)hk << (8 * k);
        }
#pragma unroll
        for (int d = 0; d < 8; d++) best_d2[d] = dmin_d[d];
        int ones = 0;
#pragma unroll
        for (int d = 0; d < 8; d++) ones += (best_d2[d] == 0);
        if (ones > 3) {
            // max(zip(obs_idx_ls, self.poly)) R:1222-1226: most hits, ties -> lexicographically largest vertex list
            int hits_best = -1, best_k = 0;
            for (int k = 0; k < e.num_obs; k++) {
                const int hk = (int)((hits >> (8 * k)) & 0xffull);
                bool take = hk > hits_best;
                if (!take && hk == hits_best) {
                    const int4 a = e.rects[k], b = e.rects[best_k];
                    take = (a.x != b.x) ? (a.x > b.x) : ((a.y != b.y) ? (a.y > b.y) : ((a.w != b.w) ? (a.w > b.w) : (a.z > b.z)));
                }
                if (take) { hits_best = hk; best_k = k; }
            }
            correct_coords(px, py, e.rects[best_k], out, status);
#pragma unroll
            for (int d = 0; d < 8; d++) best_d2[d] = -2;      // already final
        }
    }
#pragma unroll
    for (int d = 0; d < 8; d++) {
        // (110 - dist)/110; 0 = no hit, exactly 1 on the edge; straight-line code (MUFU.RSQ, ~3e-7 relative)
        const float f2 = (float)max(best_d2[d], 1);
        const float v = (110.0f - f2 * rsqrtf(f2)) * (1.0f / 110.0f);
        if (best_d2[d] != -2) out[d] = best_d2[d] < 0 ? 0.0f : (best_d2[d] == 0 ? 1.0f : v);
    }
}

// The same for the step kernel, four directions at a time: half the instruction footprint of the fully unrolled form
// (the kernel is instruction-fetch sensitive) while a thread still has four independent dependency chains in flight (the
// plain direction-major loop of the reference, one chain, costs 1.5 k cycles more per CTA).  The eight values go straight
// to the observation row (shared memory).
__device__ __forceinline__ void sensors_rects_row(const EnvView &e, int px, int py, int cand, float *row8,
                                                  uint32_t &status) {
    unsigned long long hits = 0ull;                  // 8 bits per rectangle (obs_idx_ls R:1190)
    int ones = 0;
#pragma unroll 1
    for (int d0 = 0; d0 < 8; d0 += 4) {
        int inter[4] = {0, 0, 0, 0}, dmin[4] = {-1, -1, -1, -1};
        int todo = cand;
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const int4 r = e.rects[k];
            int hk = 0;
#pragma unroll
            for (int dd = 0; dd < 4; dd++) {
                const int sx = step_dx(d0 + dd), sy = step_dy(d0 + dd);
                // edge order R:1000-1006: (p0,p1) left, (p0,p3) bottom, (p2,p1) top, (p2,p3) right
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    bool hit;
                    int d2;
                    if (s == 0 || s == 3) {
                        const int c = (s == 0) ? r.x : r.z;
                        hit = ray_hits_vedge(px, py, sx, sy, c, r.y, r.w);
                        const int ddx = px - c, ddy = clampdist(py, r.y, r.w);
                        d2 = ddx * ddx + ddy * ddy;
                    } else {
                        const int c = (s == 1) ? r.y : r.w;
                        hit = ray_hits_vedge(py, px, sy, sx, c, r.x, r.z);      // transposed
                        const int ddy = py - c, ddx = clampdist(px, r.x, r.z);
                        d2 = ddx * ddx + ddy * ddy;
                    }
                    if (inter[dd] < 2 && hit) {
                        dmin[dd] = (dmin[dd] < 0 || d2 < dmin[dd]) ? d2 : dmin[dd];
                        inter[dd]++;
                        hk++;
                    }
                }
            }
            hits += (unsigned long long)hk << (8 * k);
        }
#pragma unroll
        for (int dd = 0; dd < 4; dd++) {
            ones += (dmin[dd] == 0);
            const float f2 = (float)max(dmin[dd], 1);
            const float v = (110.0f - f2 * rsqrtf(f2)) * (1.0f / 110.0f);
            row8[d0 + dd] = dmin[dd] < 0 ? 0.0f : (dmin[dd] == 0 ? 1.0f : v);
        }
    }
    if (ones > 3) {
        // max(zip(obs_idx_ls, self.poly)) R:1222-1226: most hits, ties -> lexicographically largest vertex list
        int hits_best = -1, best_k = 0;
        for (int k = 0; k < e.num_obs; k++) {
            const int hk = (int)((hits >> (8 * k)) & 0xffull);
            bool take = hk > hits_best;
            if (!take && hk == hits_best) {
                const int4 a = e.rects[k], b = e.rects[best_k];
                take = (a.x != b.x) ? (a.x > b.x) : ((a.y != b.y) ? (a.y > b.y) : ((a.w != b.w) ? (a.w > b.w) : (a.z > b.z)));
            }
            if (take) { hits_best = hk; best_k = k; }
        }
        float out[8];
        correct_coords(px, py, e.rects[best_k], out, status);
#pragma unroll
        for (int d = 0; d < 8; d++) row8[d] = out[d];
    }
}

// the wall part of obstruction_sensors (enforce_grid_boundaries, R:1232-1259); out = the 8 sensor values
__device__ __forceinline__ void sensors_walls(const Params &P, int px, int py, float *out, uint32_t &status) {
    {
        if (px - 110 < P.bx0) { if (out[0] != 0.0f) status |= RS_ST_WALL_ASSERT; out[0] = __fdiv_rn(110.0f - fabsf((float)(px - P.bx0)), 110.0f); }
        if (py - 110 < P.by0) { if (out[6] != 0.0f) status |= RS_ST_WALL_ASSERT; out[6] = __fdiv_rn(110.0f - fabsf((float)(py - P.by0)), 110.0f); }
        if (P.bx1 <= px + 110) { if (out[4] != 0.0f) status |= RS_ST_WALL_ASSERT; out[4] = __fdiv_rn(110.0f - fabsf((float)(P.bx1 - px)), 110.0f); }
        if (P.by1 <= py + 110) { if (out[2] != 0.0f) status |= RS_ST_WALL_ASSERT; out[2] = __fdiv_rn(110.0f - fabsf((float)(P.by1 - py)), 110.0f); }
    }
}

__device__ __forceinline__ void sensors(const Params &P, const EnvView &e, int px, int py, float out[8],
                                        uint32_t &status) {
    sensors_rects(e, px, py, sensor_candidates(e, px, py), out, status);
    if (P.enforce) sensors_walls(P, px, py, out, status);
}

// in_obstruction R:1148-1170
__device__ __forceinline__ bool in_obstruction(const EnvView &e, int px, int py) {
    bool found = false, blocked = false;
    for (int k = 0; k < e.num_obs; k++) {
        const int4 r = e.rects[k];
        if (!found && in_rect_closed(px, py, r)) { found = true; blocked = in_rect_open(px, py, r); }
    }
    return blocked;
}

// measurement + sensors + observation row for one agent at (px,py).  R:498-502, 570-593
template <bool kFast>
__device__ __forceinline__ void observe(const Params &P, const EnvView &e, int px, int py, double euc, bool blocked_los,
                                        Rng &g, float *obs_row, uint32_t &status) {
    double lam;
    if (blocked_los) lam = (double)e.bkg;
    else {
        double d = euc;
        if (d == 0.0) { status |= RS_ST_LAMBDA_INF; d = 1.0; }
        lam = (P.count_law == 1) ? (double)e.intensity / (d * d) + (double)e.bkg : (double)e.intensity / d + (double)e.bkg;
    }
    const long long cnt = poisson<kFast>(g, lam);
    status |= g.status;
    float s[8];
    if (e.num_obs > 0 || P.enforce) sensors(P, e, px, py, s, status);
    else {
#pragma unroll
        for (int d = 0; d < 8; d++) s[d] = 0.0f;
    }
    obs_row[0] = (float)cnt;
    obs_row[1] = (float)((double)px * P.inv_scale);
    obs_row[2] = (float)((double)py * P.inv_scale);
#pragma unroll
    for (int d = 0; d < 8; d++) obs_row[3 + d] = s[d];
}

// ---------------------------------------------------------------------------------------------------------------
// Per-episode running standardisation of the count channel (RsConfig.standardize):
// StatisticStandardization.update + standardize (algos/multiagent/NeuralNetworkCores/RADTEAM_core.py:215-265) in the
// caller's order (train.py:311, 339, 436, 469, 548: update with a reading, then standardize that same reading);
// mode 2 = StatBuff.update (algos/test_environment/core.py:62-73) and np.clip(z, -8, 8) (test_environment/ppo.py:502).
// n = number of readings of the episode including x; update == false only standardizes (step(None) probe).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double stat_std(int mode, int n, double m2) {
    if (n <= 1) return 1.0;                                              // defaults after reset(): std = 1
    const double sd = sqrt(m2 / (double)(n - 1));                        // sample variance
    return mode == 2 ? (sd == 0.0 ? 1.0 : sd) : fmax(sd, 1.0);
}
__device__ __forceinline__ double stat_standardize(int mode, int n, double x, double &mean, double &m2, bool update) {
    if (update) {
        if (n <= 1) { mean = x; m2 = 0.0; }
        else {
            const double mean_new = mean + (x - mean) / (double)n;
            m2 = m2 + (x - mean) * (x - mean_new);
            mean = mean_new;
        }
    }
    double z = (x - mean) / stat_std(mode, n, m2);
    if (mode == 2) z = fmin(fmax(z, -8.0), 8.0);
    return z;
}

// ---------------------------------------------------------------------------------------------------------------
// RadSearch.step R:443-728 for one environment (all agents), plus the caller rules T:394-405 when auto-reset is on
// ---------------------------------------------------------------------------------------------------------------
struct StepArgs {
    const int32_t *actions;
    float *obs, *reward, *team_reward, *final_obs;
    uint8_t *done, *info, *ended;
    int n_env;
    uint32_t env_id0;
    uint64_t seed, step_ctr;
    const double *uniforms;
    int n_uniforms, flags;
    int parity;         // refill list that consumed prefetch slots are pushed to
};

// ---------------------------------------------------------------------------------------------------------------
// reset R:730-797: scenario sampling (Philox domain 1), per-episode tables, initial observation (step(None) probe).
// Cooperative form: `nl` lanes (32, 8 or 1 on the GPU depending on how many envs reset; 1 in the host emulation)
// share one environment; sync_mask names them.  The sequential
// rejection sampling runs redundantly on every lane (same Philox stream, no divergence); the per-corner work
// (visibility rows, source visibility, Dijkstra relaxations, table stores) is strided over the lanes through the
// warp's shared-memory scratch, separated by __syncwarp().
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rand_point(const Params &P, Rng &g, int &x, int &y) {     // R:1026-1036
    const uint32_t span = (uint32_t)(P.sx1 - P.sx0);
    x = P.sx0 + (int)g.below(span);
    y = P.sx0 + (int)g.below(span);
}

__device__ __forceinline__ int create_obstructions(const Params &P, Rng &g, Col<int4> rects, uint32_t &status) {
    const int hx = (int)((double)P.sx1 * 0.9), hy = (int)((double)P.sy1 * 0.9);            // R:961-966
    int num_obs = 0;
    for (int attempt = 0; attempt < 256; attempt++) {
        num_obs = (P.obstruction_count == -1) ? 1 + (int)g.below(5) : P.obstruction_count;  // R:745-750
        int ii = 0, tries = 0;
        while (ii < num_obs && tries < 4096) {
            tries++;
            const int sx = P.sx0 + (int)g.below((uint32_t)(hx - P.sx0));
            const int sy = P.sy0 + (int)g.below((uint32_t)(hy - P.sy0));
            const int ex = P.lo + (int)g.below((uint32_t)(P.hi - P.lo));
            const int ey = P.lo + (int)g.below((uint32_t)(P.hi - P.lo));
            const int4 r = make_int4(sx, sy, sx + ex, sy + ey);
            bool touch = false;
            for (int kk = 0; kk < ii; kk++) touch = touch || rects_touch(rects[kk], r);       // R:985-992
            if (!touch) { rects[ii] = r; ii++; }
        }
        if (ii < num_obs) { status |= RS_ST_REJECT_CAP; num_obs = ii; }
        bool nested = false;                                                                // is_valid R:788-791
        for (int i = 0; i < num_obs; i++)
            for (int j = i + 1; j < num_obs; j++) nested = nested || rects_nested(rects[i], rects[j]);
        if (!nested) return num_obs;
    }
    status |= RS_ST_REJECT_CAP;
    return num_obs;
}

__device__ __forceinline__ bool los_blocked_any(const EnvView &e, int px, int py, int qx, int qy) {
    bool hit = false;
    for (int k = 0; k < e.num_obs; k++) hit = hit || los_blocked_rect(px, py, qx, qy, e.rects[k]);
    return hit;
}

__device__ __forceinline__ void sample_source_loc_pos(const Params &P, Rng &g, const EnvView &e, int &sx, int &sy,
                                                      int &dx_, int &dy_, uint32_t &status) {
    int srcx, srcy, detx, dety;
    rand_point(P, g, srcx, srcy);
    rand_point(P, g, detx, dety);
    int tries = 0;
    for (;;) {                                                                               // R:1057-1076
        bool inside = false;
        for (int k = 0; k < e.num_obs; k++) inside = inside || in_rect_closed(detx, dety, e.rects[k]);
        if (!inside) break;
        if (++tries > 100000) { status |= RS_ST_REJECT_CAP; break; }
        rand_point(P, g, detx, dety);
    }
    int num_retry = 0;
    tries = 0;
    for (;;) {                                                                               // R:1091-1129
        for (;;) {
            const int ddx = detx - srcx, ddy = dety - srcy;
            if (ddx * ddx + ddy * ddy >= 1000000) break;
            if (++tries > 100000) { status |= RS_ST_REJECT_CAP; break; }
            rand_point(P, g, srcx, srcy);
        }
        bool resamp = false, inter = false;
        for (int k = 0; k < e.num_obs; k++) {
            if (resamp) break;
            const int4 r = e.rects[k];
            if (in_rect_closed(srcx, srcy, r)) resamp = true;
            if (!resamp && los_blocked_rect(detx, dety, srcx, srcy, r)) inter = true;
        }
        if (e.num_obs == 0 || (num_retry > 20 && !resamp)) break;
        else if (resamp || !inter) { rand_point(P, g, srcx, srcy); num_retry++; }
        else break;
        if (++tries > 100000) { status |= RS_ST_REJECT_CAP; break; }
    }
    sx = srcx; sy = srcy; dx_ = detx; dy_ = dety;
}

struct ResetArgs {
    float *obs;
    int n_env;
    uint32_t env_id0;
    uint64_t seed, step_ctr;
    const double *uniforms;
    int n_uniforms;
    // scenario injection (all nullptr when sampling)
    const int32_t *in_src, *in_det, *in_intensity, *in_bkg, *in_rects, *in_num_obs;
    int k_in;
    int prepare;        // 1: write the scenario of episode epi[n]+1 into the prefetch buffers (RsState.nx_*) only
    int parity;         // which refill list a synchronous reset pushes the env to (-1: none)
};

// shortest path source -> (px,py) of env n, for rs_query_shortest_path
__device__ __forceinline__ double query_sp(const RsState &S, int n, int N, int k_max, int px, int py, int variant,
                                           Col<int4> rects) {
    EnvView e;
    e.rects = rects;
    e.dsrc = Col<double>{S.dsrc + (size_t)n * 4 * k_max, 1};          // env-major table row
    const int meta = S.meta[n];
    e.num_obs = meta & 0xff;
    const int2 src = reinterpret_cast<const int2 *>(S.src)[n];
    e.sx = src.x; e.sy = src.y; e.intensity = 0; e.bkg = 0;
    for (int k = 0; k < e.num_obs; k++) rects[k] = reinterpret_cast<const int4 *>(S.rects)[(size_t)k * N + n];
    if (variant == 1) return shortest_path(e, px, py);
    bool direct, blocked;
    source_segment(e, px, py, direct, blocked);
    int hint = 31;
    return direct ? dist_int(px - e.sx, py - e.sy) : shortest_path_pruned(e, e.dsrc.p, px, py, hint);
}

#ifdef RS_HOST_EMU
#define RS_SYNCWARP(m)
#else
#define RS_SYNCWARP(m) __syncwarp(m)
#endif

template <bool kFast>
__device__ __forceinline__ void reset_env(const Params &P, const RsState &S, const ResetArgs &a, int n,
                                          bool new_obstacles, int lane, int nl, uint32_t sync_mask,
                                          Col<int4> w_rects, Col<double> w_dsrc, Col<uint32_t> w_vis) {
    const int N = a.n_env, A = P.n_agents;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    uint32_t status = 0;
    EnvView e;
    e.rects = w_rects;
    e.dsrc = w_dsrc;
    // the draws of a reset are keyed by the env's episode number, not by wall-clock step: the scenario of episode e
    // is a pure function of (seed, env id, e, obstructions), whoever computes it and whenever
    const uint32_t ep_seq = S.epi[n] + 1u;
    const bool prepare = a.prepare != 0;
    const bool inject = a.in_src != nullptr;
    if (!prepare && !inject && !new_obstacles && a.parity >= 0) {
        // RS_F_PREFETCH: rs_prepare may already have computed this very episode (same seed, env, episode number,
        // obstructions): adopt it -- a handful of copies by one lane -- instead of recomputing it
        const uint32_t tag = *reinterpret_cast<volatile const uint32_t *>(S.nx_seq + n);
        if (tag == ep_seq) {
            __threadfence();
            {                                           // the prefetched source-distance row (env-major, 4K doubles)
                const int nc0 = 4 * (S.meta[n] & 0xff);
                const size_t row = (size_t)n * 4 * P.k_max;
                for (int c = lane; c < nc0; c += nl) {
                    S.dsrc[row + c] = S.nx_dsrc[row + c];
                    S.dsf[row + c] = S.nx_dsf[row + c];
                }
            }
            if (lane == 0) {
                const int2 s0 = reinterpret_cast<const int2 *>(S.nx_src)[n];
                const int2 r0 = reinterpret_cast<const int2 *>(S.nx_rad)[n];
                const int2 d0 = reinterpret_cast<const int2 *>(S.nx_det)[n];
                const double b0 = S.nx_best[n];
                const int meta = S.meta[n];
                const float *nxo = S.nx_obs + (size_t)n * A * RS_OBS_DIM;
                float *dst = a.obs + (size_t)n * A * RS_OBS_DIM;
                for (int ag = 0; ag < A; ag++) {
                    float row[RS_OBS_DIM];
#pragma unroll
                    for (int i = 0; i < RS_OBS_DIM; i++) row[i] = nxo[ag * RS_OBS_DIM + i];
                    const size_t ia = (size_t)ag * N + n;
                    if (P.standardize) {                                   // first reading of the episode: z = 0
                        S.st_mean[ia] = (double)row[0];
                        S.st_m2[ia] = 0.0;
                        if (S.raw_count) S.raw_count[(size_t)n * A + ag] = row[0];
                        row[0] = 0.0f;
                    }
#pragma unroll
                    for (int i = 0; i < RS_OBS_DIM; i++) dst[ag * RS_OBS_DIM + i] = row[i];
                    reinterpret_cast<int2 *>(S.det)[ia] = d0;
                    S.best[ia] = b0;
                    S.aflags[ia] = (r0.y >> 8) << 25;                     // the prefetched episode's search seed
                }
                reinterpret_cast<int2 *>(S.src)[n] = s0;
                reinterpret_cast<int2 *>(S.rad)[n] = make_int2(r0.x, r0.y & 0xff);
                S.meta[n] = meta & 0xff;                                  // done = 0, ep_len = 0
                S.epi[n] = ep_seq;
                // an env is listed once per block of steps (episodes outlast a block: RadSearch refuses prefetch
                // otherwise); the bound keeps a caller that breaks that rule from writing past its list
                const int slot = atomicAdd(S.refill_count + a.parity, 1);
                if (slot < N) S.refill_list[(size_t)a.parity * N + slot] = n;
                else atomicOr(S.status + n, RS_ST_REFILL_OVERFLOW);
            }
            return;
        }
    }
    Rng g;
    g.init_philox(a.seed, a.env_id0 + (uint32_t)n, 1, 0, (uint64_t)ep_seq);
    RS_SYNCWARP(sync_mask);                                                      // scratch is reused between environments
    if (inject) {                                                       // refresh_environment R:799-874
        e.num_obs = min(a.in_num_obs[n], P.k_max);
        for (int k = lane; k < e.num_obs; k += nl)
            w_rects[k] = reinterpret_cast<const int4 *>(a.in_rects)[(size_t)n * a.k_in + k];
        new_obstacles = true;
    } else if (new_obstacles) {
        e.num_obs = create_obstructions(P, g, e.rects, status);         // R:744-762 (every lane writes the same values)
    } else {
        e.num_obs = S.meta[n] & 0xff;
        for (int k = lane; k < e.num_obs; k += nl) w_rects[k] = reinterpret_cast<const int4 *>(S.rects)[(size_t)k * N + n];
        for (int c = lane; c < 4 * e.num_obs; c += nl) w_vis[c] = S.vis[(size_t)c * N + n];
    }
    RS_SYNCWARP(sync_mask);
    const int nc = 4 * e.num_obs;
    if (new_obstacles) {
        // corner-to-corner visibility (depends on the obstructions only).  The predicate is exact integer arithmetic and
        // therefore symmetric: every unordered pair is tested once and sets both bits.  Two corners of the SAME sampled
        // rectangle need no test: the sampler keeps the rectangles apart (R:985-992, is_valid R:788-791), so an edge is
        // always free and a diagonal always crosses its own interior (an injected scenario may break that: tested there).
        const bool own_known = !inject && !(status & RS_ST_REJECT_CAP);
        for (int c = lane; c < nc; c += nl) w_vis[c] = 0u;
        RS_SYNCWARP(sync_mask);
        for (int c = 0; c + 1 < nc; c++) {
            const int4 rc = w_rects[c >> 2];
            const int cx = corner_x(rc, c & 3), cy = corner_y(rc, c & 3);
            uint32_t m = 0u;
            for (int c2 = c + 1 + lane; c2 < nc; c2 += nl) {
                bool v;
                if (own_known && (c2 >> 2) == (c >> 2)) v = ((c ^ c2) & 1) != 0;      // neighbours along an edge
                else {
                    const int4 r2 = w_rects[c2 >> 2];
                    v = visible(e, cx, cy, corner_x(r2, c2 & 3), corner_y(r2, c2 & 3));
                }
                if (v) {
                    m |= 1u << c2;
                    if (nl == 1) w_vis[c2] |= 1u << c;
                    else atomicOr(&w_vis[c2], 1u << c);
                }
            }
            if (nl == 1) w_vis[c] |= m;
            else if (m) atomicOr(&w_vis[c], m);
        }
        RS_SYNCWARP(sync_mask);
        for (int c = lane; c < nc; c += nl) S.vis[(size_t)c * N + n] = w_vis[c];
        for (int k = lane; k < e.num_obs; k += nl) reinterpret_cast<int4 *>(S.rects)[(size_t)k * N + n] = w_rects[k];
    }
    int detx, dety;
    if (inject) {
        e.sx = a.in_src[2 * n]; e.sy = a.in_src[2 * n + 1];
        detx = a.in_det[2 * n]; dety = a.in_det[2 * n + 1];
        e.intensity = a.in_intensity[n]; e.bkg = a.in_bkg[n];
    } else {
        sample_source_loc_pos(P, g, e, e.sx, e.sy, detx, dety, status);  // R:764-769
        e.intensity = 1000000 + (int)g.below(9000000u);                  // R:778
        e.bkg = 10 + (int)g.below(41u);                                  // R:779
    }
    // dsrc[c] = shortest path length source -> corner c: Dijkstra on the corner visibility graph, sums accumulated
    // from the source outwards (the order Polyline::length() adds them)
    for (int c = lane; c < nc; c += nl) {
        const int4 r = w_rects[c >> 2];
        const int cx = corner_x(r, c & 3), cy = corner_y(r, c & 3);
        w_dsrc[c] = (!corner_hidden_by_own_rect(r, c & 3, e.sx, e.sy) && visible(e, e.sx, e.sy, cx, cy))
                        ? dist_int(cx - e.sx, cy - e.sy) : inf;
    }
    RS_SYNCWARP(sync_mask);
    uint32_t fin = 0;
    for (int it = 0; it < nc; it++) {
        int u = -1;
        double du = inf;
#ifndef RS_HOST_EMU
        if (nl == 32) {
            // lane c holds corner c: butterfly min over (distance, index), lowest index on ties like the scan below
            if (lane < nc && !((fin >> lane) & 1)) { du = w_dsrc[lane]; if (du < inf) u = lane; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, du, o);
                const int ou = __shfl_xor_sync(0xffffffffu, u, o);
                if (ou >= 0 && (u < 0 || od < du || (od == du && ou < u))) { du = od; u = ou; }
            }
        } else
#endif
        {
            uint32_t rem = (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u)) & ~fin;     // unsettled corners only, lowest index
            while (rem) {                                                           // first: every lane scans, uniform result
                const int c = __ffs(rem) - 1;
                rem &= rem - 1;
                const double d = w_dsrc[c];
                if (d < du) { du = d; u = c; }
            }
        }
        if (u < 0) break;
        fin |= 1u << u;
        const int4 ru = w_rects[u >> 2];
        const int ux = corner_x(ru, u & 3), uy = corner_y(ru, u & 3);
        const uint32_t m = w_vis[u] & ~fin;
        RS_SYNCWARP(sync_mask);
        auto relax = [&](int w) {
            const int4 rw = w_rects[w >> 2];
            const int dx = ux - corner_x(rw, w & 3), dy = uy - corner_y(rw, w & 3);
            const double old = w_dsrc[w];
            // |u - w| >= max(|dx|, |dy|), an integer: rounding is monotone, so a label this bound cannot beat stands
            // (most relaxations end here, without the square root)
            if (du + (double)max(abs(dx), abs(dy)) >= old) return;
            const double nd = du + dist_int(dx, dy);
            if (nd < old) w_dsrc[w] = nd;
        };
        if (nl == 1) {
            // one thread per environment: the thread walks ITS corner's open neighbours only (the lanes of a warp are at
            // different corners of different scenes, so a loop over all corners runs as long as the union of their masks)
            uint32_t mm = m;
            while (mm) {
                const int w = __ffs(mm) - 1;
                mm &= mm - 1;
                relax(w);
            }
        } else {
            for (int w = lane; w < nc; w += nl)
                if ((m >> w) & 1u) relax(w);
        }
        RS_SYNCWARP(sync_mask);
    }
    const double euc = dist_int(detx - e.sx, dety - e.sy);
    bool direct, blocked_raw;
    source_segment(e, detx, dety, direct, blocked_raw);
    // prev_det_dist R:771-776 = shortest_path(e, det): the lanes evaluate one corner each, then everybody takes the min
    double *dsrc_out = (prepare ? S.nx_dsrc : S.dsrc) + (size_t)n * 4 * P.k_max;
    float *dsf_out = (prepare ? S.nx_dsf : S.dsf) + (size_t)n * 4 * P.k_max;     // lower bounds for the marking pass
    double lane_best = inf;
    for (int c = lane; c < nc; c += nl) {
        const double ds = w_dsrc[c];
        dsrc_out[c] = ds;
        dsf_out[c] = __double2float_rd(ds);
        const int4 r = w_rects[c >> 2];
        const int cx = corner_x(r, c & 3), cy = corner_y(r, c & 3);
        double cand = inf;
        // a corner whose lower bound ds + max(|dx|, |dy|) does not beat this lane's best so far cannot be the minimum
        if (!direct && ds + (double)max(abs(detx - cx), abs(dety - cy)) < lane_best &&
            !corner_hidden_by_own_rect(r, c & 3, detx, dety) && visible(e, detx, dety, cx, cy)) {
            cand = ds + dist_int(detx - cx, dety - cy);
            lane_best = fmin(lane_best, cand);
        }
        w_dsrc[c] = cand;
    }
    RS_SYNCWARP(sync_mask);
    // the first step's seed for the pruned search: the corner the initial shortest path goes through (aflags bits 25..29)
    double sp = direct ? euc : inf;
    int hint0 = 0;
    if (!direct)
        for (int c = 0; c < nc; c++) {
            const double d = w_dsrc[c];
            if (d < sp) { sp = d; hint0 = c; }
        }
    const bool blocked_los = blocked_raw && !isclose_quirk(euc, sp);
    if (lane == 0) {
        if (prepare) {
            reinterpret_cast<int2 *>(S.nx_src)[n] = make_int2(e.sx, e.sy);
            reinterpret_cast<int2 *>(S.nx_rad)[n] = make_int2(e.intensity, e.bkg | (hint0 << 8));    // bkg < 256 R:779
            reinterpret_cast<int2 *>(S.nx_det)[n] = make_int2(detx, dety);
            S.nx_best[n] = sp;
        } else {
            reinterpret_cast<int2 *>(S.src)[n] = make_int2(e.sx, e.sy);
            reinterpret_cast<int2 *>(S.rad)[n] = make_int2(e.intensity, e.bkg);
            // done = 0, ep_len = 0 R:739-740; bits 9..15: rectangles that hold the source strictly inside (an injected
            // scenario may have them; the sampler rejects such sources R:1091-1129) -- see rs_step1.cuh::source_segment1
            int src_in = 0;
            if (inject)
                for (int k = 0; k < e.num_obs && k < 7; k++) src_in |= (int)in_rect_open(e.sx, e.sy, w_rects[k]) << k;
            S.meta[n] = e.num_obs | (src_in << 9);
            S.epi[n] = ep_seq;
        }
    }
    float *obs_out = prepare ? S.nx_obs : a.obs;
    for (int ag = 0; ag < A; ag++) {
        const size_t ia = (size_t)ag * N + n;
        Rng gp;
        if (a.uniforms) gp.init_inject(a.uniforms + ((size_t)n * A + ag) * a.n_uniforms, a.n_uniforms);
        else gp.init_philox(a.seed, a.env_id0 + (uint32_t)n, 2, (uint32_t)ag, (uint64_t)ep_seq);
        float row[RS_OBS_DIM];
        observe<kFast>(P, e, detx, dety, euc, blocked_los, gp, row, status);   // R:794
        if (lane == 0) {
            if (!prepare) {
                reinterpret_cast<int2 *>(S.det)[ia] = make_int2(detx, dety);
                S.best[ia] = sp;
                S.aflags[ia] = hint0 << 25;                              // Agent.reset R:289-300 (+ the search seed)
                if (P.standardize) {                                     // stat_buffers[id].reset(); .update(obs[0])
                    S.st_mean[ia] = (double)row[0];                      // T:509, 548: first reading, z = 0
                    S.st_m2[ia] = 0.0;
                    if (S.raw_count) S.raw_count[(size_t)n * A + ag] = row[0];
                    row[0] = 0.0f;
                }
            }
#pragma unroll
            for (int i = 0; i < RS_OBS_DIM; i++) obs_out[((size_t)n * A + ag) * RS_OBS_DIM + i] = row[i];
        }
    }
    if (prepare) {
        // publish: the tag is written last, after the data (lanes' table stores included)
        __threadfence();
        RS_SYNCWARP(sync_mask);
        if (lane == 0) {
#ifdef RS_HOST_EMU
            S.nx_seq[n] = ep_seq;
#else
            atomicExch(S.nx_seq + n, ep_seq);
#endif
        }
    } else if (lane == 0 && a.parity >= 0 && S.refill_list) {
        const int slot = atomicAdd(S.refill_count + a.parity, 1);
        if (slot < N) S.refill_list[(size_t)a.parity * N + slot] = n;
        else status |= RS_ST_REFILL_OVERFLOW;
    }
    status |= g.status;
    if (status && lane == 0) S.status[n] |= status;
}

}  // namespace rs
