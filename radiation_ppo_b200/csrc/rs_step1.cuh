// RadSearch.step (R:443-728) + caller rules (T:394-405) for the single-agent case (number_agents == 1, every BASELINE
// configuration but the RAD-TEAM one): ONE THREAD PER ENVIRONMENT, WARP-AUTONOMOUS.  A warp owns 32 consecutive
// environments and never meets another warp after the tile has landed: no CTA barrier, no work list, no atomics on the
// hot path.  The design follows what the round-1 profile of the tile program (rs_step_tiled.cuh) showed -- the step is
// bound by instruction issue, and a third of its instructions ran at 4-14 active lanes inside data-dependent loops:
//
//   * every per-rectangle loop is unrolled over the template bound KMAX and runs branch-free on all 32 lanes (take_action /
//     in_obstruction, the detector->source segment, the visibility of the hint corner, the corner marking pass);
//   * the shortest path keeps round 1's idea (upper bound through last step's best corner, then only corners that can
//     still improve on it) but the marking pass now applies the EXACT improvement test in directed-rounding fp32
//     (d^2 < (best - dsrc[c])^2 on a float lower bound of the per-episode table, RsState.dsf); about one corner per unit
//     survives it, and the survivors of the whole warp are evaluated as (unit, corner) pairs, one per lane;
//   * the Poisson draw is finished in place: RS_F_FAST_POISSON takes Poisson(bkg) -- 9 units in 10: the line of sight is
//     blocked, R:498-502 -- from an alias table (rs_poisson_alias.h) and the rest from the fp32 PTRS sampler fed by the same
//     Philox block; the numpy-exact sampler tries the squeeze first and calls the full sampler only where it fails;
//   * the other cross-lane phase are the 8-direction ray casts, as (unit, direction) work items of the warp.
//
// Results are identical to the tile program's (and the oracle's): the integer geometry is the same code, the fp64
// values are the same expressions in the same order.  The functions below are plain per-unit code, also compiled as
// host C++ by tests/emu; the kernel in rs_kernels.cu adds the tile movement and the warp-level item distribution.
// R: = /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py, T: = algos/multiagent/train.py
#pragma once
#include "rs_step_tiled.cuh"
#include "rs_poisson_alias.h"

// RS_S1_ROLL: bit set = that per-rectangle loop stays a loop instead of being unrolled over KMAX (1 in_obstruction, 2 source
// segment, 4 visibility, 8 marking pass, 16 sensor candidates).  The bodies are branch-free either way; rolled, the kernel's
// hot code is ~9 KB smaller (46 KB of hot instruction lines against a 32 KB instruction cache) and measures 2-4 % faster.
#ifndef RS_S1_ROLL
#define RS_S1_ROLL 31
#endif

namespace rs {

constexpr int kS1Roll = RS_S1_ROLL;

// x / d for a divisor d that is a launch constant, rd = 1 / d rounded to nearest: q0 = x * rd, then one correction with
// the exact residual.  Equals the IEEE quotient (Markstein: a correctly rounded reciprocal and a faithful first quotient
// give the correctly rounded result; checked against `/` on 10^8 operands in tests/test_kernel_logic_emu.py) for finite x.
__device__ __forceinline__ double div_const(double x, double d, double rd) {
    if (!(fabs(x) < __longlong_as_double(0x7ff0000000000000LL))) return x / d;
    const double q0 = __dmul_rn(x, rd);
    const double r = __fma_rn(-q0, d, x);
    return __fma_rn(r, rd, q0);
}

// Python round(x, 2) R:613 (see round2() in rs_device.cuh), the final division by 100 through div_const
__device__ __forceinline__ double round2_fast(double x) {
    const double p = __dmul_rn(x, 100.0);
    const double e = __fma_rn(x, 100.0, -p);
    double n = rint(p);
    const double diff = p - n;
    if (diff == 0.5 || diff == -0.5) {
        if (e > 0) n = floor(p) + 1.0;
        else if (e < 0) n = floor(p);
    }
    return div_const(n, 100.0, 0.01);
}

// in_obstruction R:1148-1170 over the unit's rectangle column, unrolled and branch-free: the FIRST rectangle (index order)
// whose closed set holds the point decides, so the rectangles are visited last to first and each overrides the later ones
template <int KMAX>
__device__ __forceinline__ bool in_obstruction1(const int4 *rects, int rstride, int num_obs, int px, int py) {
    bool blocked = false;
#pragma unroll((kS1Roll & 1) ? 1 : (KMAX > 0 ? KMAX : 1))
    for (int k = KMAX - 1; k >= 0; k--) {
        const int4 r = rects[k * rstride];
        const bool closed = (k < num_obs) & (r.x <= px) & (px <= r.z) & (r.y <= py) & (py <= r.w);
        const bool open = (r.x < px) & (px < r.z) & (r.y < py) & (py < r.w);
        blocked = closed ? open : blocked;
    }
    return blocked;
}

// the near-corner clause of boundary_distance < 0.001 (see corner_grazes) as a call: it is needed for about one segment
// in a thousand, and inlined it would keep four cross products alive through the whole rectangle loop
__device__ __noinline__ bool graze_call(int px, int py, int qx, int qy, int4 r) {
    int cr[4];
    seg_rect(px, py, qx, qy, r, cr);
    const int dx = qx - px, dy = qy - py;
    return corner_grazes(px, py, dx, dy, dx * dx + dy * dy, r, cr);
}

// source_segment() of rs_env_impl.cuh without the box pre-filter and the per-lane rectangle list: every rectangle, every
// lane, no branch but the (rare) near-corner call; rectangles past num_obs are computed and masked out.
// src_in: the rectangles that hold the source strictly inside (meta bits 9..15, set by rs_load_scenarios; never in a
// sampled scenario): there, a detector strictly inside the same rectangle meets no boundary.
template <int KMAX>
__device__ __forceinline__ void source_segment1(const int4 *rects, int rstride, int num_obs, int src_in, int px, int py, int sx,
                                                int sy, bool &direct, bool &blocked) {
    const Seg1 s = make_seg1(px, py, sx, sy);
    const bool far = s.dx * s.dx + s.dy * s.dy > 1000000;
    bool any_open = false, any_closed = false;
#pragma unroll((kS1Roll & 2) ? 1 : (KMAX > 0 ? KMAX : 1))
    for (int k = 0; k < KMAX; k++) {
        const bool on = k < num_obs;
        const int4 r = rects[k * rstride];
        int cr[4];
        bool open, closed;
        seg_both1(s, r, open, closed, cr);
        if (src_in) closed = closed & !(((src_in >> k) & 1) && in_rect_open(px, py, r));
        // near-corner clause (|cross| <= 3, |pq| > 1000): guarded by one unsigned minimum over the four cross products
        const unsigned g = min(min((unsigned)(cr[0] + 3), (unsigned)(cr[1] + 3)),
                               min((unsigned)(cr[2] + 3), (unsigned)(cr[3] + 3)));
        if (on & far & (g <= 6u) & !closed) closed = graze_call(px, py, sx, sy, r);
        any_open |= on & open;
        any_closed |= on & closed;
    }
    direct = !any_open;
    blocked = any_closed;
}

// visible() of rs_env_impl.cuh, unrolled and branch-free (rectangles past num_obs are computed and masked out)
template <int KMAX>
__device__ __forceinline__ bool visible1(const int4 *rects, int rstride, int num_obs, int px, int py, int qx, int qy) {
    const Seg1 s = make_seg1(px, py, qx, qy);
    bool hit = false;
#pragma unroll((kS1Roll & 4) ? 1 : (KMAX > 0 ? KMAX : 1))
    for (int k = 0; k < KMAX; k++) hit |= (k < num_obs) & seg_open1(s, rects[k * rstride]);
    return !hit;
}

// Marking pass of the pruned shortest path: the corners that may still improve on the upper bound `best`.  A corner c
// improves iff it is tangent (see sp_seed_and_mask) and dsrc[c] + |c - p| < best, i.e. |c - p|^2 < (best - dsrc[c])^2 with
// best - dsrc[c] > 0.  Evaluated in fp32 with every rounding directed to the side that keeps a corner: dsf[c] <= dsrc[c]
// (table rounded down at reset), bf >= best * (1 + 1e-6) (rounded up and inflated: the margin covers the two fp64
// roundings of the exact candidate), the difference and its square rounded up, |c - p|^2 rounded down (exact below 2^24).
// A marked corner is then evaluated exactly, so marking too many is harmless and marking too few impossible.
template <int KMAX>
__device__ __forceinline__ uint32_t mark1(const int4 *rects, int rstride, int num_obs, const float *dsf, int px, int py,
                                          float bf) {
    uint32_t mask = 0u;
#pragma unroll((kS1Roll & 8) ? 1 : (KMAX > 0 ? KMAX : 1))
    for (int k = 0; k < KMAX; k++) {
        const int4 r = rects[k * rstride];
        const float4 d = *reinterpret_cast<const float4 *>(dsf + 4 * k);
        const int ux0 = r.x - px, ux1 = r.z - px, uy0 = r.y - py, uy1 = r.w - py;
        const int qx0 = ux0 * ux0, qx1 = ux1 * ux1, qy0 = uy0 * uy0, qy1 = uy1 * uy1;
        const float t0 = __fsub_ru(bf, d.x), t1 = __fsub_ru(bf, d.y), t2 = __fsub_ru(bf, d.z), t3 = __fsub_ru(bf, d.w);
        // corners p0 (x0,y0), p1 (x0,y1), p2 (x1,y1), p3 (x1,y0); tangent: u.x*u.y <= 0 at p0/p2, >= 0 at p1/p3
        const bool m0 = (ux0 * uy0 <= 0) & (t0 > 0.0f) & (__int2float_rd(qx0 + qy0) < __fmul_ru(t0, t0));
        const bool m1 = (ux0 * uy1 >= 0) & (t1 > 0.0f) & (__int2float_rd(qx0 + qy1) < __fmul_ru(t1, t1));
        const bool m2 = (ux1 * uy1 <= 0) & (t2 > 0.0f) & (__int2float_rd(qx1 + qy1) < __fmul_ru(t2, t2));
        const bool m3 = (ux1 * uy0 >= 0) & (t3 > 0.0f) & (__int2float_rd(qx1 + qy0) < __fmul_ru(t3, t3));
        const uint32_t m4 = (uint32_t)m0 | ((uint32_t)m1 << 1) | ((uint32_t)m2 << 2) | ((uint32_t)m3 << 3);
        mask |= (k < num_obs ? m4 : 0u) << (4 * k);         // rectangles past num_obs: computed and masked out
    }
    return mask;
}

// Pruned shortest path source -> p for a unit whose source segment is obstructed, in two halves.
// (1) sp_hint_mark1: the upper bound through the corner that was optimal at the previous step (`hint`; ds_hint = its
//     dsrc entry, fetched by the caller ahead of time; any hint is allowed, it only seeds the bound), then the marking
//     pass.  best / besti = the value through the hint (inf / -1 when it is unusable); returns the marked corners.
// (2) sp_pair1, once per marked corner: its exact candidate if it beats `best` and is visible from p, else inf.
// The minimum of best and the pair values is exactly shortest_path() / shortest_path_pruned() of rs_env_impl.cuh.
template <int KMAX>
__device__ __forceinline__ uint32_t sp_hint_mark1(const int4 *rects, int rstride, int num_obs, const float *dsf, int px,
                                                  int py, int hint, double ds_hint, double &best, int &besti) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int nc = 4 * num_obs;
    // evaluated on every lane, usable or not: no branch around the root and the visibility test
    const int hc = hint < nc ? hint : 0;
    const int4 r = rects[(hc >> 2) * rstride];
    const int cx = corner_x(r, hc & 3), cy = corner_y(r, hc & 3);
    const bool vis = visible1<KMAX>(rects, rstride, num_obs, px, py, cx, cy);
    const double cand = ds_hint + dist_int(px - cx, py - cy);
    const bool ok = hint < nc && ds_hint < inf && vis;
    best = ok ? cand : inf;
    besti = ok ? hint : -1;
    const float bf = ok ? __fmul_ru(__double2float_ru(cand), 1.000001f) : __int_as_float(0x7f800000);
    uint32_t mask = mark1<KMAX>(rects, rstride, num_obs, dsf, px, py, bf);
    if (ok) mask &= ~(1u << hint);
    return mask;
}
template <int KMAX>
__device__ __forceinline__ double sp_pair1(const int4 *rects, int rstride, int num_obs, int px, int py, int c, double ds,
                                           double best) {
    const int4 r = rects[(c >> 2) * rstride];
    const int cx = corner_x(r, c & 3), cy = corner_y(r, c & 3);
    const double cand = ds + dist_int(px - cx, py - cy);
    const bool vis = visible1<KMAX>(rects, rstride, num_obs, px, py, cx, cy);
    return (cand < best && vis) ? cand : __longlong_as_double(0x7ff0000000000000LL);
}

// RS_F_FAST_POISSON draw from the unit's first Philox block x[0..3] (counter block 0 of the stream poisson_f32 walks):
// integer lambda in the alias table's range -> one look-up; lambda >= 10 -> fp32 PTRS, two proposals from this block,
// further blocks through poisson_f32_from(); anything else -> the generic sampler.
__device__ __noinline__ long long poisson_f32_from(uint64_t seed, uint32_t env_id, uint32_t domain, uint32_t agent,
                                                   uint64_t ctr, double lam, uint32_t first_block) {
    PtrsF32 s;
    s.init(lam);
    const uint32_t c1 = (domain << 24) | (agent << 16);
    for (uint32_t j = first_block; j < 500u; j++) {
        uint32_t x[4];
        philox4x32_10(env_id, c1 + j, (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), x);
        long long k;
        if (s.propose(x[0], x[1], true, k)) return k;
        if (s.propose(x[2], x[3], true, k)) return k;
    }
    return (long long)floor(lam);
}

__device__ __forceinline__ long long count_fast1(const StepArgs &a, int n, uint64_t step_ctr, const uint32_t x[4],
                                                 bool table, int bkg, double lam, uint32_t &status) {
    if (table) {
        const uint32_t e = rs_poisson_alias[(bkg - RS_PA_LO) * RS_PA_NK + (int)(x[0] >> 25)];
        return (x[1] >> 8) < (e >> 7) ? (long long)(x[0] >> 25) : (long long)(e & 127u);
    }
    if (lam >= 10) {
        PtrsF32 s;
        s.init(lam);
        long long k;
        if (s.propose(x[0], x[1], true, k)) return k;
        if (s.propose(x[2], x[3], true, k)) return k;
        return poisson_f32_from(a.seed, a.env_id0 + (uint32_t)n, 0, 0, step_ctr, lam, 1u);
    }
    Rng g;
    g.init_philox(a.seed, a.env_id0 + (uint32_t)n, 0, 0, step_ctr);
    const long long k = poisson<true>(g, lam);
    status |= g.status;
    return k;
}

// the numpy-exact sampler (Philox stream shared with the oracle, or the injected uniforms): squeeze first, then the whole
// sampler from the start of the same stream where the squeeze missed -- poisson<false>()'s value in every case
__device__ __noinline__ long long count_exact_retry(const StepArgs &a, int n, uint64_t step_ctr, double lam,
                                                    uint32_t &status) {
    return unit_count<false>(a, n, 1, 0, step_ctr, lam, status);
}
__device__ __forceinline__ long long count_exact1(const StepArgs &a, int n, uint64_t step_ctr, double lam, uint32_t &status) {
    Rng g;
    unit_rng(g, a, n, 1, 0, step_ctr);
    long long k;
    const bool ok = poisson_first(g, lam, k);
    status |= g.status;
    return ok ? k : count_exact_retry(a, n, step_ctr, lam, status);
}

constexpr int UF1_IDLE = 512;   // unit flag of this kernel (next to UF_* of rs_step_tiled.cuh): the action was 8 (idle)

// What take_action and the segment to the source leave in registers.
struct Move1 {
    int2 det;            // position after take_action
    int af;              // aflags with this step's changes (oob count, blocked bit)
    int uf;              // UF_MOVED | UF_OOB | UF_NEED_D | sensor candidate rectangles << 16
    int d2;              // |det - src|^2
    bool direct;         // source and detector see each other: the shortest path is the segment
    bool blocked_raw;    // boundary_distance(segment, some rectangle) < 0.001 (R:1139-1141 without the isclose clause)
    uint32_t status;
};

// take_action R:876-946 (one agent: no collision), the detector -> source segment, the sensor candidates.  rects: the
// unit's rectangle column in shared memory (element k at rects[k * rstride]).
template <int KMAX>
__device__ __forceinline__ Move1 unit1_move(const Params &P, const int4 *rects, int rstride, int2 src, int meta, int action,
                                            int2 det, int af) {
    Move1 m;
    const int num_obs = meta & 0xff;
    int uf = 0;
    uint32_t status = 0;
    if (action == 8) uf |= UF1_IDLE;                                    // all the commit phase needs to know of the action
    if (action >= 0) {
        const int tx = det.x + step_dx(action), ty = det.y + step_dy(action);
        bool roll = false;
        if (P.enforce) {
            if (tx < P.bx0 || ty < P.by0 || P.bx1 <= tx || P.by1 <= ty) { uf |= UF_OOB; af += 1; roll = true; }
        } else {
            if (det.x < P.sx0 || det.y < P.sy0 || P.sx1 < det.x || P.sy1 < det.y) { uf |= UF_OOB; af += 1; }
        }
        if (in_obstruction1<KMAX>(rects, rstride, num_obs, tx, ty)) { roll = true; af |= 1 << 24; }
        if (!roll) { det.x = tx; det.y = ty; uf |= UF_MOVED; }
    }
    if ((unsigned)(det.x + 16383) > 32766u || (unsigned)(det.y + 16383) > 32766u) status |= RS_ST_COORD_RANGE;
    source_segment1<KMAX>(rects, rstride, num_obs, (meta >> 9) & 0x7f, det.x, det.y, src.x, src.y, m.direct, m.blocked_raw);
    int cand = 0;                                                       // sensor candidates: a ray is at most 100 long
#pragma unroll((kS1Roll & 16) ? 1 : (KMAX > 0 ? KMAX : 1))
    for (int k = 0; k < KMAX; k++) {
        const int4 r = rects[k * rstride];
        const bool near = (k < num_obs) & (r.x - 100 <= det.x) & (det.x <= r.z + 100) & (r.y - 100 <= det.y) & (det.y <= r.w + 100);
        cand |= (int)near << k;
    }
    if (cand) uf |= UF_NEED_D | (cand << 16);
    const int ddx = det.x - src.x, ddy = det.y - src.y;
    m.d2 = ddx * ddx + ddy * ddy;
    m.det = det; m.af = af; m.uf = uf; m.status = status;
    return m;
}

// What the measurement leaves for the sensing and commit halves.
struct Unit1 {
    int2 det;
    int af;              // aflags incl. the hint corner of the next step
    int uf;
    double sp;           // shortest-path length
    bool blocked_los;    // is_intersect R:1133-1146
    float count;         // raw Poisson count
    uint32_t status;
};

// Shortest-path value -> line of sight -> expected counts -> Poisson draw.  sp_blocked / hint: the pruned search's result for
// a unit whose segment is obstructed (ignored for direct units); x = Philox block 0 of the unit (kFast).
template <bool kFast>
__device__ __forceinline__ Unit1 unit1_measure(const Params &P, const StepArgs &a, const Move1 &m, int n, int2 rad,
                                               double sp_blocked, int hint, uint64_t step_ctr, const uint32_t x[4]) {
    Unit1 o;
    uint32_t status = m.status;
    // euc is needed where it is the answer (direct), where it sets the expected count (line of sight free) and for the
    // isclose leftover (euc <= 2, i.e. d2 <= 4); the other units -- most of them -- never take the root
    double euc = 0.0;
    if (m.direct || !m.blocked_raw || m.d2 <= 4) euc = sqrt((double)m.d2);
    const double sp = m.direct ? euc : sp_blocked;
    int af = m.af;
    if (!m.direct) af = (af & ~(31 << 25)) | (hint << 25);
    // is_intersect R:1133-1146 = blocked_raw && !isclose(sqrt(euc), sp, abs_tol=0.1); the isclose clause can only hold for
    // euc <= 2 (see rs_step_tiled.cuh)
    bool blocked_los = m.blocked_raw;
    if (m.d2 <= 4) blocked_los = blocked_los && !isclose_quirk(euc, sp);
    EnvView e;
    e.intensity = rad.x; e.bkg = rad.y;
    long long k;
    if (kFast && !a.uniforms) {
        const bool table = blocked_los && rad.y >= RS_PA_LO && rad.y <= RS_PA_HI;
        double lam = 0.0;
        if (!table) lam = unit_lambda(P, e, euc, blocked_los, status);
        k = count_fast1(a, n, step_ctr, x, table, rad.y, lam, status);
    } else {
        const double lam = unit_lambda(P, e, euc, blocked_los, status);
        k = count_exact1(a, n, step_ctr, lam, status);
    }
    o.det = m.det; o.af = af; o.uf = m.uf; o.sp = sp; o.blocked_los = blocked_los; o.count = (float)k; o.status = status;
    return o;
}

// One direction d of obstruction_sensors (R:1186-1217) for a detector at (px,py): the candidate rectangles in index
// order, their edges in the order (p0,p1) left, (p0,p3) bottom, (p2,p1) top, (p2,p3) right (R:1000-1006), at most two
// scored hits per direction.  Returns the squared distance to the nearest scored edge (-1: none); hits += scored edges of
// rectangle k << 8k (obs_idx_ls R:1190).  d is a run-time value here (one lane per direction): no branch depends on it.
__device__ __forceinline__ int sense_dir1(const int4 *rects, int rstride, int cand, int px, int py, int d,
                                          unsigned long long &hits) {
    const int sx = step_dx(d), sy = step_dy(d);
    const int ax = sx < 0 ? -sx : sx, ay = sy < 0 ? -sy : sy;
    const int gx = sx > 0 ? 1 : -1, gy = sy > 0 ? 1 : -1;               // only used when the component is non-zero
    const int ryx = (sx == 0 || sy == 0) ? 0 : gx * gy;                 // dy/dx along the ray (+-1 on the diagonals)
    const int xlo = min(px, px + sx), xhi = max(px, px + sx), ylo = min(py, py + sy), yhi = max(py, py + sy);
    int inter = 0, dmin = -1;
    int todo = cand;
    while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int4 r = rects[k * rstride];
        int hk = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            bool hit;
            int d2;
            if (s == 0 || s == 3) {                                     // vertical edge x = c, y in [r.y, r.w]
                const int c = (s == 0) ? r.x : r.z;
                const int t = c - px;
                const int yat = py + t * ryx;
                const bool cross = (unsigned)(t * gx) <= (unsigned)ax && r.y <= yat && yat <= r.w;
                const bool along = px == c && ylo <= r.w && r.y <= yhi;
                hit = sx == 0 ? along : cross;
                const int ex = px - c, ey = clampdist(py, r.y, r.w);
                d2 = ex * ex + ey * ey;
            } else {                                                    // horizontal edge y = c, x in [r.x, r.z]
                const int c = (s == 1) ? r.y : r.w;
                const int t = c - py;
                const int xat = px + t * ryx;
                const bool cross = (unsigned)(t * gy) <= (unsigned)ay && r.x <= xat && xat <= r.z;
                const bool along = py == c && xlo <= r.z && r.x <= xhi;
                hit = sy == 0 ? along : cross;
                const int ey = py - c, ex = clampdist(px, r.x, r.z);
                d2 = ex * ex + ey * ey;
            }
            if (inter < 2 && hit) {
                dmin = (dmin < 0 || d2 < dmin) ? d2 : dmin;
                inter++;
                hk++;
            }
        }
        hits += (unsigned long long)hk << (8 * k);
    }
    return dmin;
}

// can the ray of direction d from (px,py) touch any candidate rectangle at all?  (closed boxes: a superset of sense_dir1's hits)
__device__ __forceinline__ bool sense_box1(const int4 *rects, int rstride, int cand, int px, int py, int d) {
    const int sx = step_dx(d), sy = step_dy(d);
    const int xlo = min(px, px + sx), xhi = max(px, px + sx), ylo = min(py, py + sy), yhi = max(py, py + sy);
    bool any = false;
    int todo = cand;
    while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int4 r = rects[k * rstride];
        any = any || (r.x <= xhi && xlo <= r.z && r.y <= yhi && ylo <= r.w);
    }
    return any;
}

// (110 - dist)/110 of a scored edge; 0 = no hit, exactly 1 on the edge (MUFU.RSQ, ~3e-7 relative, as sensors_rects_row)
__device__ __forceinline__ float sense_value(int dmin) {
    const float f2 = (float)max(dmin, 1);
    const float v = (110.0f - f2 * rsqrtf(f2)) * (1.0f / 110.0f);
    return dmin < 0 ? 0.0f : (dmin == 0 ? 1.0f : v);
}

// max(zip(obs_idx_ls, self.poly)) R:1222-1226: most hits, ties -> lexicographically largest vertex list
__device__ __forceinline__ int sense_correct_rect(const int4 *rects, int rstride, int num_obs, unsigned long long hits) {
    int hits_best = -1, best_k = 0;
    for (int k = 0; k < num_obs; k++) {
        const int hk = (int)((hits >> (8 * k)) & 0xffull);
        bool take = hk > hits_best;
        if (!take && hk == hits_best) {
            const int4 p = rects[k * rstride], q = rects[best_k * rstride];
            take = (p.x != q.x) ? (p.x > q.x) : ((p.y != q.y) ? (p.y > q.y) : ((p.w != q.w) ? (p.w > q.w) : (p.z > q.z)));
        }
        if (take) { hits_best = hk; best_k = k; }
    }
    return best_k;
}

// What a unit writes back.
struct Commit1 {
    float reward;
    int done, info, ended, meta;
    double best;
    bool scheduled;
};

// reward / terminal / caller rules for one unit (phase_commit of rs_step_tiled.cuh for one agent).  row = the unit's
// observation row (shared memory) holding the sensors; count / coordinates / standardisation / walls are finished here.
__device__ __forceinline__ Commit1 unit1_commit(const Params &P, const StepArgs &a, const Unit1 &o, int meta, int action,
                                                double best, float *row, double *st_mean, double *st_m2, float *raw,
                                                uint32_t &status) {
    Commit1 c;
    int done = (meta >> 8) & 1;
    int ep_len = meta >> 16;
    const bool have_act = a.actions != nullptr;
    float cnt = o.count;
    if (st_mean) {                                                      // T:436 update(next_obs[0]), T:339/469 standardize
        if (raw) *raw = cnt;
        cnt = (float)stat_standardize(P.standardize, ep_len + (have_act ? 2 : 1), (double)cnt, *st_mean, *st_m2, have_act);
    }
    row[0] = cnt;
    row[1] = (float)((double)o.det.x * P.inv_scale);
    row[2] = (float)((double)o.det.y * P.inv_scale);
    if (P.enforce) sensors_walls(P, o.det.x, o.det.y, row + 3, status);  // R:1232-1259
    int info = (o.uf & UF_OOB ? RS_I_OOB : 0) | (o.blocked_los ? RS_I_LOS_BLOCKED : 0);
    const double sp = o.sp;
    double reward;
    if (o.uf & UF_MOVED) {                                              // R:507-522
        info |= RS_I_MOVED;
        if (sp < 110) { reward = 0.1; done = 1; }
        else if (sp < best) { reward = 0.1; best = sp; }
        else if (o.uf & UF1_IDLE) reward = div_const(-1.0 * sp, P.max_dist, P.inv_max_dist);
        else reward = div_const(-0.5 * sp, P.max_dist, P.inv_max_dist);
    } else {
        reward = div_const(-0.5 * sp, P.max_dist, P.inv_max_dist);      // R:549, 567
    }
    reward = round2_fast(reward);                                       // R:613
    if (o.af & (1 << 24)) info |= RS_I_BLOCKED;
    int ended = done ? RS_E_TERMINAL : 0;
    if (have_act) ep_len += 1;
    bool scheduled = false;
    if (a.flags & RS_F_AUTO_RESET) {                                    // T:394-405, 446-548
        const bool timeout = ep_len == P.max_ep_len;
        if (timeout) ended |= RS_E_TIMEOUT;
        if (done || timeout || (a.flags & RS_F_EPOCH_END)) { ended |= RS_E_RESET; scheduled = true; }
    }
    c.reward = (float)reward; c.done = done; c.info = info; c.ended = ended; c.best = best; c.scheduled = scheduled;
    c.meta = (meta & 0xfeff) | (done << 8) | (ep_len << 16);
    return c;
}

}  // namespace rs
