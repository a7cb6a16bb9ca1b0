// RadSearch.step (R:443-728) + caller rules (T:394-405) as a TILE program: a CTA owns E consecutive environments whose
// structure-of-arrays state rows are staged in shared memory exactly as they lie in HBM (bulk-async copies in, bulk-async
// copies out), and the per-(env, agent) work is cut into phases so that the expensive, data-dependent parts run
// COMPACTED -- only the units that need them, packed into whole warps -- instead of every warp paying for every branch:
//
//   phase_move   every unit   take_action, in_obstruction, detector->source segment (direct? blocked?), first Poisson
//                             proposal (accepted by the PTRS squeeze for ~85 % of the units), sensor candidate mask
//   phase_path   list B       shortest path around the rectangles for units whose source segment is obstructed
//   phase_sense  list D       8-direction ray casts for units within 100 of an obstruction
//   phase_count  list P       the full numpy PTRS sampler for units whose first proposal missed the squeeze (the
//                             counter-based stream restarts, so the result equals the one-piece sampler's)
//   phase_commit every env    walls, reward, terminal, team reward, caller rules, reset work list, state + outputs
//
// The functions here are plain per-unit code (also compiled as host C++ by tests/emu); the CUDA kernel in rs_kernels.cu
// adds the tile movement and the CTA-level list compaction.
// R: = /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py, T: = algos/multiagent/train.py
#pragma once
#include "rs_env_impl.cuh"

namespace rs {

enum : int {
    UF_MOVED = 1, UF_DIRECT = 2, UF_BLOCKED_RAW = 4, UF_NEED_B = 8, UF_NEED_D = 16, UF_NEED_P = 32, UF_RARE = 64,
    UF_COLLISION = 128, UF_OOB = 256,
};

// Shared-memory image of E environments x A agents.  unit u = agent * E + env slot: the per-agent rows are indexed by u.
struct Tile {
    int E, A, K, U;
    // state rows (global layout of the tile)
    int4 *rects;          // [K][E]
    int2 *src, *rad;      // [E]
    int *meta;            // [E]
    int2 *det;            // [A][E]
    double *best;         // [A][E]
    int *af;              // [A][E]
    double *stm, *stq;    // [A][E]   running mean / squared distance of the count channel (nullptr: standardize off)
    const int *act;       // [E][A]   (nullptr: step(None) probe)
    // output rows (global layout of the tile)
    float *obs;           // [E][A][11]
    float *reward;        // [E][A]
    float *team;          // [E]
    uint8_t *done, *info; // [E][A]
    uint8_t *ended;       // [E]
    float *raw;           // [E][A]   unstandardised counts (standardize on)
    // per-unit scratch
    int2 *ndet;           // [U] position after take_action (committed by phase_commit: other agents' proposals are
                          //     computed from the old positions, R:645-648)
    double *sp;           // [U] shortest-path length
    unsigned long long *hkey;   // [U] (candidate bits & ~31) | corner of the best improving pair (~0: none): next step's hint
    int *uflag;           // [U] UF_* | sensor candidate rectangles << 16
};

struct TileLayout {       // byte offsets of the arrays above inside the CTA's dynamic shared memory (all 16-aligned)
    int rects, src, rad, meta, det, best, af, act, obs, reward, team, done, info, ended, ndet, sp, hkey, uflag, lists,
        counters, mbar, stm, stq, raw, total;
};

__host__ __device__ inline int align16(int v) { return (v + 15) & ~15; }

__host__ __device__ inline TileLayout make_layout(int E, int A, int K, int threads, int standardize = 0) {
    TileLayout L;
    const int U = E * A;
    int o = 0;
    L.rects = o;   o += align16(K * E * 16);
    L.src = o;     o += align16(E * 8);
    L.rad = o;     o += align16(E * 8);
    L.meta = o;    o += align16(E * 4);
    L.det = o;     o += align16(U * 8);
    L.best = o;    o += align16(U * 8);
    L.af = o;      o += align16(U * 4);
    L.act = o;     o += align16(U * 4);
    L.obs = o;     o += align16(U * RS_OBS_DIM * 4);
    L.reward = o;  o += align16(U * 4);
    L.team = o;    o += align16(E * 4);
    L.done = o;    o += align16(U);
    L.info = o;    o += align16(U);
    L.ended = o;   o += align16(E);
    L.ndet = o;    o += align16(U * 8);
    L.sp = o;      o += align16(U * 8);
    L.hkey = o;    o += align16(U * 8);
    L.uflag = o;   o += align16(U * 4);
    L.lists = o;   o += align16(3 * U * 2);
    L.counters = o; o += 32;
    L.mbar = o;    o += 16;
    L.stm = L.stq = L.raw = -1;                   // the default configuration keeps its footprint (8 CTAs per SM)
    if (standardize) {
        L.stm = o; o += align16(U * 8);
        L.stq = o; o += align16(U * 8);
        L.raw = o; o += align16(U * 4);
    }
    L.total = o;
    return L;
}

__host__ __device__ inline Tile carve_tile(unsigned char *base, const TileLayout &L, int E, int A, int K, bool have_act) {
    Tile T;
    T.E = E; T.A = A; T.K = K; T.U = E * A;
    T.rects = reinterpret_cast<int4 *>(base + L.rects);
    T.src = reinterpret_cast<int2 *>(base + L.src);
    T.rad = reinterpret_cast<int2 *>(base + L.rad);
    T.meta = reinterpret_cast<int *>(base + L.meta);
    T.det = reinterpret_cast<int2 *>(base + L.det);
    T.best = reinterpret_cast<double *>(base + L.best);
    T.af = reinterpret_cast<int *>(base + L.af);
    T.act = have_act ? reinterpret_cast<const int *>(base + L.act) : nullptr;
    T.obs = reinterpret_cast<float *>(base + L.obs);
    T.reward = reinterpret_cast<float *>(base + L.reward);
    T.team = reinterpret_cast<float *>(base + L.team);
    T.done = base + L.done;
    T.info = base + L.info;
    T.ended = base + L.ended;
    T.ndet = reinterpret_cast<int2 *>(base + L.ndet);
    T.sp = reinterpret_cast<double *>(base + L.sp);
    T.hkey = reinterpret_cast<unsigned long long *>(base + L.hkey);
    T.uflag = reinterpret_cast<int *>(base + L.uflag);
    T.stm = L.stm >= 0 ? reinterpret_cast<double *>(base + L.stm) : nullptr;
    T.stq = L.stq >= 0 ? reinterpret_cast<double *>(base + L.stq) : nullptr;
    T.raw = L.raw >= 0 ? reinterpret_cast<float *>(base + L.raw) : nullptr;
    return T;
}

__device__ __forceinline__ EnvView tile_env(const Tile &T, const RsState &S, int t, int n) {
    EnvView e;
    e.rects = Col<int4>{T.rects + t, T.E};
    e.dsrc = Col<double>{S.dsrc + (size_t)n * 4 * T.K, 1};      // env-major table row, read on demand (phase_path only)
    e.num_obs = T.meta[t] & 0xff;
    const int2 s = T.src[t], r = T.rad[t];
    e.sx = s.x; e.sy = s.y; e.intensity = r.x; e.bkg = r.y;
    return e;
}

__device__ __forceinline__ void unit_rng(Rng &g, const StepArgs &a, int n, int A, int ag, uint64_t step_ctr) {
    if (a.uniforms) g.init_inject(a.uniforms + ((size_t)n * A + ag) * a.n_uniforms, a.n_uniforms);
    else g.init_philox(a.seed, a.env_id0 + (uint32_t)n, 0, (uint32_t)ag, step_ctr);
}

// expected counts R:498-502 (see observe())
__device__ __forceinline__ double unit_lambda(const Params &P, const EnvView &e, double euc, bool blocked_los,
                                              uint32_t &status) {
    if (blocked_los) return (double)e.bkg;
    double d = euc;
    if (d == 0.0) { status |= RS_ST_LAMBDA_INF; d = 1.0; }
    return (P.count_law == 1) ? (double)e.intensity / (d * d) + (double)e.bkg : (double)e.intensity / d + (double)e.bkg;
}

// The first proposal of numpy's PTRS loop, accepted only by its squeeze test (`us >= 0.07 && V <= vr`): exactly what the
// first iteration of poisson<>() computes, so "accept here, else run poisson<>() from the start of the same stream"
// returns poisson<>()'s value in every case.
__device__ __forceinline__ bool poisson_first(Rng &g, double lam, long long &k_out) {
    if (!(lam >= 10)) return false;
    const double slam = sqrt(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 / (b - 2);
    const double U = g.next_double() - 0.5;
    const double V = g.next_double();
    const double us = 0.5 - fabs(U);
    k_out = (long long)floor((2 * a / us + b) * U + lam + 0.43);
    return us >= 0.07 && V <= vr;
}

__device__ __forceinline__ void raise_status(const RsState &S, int n, uint32_t status) {
    if (!status) return;
#ifdef RS_HOST_EMU
    S.status[n] |= status;
#else
    atomicOr(S.status + n, status);
#endif
}

// the whole sampler for one unit (numpy's PTRS / multiplication method, or the single-precision PTRS when kFast)
template <bool kFast>
__device__ __forceinline__ long long unit_count(const StepArgs &a, int n, int A, int ag, uint64_t step_ctr, double lam,
                                                uint32_t &status) {
    if (kFast && lam >= 10) return poisson_f32(a.seed, a.env_id0 + (uint32_t)n, 0, (uint32_t)ag, step_ctr, lam);
    Rng g;
    unit_rng(g, a, n, A, ag, step_ctr);
    const long long k = poisson<kFast>(g, lam);
    status |= g.status;
    return k;
}

// ---- phase_move: every unit -----------------------------------------------------------------------------------------
template <bool kFast>
__device__ __forceinline__ int phase_move(const Params &P, const RsState &S, const StepArgs &a, const Tile &T, int n0,
                                          int u, uint64_t step_ctr) {
    const int E = T.E, A = T.A;
    const int ag = u / E, t = u - ag * E, n = n0 + t;
    const EnvView e = tile_env(T, S, t, n);
    int2 det = T.det[u];
    int af = T.af[u];
    const int action = T.act ? T.act[t * A + ag] : -1;
    int uf = 0;
    uint32_t status = 0;
    if (action >= 0) {                                                  // take_action R:876-946
        const int tx = det.x + step_dx(action), ty = det.y + step_dy(action);
        int cnt = 0;
        if (A > 1)
            for (int i = 0; i < A; i++) {                               // proposals of every agent from the OLD positions
                const int2 d = T.det[i * E + t];
                const int ai = T.act[t * A + i];
                cnt += (d.x + step_dx(ai) == tx && d.y + step_dy(ai) == ty);
            }
        if (cnt > 1) uf |= UF_COLLISION;
        else {
            bool roll = false;
            if (P.enforce) {
                if (tx < P.bx0 || ty < P.by0 || P.bx1 <= tx || P.by1 <= ty) { uf |= UF_OOB; af += 1; roll = true; }
            } else {
                if (det.x < P.sx0 || det.y < P.sy0 || P.sx1 < det.x || P.sy1 < det.y) { uf |= UF_OOB; af += 1; }
            }
            if (in_obstruction(e, tx, ty)) { roll = true; af |= 1 << 24; }
            if (!roll) { det.x = tx; det.y = ty; uf |= UF_MOVED; }
        }
    }
    if ((unsigned)(det.x + 16383) > 32766u || (unsigned)(det.y + 16383) > 32766u) status |= RS_ST_COORD_RANGE;
    // the reference keeps stale sp/euc when the detector did not move; recomputing them at the unchanged position gives
    // the same numbers (R:528-567)
    const double euc = dist_int(det.x - e.sx, det.y - e.sy);
    bool direct, blocked_raw;
    source_segment(e, det.x, det.y, direct, blocked_raw);
    if (direct) uf |= UF_DIRECT;
    else {
        uf |= UF_NEED_B;
#ifndef RS_HOST_EMU
        {                       // the env-major source-distance row (4K doubles) is wanted by phase_path: start it now
            const char *row = reinterpret_cast<const char *>(S.dsrc + (size_t)n * 4 * T.K);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
            if (T.K > 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
        }
#endif
    }
    if (blocked_raw) uf |= UF_BLOCKED_RAW;
    // sensor candidates: a ray is at most 100 long (71 per axis)
    int cand = 0;
    for (int k = 0; k < e.num_obs; k++) {
        const int4 r = e.rects[k];
        if (r.x - 100 <= det.x && det.x <= r.z + 100 && r.y - 100 <= det.y && det.y <= r.w + 100) cand |= 1 << k;
    }
    if (cand) uf |= UF_NEED_D | (cand << 16);
    float *row = T.obs + (t * A + ag) * RS_OBS_DIM;
    // is_intersect R:1133-1146 = blocked_raw && !isclose(sqrt(euc), sp, abs_tol=0.1); the isclose clause can only hold
    // for euc <= 2 (sp >= euc up to rounding, and euc - sqrt(euc) > 0.1 beyond 1.2): then wait for sp (phase_commit)
    if (euc > 2.0) {
        const double lam = unit_lambda(P, e, euc, blocked_raw, status);
        long long k;
        if (kFast && lam >= 10) {                                        // single-precision sampler (Philox only)
            PtrsF32 s;
            s.init(lam);
            uint32_t x[4];
            philox4x32_10(a.env_id0 + (uint32_t)n, (uint32_t)ag << 16, (uint32_t)step_ctr, (uint32_t)(step_ctr >> 32),
                          (uint32_t)a.seed, (uint32_t)(a.seed >> 32), x);
            if (s.propose(x[0], x[1], false, k) == 1) row[0] = (float)k;
            else uf |= UF_NEED_P;
        } else {
            Rng g;
            unit_rng(g, a, n, A, ag, step_ctr);
            if (poisson_first(g, lam, k)) row[0] = (float)k;
            else uf |= UF_NEED_P;
            status |= g.status;
        }
    } else {
        uf |= UF_RARE;
    }
    row[1] = (float)((double)det.x * P.inv_scale);
    row[2] = (float)((double)det.y * P.inv_scale);
#pragma unroll
    for (int d = 0; d < 8; d++) row[3 + d] = 0.0f;
    T.ndet[u] = det;
    T.af[u] = af;
    T.sp[u] = euc;                                                      // final for direct units
    T.uflag[u] = uf;
    raise_status(S, n, status);
    return uf;
}

// ---- phase_path: units whose source segment is obstructed (list B) ----------------------------------------------------
__device__ __forceinline__ void phase_path(const RsState &S, const Tile &T, int n0, int u) {
    const int E = T.E;
    const int ag = u / E, t = u - ag * E;
    const EnvView e = tile_env(T, S, t, n0 + t);
    const int2 det = T.ndet[u];
    int af = T.af[u];
    int hint = (af >> 25) & 31;
    T.sp[u] = shortest_path_pruned(e, e.dsrc.p, det.x, det.y, hint);
    T.hkey[u] = ~0ull;
    T.af[u] = (af & ~(31 << 25)) | (hint << 25);
}

// The same phase cut in two for the CUDA kernel (rs_kernels.cu): the seed + marking pass per unit, then the marked
// (unit, corner) pairs of the whole tile as independent work items -- a unit has ~2 marked corners on average but a
// warp's slowest lane ~9, so walking them per thread leaves 4 lanes in 5 idle.
__device__ __forceinline__ uint32_t phase_path_seed(const RsState &S, const Tile &T, int n0, int u, double &best,
                                                    int &besti) {
    const int E = T.E;
    const int ag = u / E, t = u - ag * E;
    const EnvView e = tile_env(T, S, t, n0 + t);
    const int2 det = T.ndet[u];
    return sp_seed_and_mask(e, e.dsrc.p, det.x, det.y, (T.af[u] >> 25) & 31, best, besti);
}
__device__ __forceinline__ void phase_path_finish(const Tile &T, int u, double best, int besti) {
    T.sp[u] = best;
    T.hkey[u] = ~0ull;
    if (besti >= 0) T.af[u] = (T.af[u] & ~(31 << 25)) | (besti << 25);
}
// the rest of the per-thread walk (units whose pairs did not fit into the tile's pair list)
__device__ __forceinline__ void phase_path_walk(const RsState &S, const Tile &T, int n0, int u, uint32_t mask,
                                                double best, int besti) {
    const int E = T.E;
    const int ag = u / E, t = u - ag * E;
    const EnvView e = tile_env(T, S, t, n0 + t);
    const int2 det = T.ndet[u];
    while (mask) {
        const int c = __ffs(mask) - 1;
        mask &= mask - 1;
        const double cand = sp_corner_candidate(e, e.dsrc.p, det.x, det.y, c, best);
        if (cand < best) { best = cand; besti = c; }
    }
    phase_path_finish(T, u, best, besti);
}
// one (unit, corner) pair against the unit's current minimum `cur`: the candidate, or inf
__device__ __forceinline__ double phase_path_pair(const RsState &S, const Tile &T, int n0, int u, int c, double cur) {
    const int E = T.E;
    const int ag = u / E, t = u - ag * E;
    const EnvView e = tile_env(T, S, t, n0 + t);
    const int2 det = T.ndet[u];
    return sp_corner_candidate(e, e.dsrc.p, det.x, det.y, c, cur);
}
// the exact candidate of a pair again (hint update of the winners)
__device__ __forceinline__ double phase_path_pair_value(const RsState &S, const Tile &T, int n0, int u, int c) {
    const int E = T.E;
    const int ag = u / E, t = u - ag * E;
    const int4 r = T.rects[(c >> 2) * E + t];
    const int2 det = T.ndet[u];
    return S.dsrc[(size_t)(n0 + t) * 4 * T.K + c] + dist_int(det.x - corner_x(r, c & 3), det.y - corner_y(r, c & 3));
}

// ---- phase_sense: units within reach of an obstruction (list D) -------------------------------------------------------
__device__ __forceinline__ void phase_sense(const RsState &S, const Tile &T, int n0, int u) {
    const int E = T.E, A = T.A;
    const int ag = u / E, t = u - ag * E, n = n0 + t;
    const EnvView e = tile_env(T, S, t, n);
    const int2 det = T.ndet[u];
    uint32_t status = 0;
    sensors_rects_row(e, det.x, det.y, (T.uflag[u] >> 16) & 0xff, T.obs + (t * A + ag) * RS_OBS_DIM + 3, status);
    raise_status(S, n, status);
}

// ---- phase_count: units whose first Poisson proposal missed the squeeze (list P) ----------------------------------------
template <bool kFast>
__device__ __forceinline__ void phase_count(const Params &P, const RsState &S, const StepArgs &a, const Tile &T, int n0,
                                            int u, uint64_t step_ctr) {
    const int E = T.E, A = T.A;
    const int ag = u / E, t = u - ag * E, n = n0 + t;
    const EnvView e = tile_env(T, S, t, n);
    const int2 det = T.ndet[u];
    uint32_t status = 0;
    const double euc = dist_int(det.x - e.sx, det.y - e.sy);
    const double lam = unit_lambda(P, e, euc, (T.uflag[u] & UF_BLOCKED_RAW) != 0, status);
    T.obs[(t * A + ag) * RS_OBS_DIM] = (float)unit_count<kFast>(a, n, A, ag, step_ctr, lam, status);
    raise_status(S, n, status);
}

// ---- phase_commit: every environment (its agents in order) ------------------------------------------------------------
// Returns true when the env was scheduled for reset.
template <bool kFast>
__device__ __forceinline__ bool phase_commit(const Params &P, const RsState &S, const StepArgs &a, const Tile &T, int n0,
                                             int t, uint64_t step_ctr) {
    const int E = T.E, A = T.A, n = n0 + t;
    const EnvView e = tile_env(T, S, t, n);
    const int meta = T.meta[t];
    int done = (meta >> 8) & 1;
    int ep_len = meta >> 16;
    uint32_t status = 0;
    bool have_max = false;
    double max_reward = 0.0;
    for (int ag = 0; ag < A; ag++) {
        const int u = ag * E + t;
        const int uf = T.uflag[u];
        const int2 det = T.ndet[u];
        const double sp = T.sp[u];
        double best = T.best[u];
        int af = T.af[u];
        if (uf & UF_NEED_B) {                                           // the pair phase's best improving corner: next hint
            const unsigned long long hk = T.hkey[u];
            if (hk != ~0ull) { af = (af & ~(31 << 25)) | ((int)(hk & 31ull) << 25); T.af[u] = af; }
        }
        const int action = T.act ? T.act[t * A + ag] : -1;
        float *row = T.obs + (t * A + ag) * RS_OBS_DIM;
        bool blocked_los = (uf & UF_BLOCKED_RAW) != 0;
        if (uf & UF_RARE) {                                             // detector within 2 of the source
            const double euc = dist_int(det.x - e.sx, det.y - e.sy);
            blocked_los = blocked_los && !isclose_quirk(euc, sp);
            const double lam = unit_lambda(P, e, euc, blocked_los, status);
            row[0] = (float)unit_count<kFast>(a, n, A, ag, step_ctr, lam, status);
        }
        if (T.stm) {                                                    // T:436 update(next_obs[0]), T:339/469 standardize
            const float x = row[0];
            T.raw[t * A + ag] = x;
            // readings of the episode so far: the reset observation + one per step (this one included)
            row[0] = (float)stat_standardize(P.standardize, ep_len + (T.act ? 2 : 1), (double)x, T.stm[u], T.stq[u],
                                             T.act != nullptr);
        }
        if (P.enforce) sensors_walls(P, det.x, det.y, row + 3, status);  // R:1232-1259
        int info = (uf & UF_COLLISION ? RS_I_COLLISION : 0) | (uf & UF_OOB ? RS_I_OOB : 0) |
                   (blocked_los ? RS_I_LOS_BLOCKED : 0);
        double reward;
        if (uf & UF_MOVED) {                                            // R:507-522
            info |= RS_I_MOVED;
            if (sp < 110) { reward = 0.1; done = 1; }
            else if (sp < best) { reward = 0.1; best = sp; }
            else if (action == 8) reward = -1.0 * sp / P.max_dist;
            else reward = -0.5 * sp / P.max_dist;
        } else {
            reward = -0.5 * sp / P.max_dist;                            // R:549, 567
        }
        reward = round2(reward);                                        // R:613
        if (!have_max || max_reward == 0.0) { max_reward = reward; have_max = true; }   // R:661-665
        else if (max_reward < reward) max_reward = reward;
        if (af & (1 << 24)) info |= RS_I_BLOCKED;
        T.det[u] = det;
        T.best[u] = best;
        T.reward[t * A + ag] = (float)reward;
        T.done[t * A + ag] = (uint8_t)done;
        T.info[t * A + ag] = (uint8_t)info;
    }
    T.team[t] = (float)max_reward;
    int ended = done ? RS_E_TERMINAL : 0;
    if (T.act) ep_len += 1;
    bool scheduled = false;
    if (a.flags & RS_F_AUTO_RESET) {                                    // T:394-405, 446-548
        const bool timeout = ep_len == P.max_ep_len;
        if (timeout) ended |= RS_E_TIMEOUT;
        if (done || timeout || (a.flags & RS_F_EPOCH_END)) {
            ended |= RS_E_RESET;
            scheduled = true;
            if (a.final_obs)
                for (int i = 0; i < A * RS_OBS_DIM; i++)
                    a.final_obs[(size_t)n * A * RS_OBS_DIM + i] = T.obs[t * A * RS_OBS_DIM + i];
        }
    }
    T.ended[t] = (uint8_t)ended;
    T.meta[t] = (meta & 0xfeff) | (done << 8) | (ep_len << 16);      // bits 9..15: see rs_step1.cuh::source_segment1
    raise_status(S, n, status);
    return scheduled;
}

}  // namespace rs
