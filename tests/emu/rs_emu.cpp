// TEST INFRASTRUCTURE ONLY -- host build of the step / reset kernel bodies (one loop iteration per CUDA thread) with
// the same C entry-point shapes as include/radsearch_b200.h, operating on host memory.  Used by
// tests/test_kernel_logic_emu.py to compare the kernel logic with the oracle where no GPU exists.
#define RS_HOST_EMU 1
#include "../../radiation_ppo_b200/csrc/rs_step1.cuh"
#include "../../radiation_ppo_b200/csrc/rs_rcp.cuh"

#include <vector>

// The single-agent kernel's per-unit code (rs_step1.cuh) for every environment in turn; the warp-level distribution of
// the (unit, direction) sensor items is replayed as a plain loop over the eight directions.
template <bool kFast, int KMAX>
static void step1_env(const rs::Params &P, const RsState *st, const rs::StepArgs &a, int n, uint64_t step_ctr) {
    const int K = P.k_max, N = a.n_env;
    int4 rects[RS_MAX_K];
    for (int k = 0; k < K; k++) rects[k] = reinterpret_cast<const int4 *>(st->rects)[(size_t)k * N + n];
    alignas(16) float dsf[4 * RS_MAX_K];
    for (int c = 0; c < 4 * K; c++) dsf[c] = st->dsf[(size_t)n * 4 * K + c];
    const int2 src = reinterpret_cast<const int2 *>(st->src)[n], rad = reinterpret_cast<const int2 *>(st->rad)[n];
    const int2 det = reinterpret_cast<const int2 *>(st->det)[n];
    const int meta = st->meta[n], af = st->aflags[n];
    const int action = a.actions ? a.actions[n] : -1;
    double best = st->best[n], stm = 0.0, stq = 0.0;
    if (P.standardize) { stm = st->st_mean[n]; stq = st->st_m2[n]; }
    uint32_t x[4] = {0, 0, 0, 0};
    if (kFast)
        rs::philox4x32_10(a.env_id0 + (uint32_t)n, 0u, (uint32_t)step_ctr, (uint32_t)(step_ctr >> 32), (uint32_t)a.seed,
                          (uint32_t)(a.seed >> 32), x);
    float row[RS_OBS_DIM] = {0};
    const int num_obs = meta & 0xff, hint = (af >> 25) & 31;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double *drow = st->dsrc + (size_t)n * 4 * K;
    const rs::Move1 mv = rs::unit1_move<KMAX>(P, rects, 1, src, meta, action, det, af);
    double best_sp = 0.0;
    int besti = -1;
    if (!mv.direct) {
        const double ds_hint = hint < 4 * num_obs ? drow[hint] : inf;
        uint32_t marked = rs::sp_hint_mark1<KMAX>(rects, 1, num_obs, dsf, mv.det.x, mv.det.y, hint, ds_hint, best_sp, besti);
        const double best0 = best_sp;                  // every pair is judged against the hint's bound, as on the GPU
        while (marked) {
            const int c = __ffs(marked) - 1;
            marked &= marked - 1;
            const double v = rs::sp_pair1<KMAX>(rects, 1, num_obs, mv.det.x, mv.det.y, c, drow[c], best0);
            if (v < best_sp) { best_sp = v; besti = c; }
        }
    }
    const rs::Unit1 o = rs::unit1_measure<kFast>(P, a, mv, n, rad, best_sp, besti >= 0 ? besti : hint, step_ctr, x);
    uint32_t status = o.status;
    if (o.uf & rs::UF_NEED_D) {
        unsigned long long hits = 0ull;
        int dmin[8], ones = 0;
        for (int d = 0; d < 8; d++) {
            dmin[d] = rs::sense_dir1(rects, 1, (o.uf >> 16) & 0xff, o.det.x, o.det.y, d, hits);
            ones += dmin[d] == 0;
            row[3 + d] = rs::sense_value(dmin[d]);
        }
        if (ones > 3) {
            float out[8];
            rs::correct_coords(o.det.x, o.det.y, rects[rs::sense_correct_rect(rects, 1, meta & 0xff, hits)], out, status);
            for (int d = 0; d < 8; d++) row[3 + d] = out[d];
        }
    }
    float raw = 0.0f;
    const rs::Commit1 c = rs::unit1_commit(P, a, o, meta, action, best, row, P.standardize ? &stm : nullptr, &stq, &raw, status);
    st->meta[n] = c.meta;
    reinterpret_cast<int2 *>(st->det)[n] = o.det;
    st->best[n] = c.best;
    st->aflags[n] = o.af;
    if (P.standardize) {
        st->st_mean[n] = stm; st->st_m2[n] = stq;
        if (st->raw_count) st->raw_count[n] = raw;
    }
    for (int i = 0; i < RS_OBS_DIM; i++) a.obs[(size_t)n * RS_OBS_DIM + i] = row[i];
    if (a.reward) a.reward[n] = c.reward;
    if (a.team_reward) a.team_reward[n] = c.reward;
    if (a.done) a.done[n] = (uint8_t)c.done;
    if (a.info) a.info[n] = (uint8_t)c.info;
    if (a.ended) a.ended[n] = (uint8_t)c.ended;
    if (c.scheduled) {
        if (a.final_obs) for (int i = 0; i < RS_OBS_DIM; i++) a.final_obs[(size_t)n * RS_OBS_DIM + i] = row[i];
        st->reset_list[(*st->reset_count)++] = n;
    }
    if (status) st->status[n] |= status;
}

extern "C" {

int emu_step(const RsConfig *cfg, const RsState *st, const int32_t *actions, float *obs, float *reward,
             float *team_reward, uint8_t *done, uint8_t *info, uint8_t *ended, float *final_obs, int32_t n_env,
             uint32_t env_id0, uint64_t seed, uint64_t step_ctr, const double *uniforms, int32_t n_uniforms,
             int32_t flags) {
    rs::Params P = rs::make_params(*cfg);
    rs::StepArgs a;
    a.actions = actions; a.obs = obs; a.reward = reward; a.team_reward = team_reward; a.final_obs = final_obs;
    a.done = done; a.info = info; a.ended = ended; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed;
    a.step_ctr = step_ctr; a.uniforms = uniforms; a.n_uniforms = n_uniforms; a.flags = flags;
    a.parity = (flags & RS_F_PARITY1) ? 1 : 0;
    if (flags & RS_F_AUTO_RESET) *st->reset_count = 0;
    // one-environment tiles (E = 1): copy the state rows in, run the phases in the kernel's order, copy the rows out
    const int A = cfg->n_agents, K = cfg->k_max, N = n_env;
    const rs::TileLayout L = rs::make_layout(1, A, K, 1, cfg->standardize);
    std::vector<unsigned char> buf(L.total + 16);
    unsigned char *base = buf.data() + ((16 - (reinterpret_cast<uintptr_t>(buf.data()) & 15)) & 15);
    rs::Tile T = rs::carve_tile(base, L, 1, A, K, actions != nullptr);
    const bool fast = (flags & RS_F_FAST_POISSON) && !uniforms;
    for (int n = 0; n < n_env; n++) {
        T.src[0] = reinterpret_cast<const int2 *>(st->src)[n];
        T.rad[0] = reinterpret_cast<const int2 *>(st->rad)[n];
        T.meta[0] = st->meta[n];
        for (int k = 0; k < K; k++) T.rects[k] = reinterpret_cast<const int4 *>(st->rects)[(size_t)k * N + n];
        for (int ag = 0; ag < A; ag++) {
            T.det[ag] = reinterpret_cast<const int2 *>(st->det)[(size_t)ag * N + n];
            T.best[ag] = st->best[(size_t)ag * N + n];
            T.af[ag] = st->aflags[(size_t)ag * N + n];
            if (T.stm) { T.stm[ag] = st->st_mean[(size_t)ag * N + n]; T.stq[ag] = st->st_m2[(size_t)ag * N + n]; }
            if (actions) const_cast<int *>(T.act)[ag] = actions[(size_t)n * A + ag];
        }
        for (int u = 0; u < A; u++) {
            if (fast) rs::phase_move<true>(P, *st, a, T, n, u, step_ctr);
            else rs::phase_move<false>(P, *st, a, T, n, u, step_ctr);
        }
        for (int u = 0; u < A; u++) {
            const int uf = T.uflag[u];
            if (uf & rs::UF_NEED_B) rs::phase_path(*st, T, n, u);
            if (uf & rs::UF_NEED_D) rs::phase_sense(*st, T, n, u);
            if (uf & rs::UF_NEED_P) {
                if (fast) rs::phase_count<true>(P, *st, a, T, n, u, step_ctr);
                else rs::phase_count<false>(P, *st, a, T, n, u, step_ctr);
            }
        }
        const bool sched = fast ? rs::phase_commit<true>(P, *st, a, T, n, 0, step_ctr)
                                : rs::phase_commit<false>(P, *st, a, T, n, 0, step_ctr);
        if (sched) st->reset_list[(*st->reset_count)++] = n;
        st->meta[n] = T.meta[0];
        for (int ag = 0; ag < A; ag++) {
            reinterpret_cast<int2 *>(st->det)[(size_t)ag * N + n] = T.det[ag];
            st->best[(size_t)ag * N + n] = T.best[ag];
            st->aflags[(size_t)ag * N + n] = T.af[ag];
            if (T.stm) {
                st->st_mean[(size_t)ag * N + n] = T.stm[ag];
                st->st_m2[(size_t)ag * N + n] = T.stq[ag];
                if (st->raw_count) st->raw_count[(size_t)n * A + ag] = T.raw[ag];
            }
            for (int i = 0; i < RS_OBS_DIM; i++) obs[((size_t)n * A + ag) * RS_OBS_DIM + i] = T.obs[ag * RS_OBS_DIM + i];
            if (reward) reward[(size_t)n * A + ag] = T.reward[ag];
            if (done) done[(size_t)n * A + ag] = T.done[ag];
            if (info) info[(size_t)n * A + ag] = T.info[ag];
        }
        if (team_reward) team_reward[n] = T.team[0];
        if (ended) ended[n] = T.ended[0];
    }
    return 0;
}

int emu_step1(const RsConfig *cfg, const RsState *st, const int32_t *actions, float *obs, float *reward,
              float *team_reward, uint8_t *done, uint8_t *info, uint8_t *ended, float *final_obs, int32_t n_env,
              uint32_t env_id0, uint64_t seed, uint64_t step_ctr, const double *uniforms, int32_t n_uniforms,
              int32_t flags) {
    if (cfg->n_agents != 1) return -1;
    rs::Params P = rs::make_params(*cfg);
    rs::StepArgs a;
    a.actions = actions; a.obs = obs; a.reward = reward; a.team_reward = team_reward; a.final_obs = final_obs;
    a.done = done; a.info = info; a.ended = ended; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed;
    a.step_ctr = step_ctr; a.uniforms = uniforms; a.n_uniforms = n_uniforms; a.flags = flags;
    a.parity = (flags & RS_F_PARITY1) ? 1 : 0;
    if (flags & RS_F_AUTO_RESET) *st->reset_count = 0;
    const bool fast = (flags & RS_F_FAST_POISSON) && !uniforms;
    const int K = cfg->k_max;
    for (int n = 0; n < n_env; n++) {
        // the same unroll bounds the GPU launch picks
        if (K == 0) { if (fast) step1_env<true, 0>(P, st, a, n, step_ctr); else step1_env<false, 0>(P, st, a, n, step_ctr); }
        else if (K <= 3) { if (fast) step1_env<true, 3>(P, st, a, n, step_ctr); else step1_env<false, 3>(P, st, a, n, step_ctr); }
        else if (K <= 5) { if (fast) step1_env<true, 5>(P, st, a, n, step_ctr); else step1_env<false, 5>(P, st, a, n, step_ctr); }
        else { if (fast) step1_env<true, 8>(P, st, a, n, step_ctr); else step1_env<false, 8>(P, st, a, n, step_ctr); }
    }
    return 0;
}

// x / d through the reciprocal (rs_step1.cuh::div_const) against the IEEE quotient: returns the number of mismatches
long long emu_div_const_mismatches(const double *x, long long n, double d) {
    const double rd = 1.0 / d;
    long long bad = 0;
    for (long long i = 0; i < n; i++) {
        const double q = rs::div_const(x[i], d, rd), w = x[i] / d;
        bad += !(q == w || (q != q && w != w));
    }
    return bad;
}
// x / k through the reciprocal table (rs_rcp.cuh::div_count) against the IEEE quotient: number of mismatches over x[] x k[]
long long emu_div_count_mismatches(const double *x, long long n, const int *k, long long nk) {
    long long bad = 0;
    for (long long j = 0; j < nk; j++)
        for (long long i = 0; i < n; i++) {
            const double q = div_count(x[i], k[j]), w = x[i] / (double)k[j];
            bad += !(q == w || (q != q && w != w));
        }
    return bad;
}
// the prepared-segment predicates of rs_device.cuh (seg_open1 | seg_both1's closed bit << 1 | disagreement with seg_cross_open1 << 2)
int emu_seg_prepared(int px, int py, int qx, int qy, int x0, int y0, int x1, int y1) {
    const rs::Seg1 s = rs::make_seg1(px, py, qx, qy);
    const int4 r = make_int4(x0, y0, x1, y1);
    int cr[4];
    bool open, closed;
    rs::seg_both1(s, r, open, closed, cr);
    const bool o1 = rs::seg_open1(s, r);
    const bool box = r.x < s.xhi && s.xlo < r.z && r.y < s.yhi && s.ylo < r.w;
    const bool o2 = box && rs::seg_cross_open1(s, r);
    return (int)o1 | ((int)closed << 1) | ((int)(o1 != open || o1 != o2) << 2);
}
double emu_round2_fast(double x) { return rs::round2_fast(x); }
double emu_round2(double x) { return rs::round2(x); }

// bit0: open segment meets the open rectangle, bit1: closed segment meets the closed rectangle (rs_device.cuh::seg_rect)
int emu_seg_rect(int px, int py, int qx, int qy, int x0, int y0, int x1, int y1) {
    return rs::seg_rect(px, py, qx, qy, make_int4(x0, y0, x1, y1));
}
int emu_los_blocked_rect(int px, int py, int qx, int qy, int x0, int y0, int x1, int y1) {
    return rs::los_blocked_rect(px, py, qx, qy, make_int4(x0, y0, x1, y1)) ? 1 : 0;
}

int emu_query_shortest_path(const RsConfig *cfg, const RsState *st, const int32_t *pts, double *out, int32_t n_env,
                            int32_t variant) {
    std::vector<int4> rects(RS_MAX_K);
    for (int n = 0; n < n_env; n++)
        out[n] = rs::query_sp(*st, n, n_env, cfg->k_max, pts[2 * n], pts[2 * n + 1], variant, rs::Col<int4>{rects.data(), 1});
    return 0;
}

static int run_reset(const RsConfig *cfg, const RsState *st, const rs::ResetArgs &a, const uint8_t *mask,
                     const uint8_t *new_mask, int flags, const int32_t *list = nullptr, const int32_t *cnt = nullptr) {
    rs::Params P = rs::make_params(*cfg);
    std::vector<int4> rects(RS_MAX_K);
    std::vector<double> dsrc(4 * RS_MAX_K);
    std::vector<uint32_t> vis(4 * RS_MAX_K);
    if ((flags & RS_F_RESET_LIST) && !list) { list = st->reset_list; cnt = st->reset_count; }
    const int count = list ? *cnt : a.n_env;
    const bool fast = (flags & RS_F_FAST_POISSON) && !a.uniforms;
    for (int i = 0; i < count; i++) {
        int n = i;
        if (list) n = list[i];
        else if (mask && !mask[n]) continue;
        const bool new_obs = !a.prepare && ((flags & RS_F_NEW_OBSTACLES) || (new_mask && new_mask[n]));
        // one "lane" plays the whole warp: the lane-strided loops degenerate to plain loops
        if (fast) rs::reset_env<true>(P, *st, a, n, new_obs, 0, 1, 1u, rs::Col<int4>{rects.data(), 1}, rs::Col<double>{dsrc.data(), 1}, rs::Col<uint32_t>{vis.data(), 1});
        else rs::reset_env<false>(P, *st, a, n, new_obs, 0, 1, 1u, rs::Col<int4>{rects.data(), 1}, rs::Col<double>{dsrc.data(), 1}, rs::Col<uint32_t>{vis.data(), 1});
    }
    return 0;
}

int emu_reset(const RsConfig *cfg, const RsState *st, const uint8_t *reset_mask, const uint8_t *new_obstacles_mask,
              float *obs, int32_t n_env, uint32_t env_id0, uint64_t seed, uint64_t step_ctr, const double *uniforms,
              int32_t n_uniforms, int32_t flags) {
    rs::ResetArgs a;
    memset(&a, 0, sizeof(a));
    a.obs = obs; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed; a.step_ctr = step_ctr;
    a.uniforms = uniforms; a.n_uniforms = n_uniforms;
    a.parity = (flags & RS_F_PREFETCH) ? ((flags & RS_F_PARITY1) ? 1 : 0) : -1;
    return run_reset(cfg, st, a, reset_mask, new_obstacles_mask, flags);
}

int emu_prepare(const RsConfig *cfg, const RsState *st, int32_t n_env, uint32_t env_id0, uint64_t seed, int32_t flags) {
    rs::ResetArgs a;
    memset(&a, 0, sizeof(a));
    a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed; a.prepare = 1; a.parity = -1;
    const int parity = (flags & RS_F_PARITY1) ? 1 : 0;
    const bool use_list = flags & RS_F_REFILL_LIST;
    return run_reset(cfg, st, a, nullptr, nullptr, flags & RS_F_FAST_POISSON,
                     use_list ? st->refill_list + (size_t)parity * n_env : nullptr,
                     use_list ? st->refill_count + parity : nullptr);
}

int emu_load_scenarios(const RsConfig *cfg, const RsState *st, const int32_t *src, const int32_t *det,
                       const int32_t *intensity, const int32_t *bkg, const int32_t *rects, int32_t k_in,
                       const int32_t *num_obs, float *obs, int32_t n_env, uint32_t env_id0, uint64_t seed,
                       uint64_t step_ctr, const double *uniforms, int32_t n_uniforms) {
    rs::ResetArgs a;
    memset(&a, 0, sizeof(a));
    a.obs = obs; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed; a.step_ctr = step_ctr;
    a.uniforms = uniforms; a.n_uniforms = n_uniforms;
    a.in_src = src; a.in_det = det; a.in_intensity = intensity; a.in_bkg = bkg; a.in_rects = rects;
    a.in_num_obs = num_obs; a.k_in = k_in;
    a.parity = -1;
    return run_reset(cfg, st, a, nullptr, nullptr, 0);
}
}
