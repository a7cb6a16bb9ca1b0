// sm_100a kernels + C-ABI entry points for the RadSearch env step / reset (include/radsearch_b200.h).
// Build: radiation_ppo_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rs_env_impl.cuh"
#include "rs_error.h"

namespace {

constexpr int kBlock = 128;

thread_local char g_err[256] = "";

int fail(const char *msg) { return rs_set_error(msg); }

// smem per CTA: rects [K][128] int4 | dsrc [4K][128] f64 | lb [4K][128] f32 | obs tile [128][A*11] f32
template <bool kFast>
__global__ void __launch_bounds__(kBlock, 4) step_kernel(rs::Params P, RsState S, rs::StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    int4 *srects = reinterpret_cast<int4 *>(smem);
    double *sdsrc = reinterpret_cast<double *>(srects + (size_t)P.k_max * kBlock);
    float *slb = reinterpret_cast<float *>(sdsrc + (size_t)4 * P.k_max * kBlock);
    float *sobs = slb + (size_t)4 * P.k_max * kBlock;
    const int row = P.n_agents * RS_OBS_DIM;
    const int n0 = blockIdx.x * kBlock;
    const int n = n0 + threadIdx.x;
    if (n < a.n_env)
        rs::step_env<kFast>(P, S, a, n, rs::Col<int4>{srects + threadIdx.x, kBlock},
                            rs::Col<double>{sdsrc + threadIdx.x, kBlock}, rs::Col<float>{slb + threadIdx.x, kBlock},
                            sobs + threadIdx.x * row);
    __syncthreads();
    // the CTA's observation rows are contiguous in a.obs: coalesced 16-byte stores
    const int cnt = min(kBlock, a.n_env - n0) * row;
    float *dst = a.obs + (size_t)n0 * row;                 // n0 * row * 4 bytes is a multiple of 16
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kBlock)
        reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(sobs)[i];
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += kBlock) dst[i] = sobs[i];
}

// end of a captured step: advance the device step counter and empty the reset list for the next replay
__global__ void bump_ctr_kernel(unsigned long long *ctr, int *reset_count) {
    *ctr += 1ull;
    if (reset_count) *reset_count = 0;
}

__global__ void __launch_bounds__(kBlock) sp_query_kernel(rs::Params P, RsState S, const int32_t *pts, double *out,
                                                           int n_env, int variant) {
    extern __shared__ __align__(16) unsigned char smem[];
    int4 *srects = reinterpret_cast<int4 *>(smem);
    double *sdsrc = reinterpret_cast<double *>(srects + (size_t)P.k_max * kBlock);
    float *slb = reinterpret_cast<float *>(sdsrc + (size_t)4 * P.k_max * kBlock);
    const int n = blockIdx.x * kBlock + threadIdx.x;
    if (n >= n_env) return;
    out[n] = rs::query_sp(S, n, n_env, P.k_max, pts[2 * n], pts[2 * n + 1], variant, rs::Col<int4>{srects + threadIdx.x, kBlock},
                          rs::Col<double>{sdsrc + threadIdx.x, kBlock}, rs::Col<float>{slb + threadIdx.x, kBlock});
}

// Reset: a persistent grid whose threads team up in groups of `nl` lanes per environment.  Few envs to reset (the
// steady state: ~N/120 per step) -> a whole warp per env for low latency; a bulk reset (epoch end, synchronised
// timeouts) -> one thread per env, which wastes no lanes on the sequential rejection sampling.
template <bool kFast>
__global__ void __launch_bounds__(kBlock) reset_kernel(rs::Params P, RsState S, rs::ResetArgs a, const uint8_t *mask,
                                                        const uint8_t *new_mask, int flags, const int32_t *list,
                                                        const int32_t *count) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int total = list ? *count : a.n_env;
    // rs_prepare runs off the critical path: one thread per env wastes no lanes; a reset the step is waiting for
    // teams up lanes for latency
    const int nl = a.prepare ? 1 : (total > 32768 ? 1 : (total > 4096 ? 8 : 32));
    const int G = kBlock / nl;                          // groups (environments in flight) per CTA
    const int g = threadIdx.x / nl, lane = threadIdx.x % nl;
    const uint32_t sync_mask = nl == 32 ? 0xffffffffu : (((1u << nl) - 1u) << ((threadIdx.x & 31) & ~(nl - 1)));
    // scratch columns, element i of group g at [i * G + g]: rects [K] int4 | dsrc [4K] f64 | vis [4K] u32
    int4 *srects = reinterpret_cast<int4 *>(smem);
    double *sdsrc = reinterpret_cast<double *>(srects + (size_t)P.k_max * kBlock);
    uint32_t *svis = reinterpret_cast<uint32_t *>(sdsrc + (size_t)4 * P.k_max * kBlock);
    const int stride = gridDim.x * G;
    for (int i = blockIdx.x * G + g; i < total; i += stride) {
        int n = i;
        if (list) n = list[i];
        else if (mask && !mask[n]) continue;
        const bool new_obs = !a.prepare && ((flags & RS_F_NEW_OBSTACLES) || (new_mask && new_mask[n]));
        rs::reset_env<kFast>(P, S, a, n, new_obs, lane, nl, sync_mask, rs::Col<int4>{srects + g, G},
                             rs::Col<double>{sdsrc + g, G}, rs::Col<uint32_t>{svis + g, G});
    }
}

int check_prefetch(const RsConfig *cfg, const RsState *st) {
    if (!st->nx_src || !st->nx_det || !st->nx_rad || !st->nx_best || !st->nx_obs || !st->nx_seq || !st->refill_list ||
        !st->refill_count)
        return fail("prefetch needs the RsState.nx_* / refill_* buffers");
    return 0;
}

int check_cfg(const RsConfig *cfg, const RsState *st, int32_t n_env) {
    if (!cfg || !st) return fail("cfg/state is NULL");
    if (n_env <= 0) return fail("n_env must be positive");
    if (cfg->n_agents < 1 || cfg->n_agents > RS_MAX_A) return fail("n_agents out of range [1, 8]");
    if (cfg->k_max < 0 || cfg->k_max > RS_MAX_K) return fail("k_max out of range [0, 8]");
    if (cfg->obstruction_count < -1 || cfg->obstruction_count > 7) return fail("obstruction_count out of range [-1, 7]");
    if (cfg->obstruction_count > cfg->k_max || (cfg->obstruction_count == -1 && cfg->k_max < 5))
        return fail("k_max smaller than the number of obstructions that can be drawn");
    if (cfg->max_ep_len < 1 || cfg->max_ep_len > 32767) return fail("max_ep_len out of range [1, 32767]");
    if (cfg->bbox[2] - cfg->obs_area[1] <= cfg->bbox[0] + cfg->obs_area[0]) return fail("empty search area");
    if (!st->src || !st->rad || !st->meta || !st->det || !st->best || !st->aflags || !st->status || !st->epi)
        return fail("RsState has NULL members");
    if (cfg->k_max > 0 && (!st->rects || !st->dsrc || !st->vis)) return fail("RsState obstruction tables are NULL");
    return 0;
}

size_t step_smem(const RsConfig *cfg) {
    return (size_t)cfg->k_max * kBlock * (sizeof(int4) + 4 * sizeof(double) + 4 * sizeof(float)) +
           (size_t)kBlock * cfg->n_agents * RS_OBS_DIM * sizeof(float);
}
size_t reset_smem(const RsConfig *cfg) {
    return (size_t)kBlock * cfg->k_max * (sizeof(int4) + 4 * sizeof(double) + 4 * sizeof(uint32_t));
}
constexpr int kResetGrid = 148 * 8;      // persistent: 8 CTAs of 4 warps per SM

}  // namespace

int rs_set_error(const char *msg) {
    std::snprintf(g_err, sizeof(g_err), "%s", msg);
    return -1;
}

extern "C" {

int rs_step(const RsConfig *cfg, const RsState *st, const int32_t *actions, float *obs, float *reward,
            float *team_reward, uint8_t *done, uint8_t *info, uint8_t *ended, float *final_obs, int32_t n_env,
            uint32_t env_id0, uint64_t seed, uint64_t step_ctr, const double *uniforms, int32_t n_uniforms,
            int32_t flags, void *stream) {
    if (int rc = check_cfg(cfg, st, n_env)) return rc;
    if (!obs) return fail("obs is NULL");
    if (uniforms && n_uniforms < 2) return fail("n_uniforms must be >= 2 when uniforms are injected");
    if ((flags & RS_F_AUTO_RESET) && (!st->reset_list || !st->reset_count)) return fail("auto-reset needs reset_list/reset_count");
    if ((flags & RS_F_DEVICE_CTR) && !st->ctr_dev) return fail("RS_F_DEVICE_CTR needs RsState.ctr_dev");
    const int parity = (flags & RS_F_PARITY1) ? 1 : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if ((flags & RS_F_AUTO_RESET) && !(flags & RS_F_DEVICE_CTR)) {     // with RS_F_DEVICE_CTR rs_bump_ctr empties the list
        cudaError_t e = cudaMemsetAsync(st->reset_count, 0, sizeof(int32_t), s);
        if (e != cudaSuccess) return (int)e;
    }
    rs::Params P = rs::make_params(*cfg);
    rs::StepArgs a;
    a.actions = actions; a.obs = obs; a.reward = reward; a.team_reward = team_reward; a.final_obs = final_obs;
    a.done = done; a.info = info; a.ended = ended; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed;
    a.step_ctr = step_ctr; a.uniforms = uniforms; a.n_uniforms = n_uniforms; a.flags = flags; a.parity = parity;
    const int grid = (n_env + kBlock - 1) / kBlock;
    const size_t smem = step_smem(cfg);
    const bool fast = (flags & RS_F_FAST_POISSON) && !uniforms;
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    if (fast) step_kernel<true><<<grid, kBlock, smem, s>>>(P, *st, a);
    else step_kernel<false><<<grid, kBlock, smem, s>>>(P, *st, a);
    return (int)cudaGetLastError();
}

static int launch_reset(const RsConfig *cfg, const RsState *st, const rs::ResetArgs &a, const uint8_t *mask,
                        const uint8_t *new_mask, int flags, const int32_t *list, const int32_t *count, cudaStream_t s) {
    rs::Params P = rs::make_params(*cfg);
    int need = (a.n_env + 3) / 4;                       // one warp per env is the widest teaming
    int cap = kResetGrid;
    if (a.prepare) {                                    // one thread per env, kept small: it shares the GPU with rs_step
        need = (a.n_env + kBlock - 1) / kBlock;
        const char *g = getenv("RS_PREPARE_GRID");
        cap = g ? atoi(g) : 148;
    }
    const int grid = need < cap ? need : cap;
    const size_t smem = reset_smem(cfg);
    const bool fast = (flags & RS_F_FAST_POISSON) && !a.uniforms;
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(reset_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(reset_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    if (fast) reset_kernel<true><<<grid, kBlock, smem, s>>>(P, *st, a, mask, new_mask, flags, list, count);
    else reset_kernel<false><<<grid, kBlock, smem, s>>>(P, *st, a, mask, new_mask, flags, list, count);
    return (int)cudaGetLastError();
}

int rs_reset(const RsConfig *cfg, const RsState *st, const uint8_t *reset_mask, const uint8_t *new_obstacles_mask,
             float *obs, int32_t n_env, uint32_t env_id0, uint64_t seed, uint64_t step_ctr, const double *uniforms,
             int32_t n_uniforms, int32_t flags, void *stream) {
    if (int rc = check_cfg(cfg, st, n_env)) return rc;
    if (!obs) return fail("obs is NULL");
    if (uniforms && n_uniforms < 2) return fail("n_uniforms must be >= 2 when uniforms are injected");
    if ((flags & RS_F_RESET_LIST) && (!st->reset_list || !st->reset_count)) return fail("list reset needs reset_list/reset_count");
    rs::ResetArgs a;
    std::memset(&a, 0, sizeof(a));
    a.obs = obs; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed; a.step_ctr = step_ctr;
    a.uniforms = uniforms; a.n_uniforms = n_uniforms;
    a.prepare = 0;
    a.parity = -1;
    if (flags & RS_F_PREFETCH) {
        if (int rc = check_prefetch(cfg, st)) return rc;
        a.parity = (flags & RS_F_PARITY1) ? 1 : 0;
    }
    const bool use_list = flags & RS_F_RESET_LIST;
    return launch_reset(cfg, st, a, reset_mask, new_obstacles_mask, flags, use_list ? st->reset_list : nullptr,
                        use_list ? st->reset_count : nullptr, static_cast<cudaStream_t>(stream));
}

int rs_prepare(const RsConfig *cfg, const RsState *st, int32_t n_env, uint32_t env_id0, uint64_t seed, int32_t flags,
               void *stream) {
    if (int rc = check_cfg(cfg, st, n_env)) return rc;
    if (int rc = check_prefetch(cfg, st)) return rc;
    rs::ResetArgs a;
    std::memset(&a, 0, sizeof(a));
    a.obs = nullptr; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed;
    a.prepare = 1;
    a.parity = -1;
    const int parity = (flags & RS_F_PARITY1) ? 1 : 0;
    const bool use_list = flags & RS_F_REFILL_LIST;
    return launch_reset(cfg, st, a, nullptr, nullptr, flags & RS_F_FAST_POISSON,
                        use_list ? st->refill_list + (size_t)parity * n_env : nullptr,
                        use_list ? st->refill_count + parity : nullptr, static_cast<cudaStream_t>(stream));
}

int rs_bump_ctr(const RsState *st, void *stream) {
    if (!st || !st->ctr_dev) return fail("RsState.ctr_dev is NULL");
    bump_ctr_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long *>(st->ctr_dev),
                                                                   st->reset_count);
    return (int)cudaGetLastError();
}

int rs_load_scenarios(const RsConfig *cfg, const RsState *st, const int32_t *src, const int32_t *det,
                      const int32_t *intensity, const int32_t *bkg, const int32_t *rects, int32_t k_in,
                      const int32_t *num_obs, float *obs, int32_t n_env, uint32_t env_id0, uint64_t seed,
                      uint64_t step_ctr, const double *uniforms, int32_t n_uniforms, void *stream) {
    if (int rc = check_cfg(cfg, st, n_env)) return rc;
    if (!obs || !src || !det || !intensity || !bkg || !num_obs) return fail("scenario arrays must not be NULL");
    if (k_in < 0 || (k_in > 0 && !rects)) return fail("rects is NULL");
    if (k_in > cfg->k_max) return fail("k_in exceeds k_max");
    if (uniforms && n_uniforms < 2) return fail("n_uniforms must be >= 2 when uniforms are injected");
    rs::ResetArgs a;
    std::memset(&a, 0, sizeof(a));
    a.obs = obs; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed; a.step_ctr = step_ctr;
    a.uniforms = uniforms; a.n_uniforms = n_uniforms;
    a.in_src = src; a.in_det = det; a.in_intensity = intensity; a.in_bkg = bkg; a.in_rects = rects;
    a.in_num_obs = num_obs; a.k_in = k_in;
    a.prepare = 0;
    a.parity = -1;
    return launch_reset(cfg, st, a, nullptr, nullptr, 0, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int rs_query_shortest_path(const RsConfig *cfg, const RsState *st, const int32_t *pts, double *out, int32_t n_env,
                           int32_t variant, void *stream) {
    if (int rc = check_cfg(cfg, st, n_env)) return rc;
    if (!pts || !out) return fail("pts/out is NULL");
    rs::Params P = rs::make_params(*cfg);
    const int grid = (n_env + kBlock - 1) / kBlock;
    const size_t smem = step_smem(cfg);
    if (smem > 48 * 1024) cudaFuncSetAttribute(sp_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sp_query_kernel<<<grid, kBlock, smem, static_cast<cudaStream_t>(stream)>>>(P, *st, pts, out, n_env, variant);
    return (int)cudaGetLastError();
}

const char *rs_last_error(void) { return g_err; }
int rs_version(void) { return RS_VERSION; }
int rs_sizeof_config(void) { return (int)sizeof(RsConfig); }
int rs_sizeof_state(void) { return (int)sizeof(RsState); }

}  // extern "C"
