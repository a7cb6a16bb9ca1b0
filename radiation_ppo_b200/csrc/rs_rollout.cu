// Caller-side bookkeeping of the batched rollout (the rules train.py applies around env.step), fused
// into two elementwise launches per step so that a rollout of N environments costs a handful of launches per step instead
// of a few dozen tiny tensor operations:
//   rs_rollout_pre   before the env step: the policy's action / state value / log-probability and the source coordinates of
//                    step t go into row t of the rollout buffer (PPOBuffer.store P:339-381; T:416-428)
//   rs_rollout_post  after the env step: bootstrap value of the trajectories that were cut (T:462-487), restart of the
//                    recurrent state of the envs whose episode ended (T:509-511), EpRet / EpLen / DoneCount / OutOfBound
//                    bookkeeping (T:361-391, 493-527) -- what rollout_stats.EpisodeStats.update does with tensor ops
// T: = /root/reference/algos/multiagent/train.py, P: = /root/reference/algos/multiagent/ppo.py
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/radsearch_b200.h"
#include "rs_error.h"

namespace {

__global__ void __launch_bounds__(256) rollout_pre_kernel(const int32_t *__restrict__ action, const float *__restrict__ val,
                                                          const float *__restrict__ logp, const int32_t *__restrict__ src,
                                                          float *__restrict__ act_row, float *__restrict__ val_row,
                                                          float *__restrict__ logp_row, float *__restrict__ src_row,
                                                          const float *__restrict__ obs, float *__restrict__ obs_row, int obs_dim,
                                                          int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (obs_row)        // the observation the policy just saw -> row t (a caller whose env step writes into fixed buffers)
        for (int j = i; j < n * obs_dim; j += gridDim.x * 256) obs_row[j] = obs[j];
    if (i >= n) return;
    act_row[i] = (float)action[i];
    val_row[i] = val[i];
    logp_row[i] = logp[i];
    if (src_row) {
        const int2 s = reinterpret_cast<const int2 *>(src)[i];
        reinterpret_cast<float2 *>(src_row)[i] = make_float2((float)s.x, (float)s.y);
    }
}

__device__ __forceinline__ void atomic_min_f64(double *p, double v) {
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p);
    unsigned long long old = *q;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long seen = atomicCAS(q, old, (unsigned long long)__double_as_longlong(v));
        if (seen == old) break;
        old = seen;
    }
}
__device__ __forceinline__ void atomic_max_f64(double *p, double v) {
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p);
    unsigned long long old = *q;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long seen = atomicCAS(q, old, (unsigned long long)__double_as_longlong(v));
        if (seen == old) break;
        old = seen;
    }
}

__global__ void __launch_bounds__(256) rollout_post_kernel(const float *__restrict__ reward, const uint8_t *__restrict__ ended,
                                                           const uint8_t *__restrict__ done, const uint8_t *__restrict__ info,
                                                           const float *__restrict__ v_next, float *__restrict__ boot_row,
                                                           float *__restrict__ hidden, int hidden_dim, double *__restrict__ ep_return,
                                                           int32_t *__restrict__ ep_steps, double *acc, double *ep_min,
                                                           double *ep_max, float *__restrict__ rew_row, uint8_t *__restrict__ end_row,
                                                           int n, int last_step) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    double a[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};           // episodes, sum EpRet, sum EpRet^2, sum EpLen, DoneCount, OutOfBound
    if (i < n) {
        const int e = ended[i];
        if (rew_row) rew_row[i] = reward[i];                 // (a caller whose env step writes into fixed buffers)
        if (end_row) end_row[i] = (uint8_t)e;
        // T:462-487: a trajectory cut by the timeout, or by the epoch's last step, is bootstrapped with V(next observation)
        const bool cut = last_step ? true : (e & RS_E_TIMEOUT) != 0;
        boot_row[i] = cut ? v_next[i] : 0.0f;
        if (hidden && !last_step && e != 0)                                  // T:509-511
            for (int h = 0; h < hidden_dim; h++) hidden[(size_t)i * hidden_dim + h] = 0.0f;
        if (ep_return) {
            const double ret = ep_return[i] + (double)reward[i];             // T:361-375
            const int steps = ep_steps[i] + 1;
            a[4] = done[i] != 0;                                             // T:388-391
            a[5] = (info[i] & RS_I_OOB) != 0;                                // T:378-384
            if (e & (RS_E_TERMINAL | RS_E_TIMEOUT)) {                        // episode_over T:394-400, logged T:493-500
                a[0] = 1.0; a[1] = ret; a[2] = ret * ret; a[3] = (double)steps;
                atomic_min_f64(ep_min, ret);
                atomic_max_f64(ep_max, ret);
            }
            const bool reset = (e & RS_E_RESET) != 0;                        // env.reset() follows T:530-535
            ep_return[i] = reset ? 0.0 : ret;
            ep_steps[i] = reset ? 0 : steps;
        }
    }
    if (acc) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
            double v = a[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(acc + k, v);
        }
    }
}

}  // namespace

extern "C" {

int rs_rollout_pre(const int32_t *action, const float *val, const float *logp, const int32_t *src, float *act_row,
                   float *val_row, float *logp_row, float *src_row, const float *obs, float *obs_row, int32_t obs_dim,
                   int32_t n, void *stream) {
    if (!action || !val || !logp || !act_row || !val_row || !logp_row) return rs_set_error("rs_rollout_pre: NULL buffer");
    if (src_row && !src) return rs_set_error("rs_rollout_pre: src_row without src");
    if (obs_row && (!obs || obs_dim <= 0)) return rs_set_error("rs_rollout_pre: obs_row without obs");
    if (n <= 0) return rs_set_error("rs_rollout_pre: n must be positive");
    rollout_pre_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(action, val, logp, src, act_row, val_row,
                                                                                     logp_row, src_row, obs, obs_row, obs_dim, n);
    return (int)cudaGetLastError();
}

int rs_rollout_post(const float *reward, const uint8_t *ended, const uint8_t *done, const uint8_t *info, const float *v_next,
                    float *boot_row, float *hidden, int32_t hidden_dim, double *ep_return, int32_t *ep_steps, double *acc,
                    double *ep_min, double *ep_max, float *rew_row, uint8_t *end_row, int32_t n, int32_t last_step,
                    void *stream) {
    if (!ended || !v_next || !boot_row) return rs_set_error("rs_rollout_post: NULL buffer");
    if (rew_row && !reward) return rs_set_error("rs_rollout_post: rew_row without reward");
    if (ep_return && (!reward || !done || !info || !ep_steps || !acc || !ep_min || !ep_max))
        return rs_set_error("rs_rollout_post: episode statistics need reward / done / info / ep_steps / acc / ep_min / ep_max");
    if (n <= 0 || (hidden && hidden_dim <= 0)) return rs_set_error("rs_rollout_post: bad sizes");
    rollout_post_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reward, ended, done, info, v_next, boot_row, hidden, hidden_dim, ep_return, ep_steps, ep_return ? acc : nullptr, ep_min,
        ep_max, rew_row, end_row, n, last_step);
    return (int)cudaGetLastError();
}

}  // extern "C"
