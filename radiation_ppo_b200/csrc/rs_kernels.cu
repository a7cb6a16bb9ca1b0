// sm_100a kernels + C-ABI entry points for the RadSearch env step / reset (include/radsearch_b200.h).
// Build: radiation_ppo_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "rs_step1.cuh"
#include "rs_error.h"

namespace {

constexpr int kBlock = 128;

thread_local char g_err[256] = "";

int fail(const char *msg) { return rs_set_error(msg); }

// ---------------------------------------------------------------------------------------------------------------
// step kernel: tile program (rs_step_tiled.cuh).  A CTA of 128 threads owns E consecutive environments (128 / 64 / 32
// for 1 / 2 / >2 agents).  Warp 0 issues one bulk-async copy (cp.async.bulk, the TMA engine's 1-D form) per state row --
// each row of the structure-of-arrays state is one contiguous run of E elements in HBM -- completing on an mbarrier;
// the phases then run on the shared-memory image, and the modified state rows and all outputs leave as bulk-async
// stores.  A partial last tile (or a caller whose pointers are not 16-byte aligned) takes plain cooperative copies.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(mbar)),
        "r"(parity)
        : "memory");
}

// One row of the tile: `bpe` bytes per environment, contiguous over the environments of the tile both in HBM (at
// g + n0 * bpe) and in shared memory (at smem + off).  The host fills the table once per launch (rs_step), so the kernel's
// staging loops are a table walk instead of per-row address arithmetic.
struct RowEnt {
    unsigned long long g;   // HBM address of the row's element for env 0 of this rank (0 = row absent)
    uint32_t off;           // byte offset inside the CTA's dynamic shared memory
    uint32_t bpe;           // bytes per environment
};
constexpr int kMaxRowsIn = 4 + RS_MAX_K + 5 * RS_MAX_A, kMaxRowsOut = 8 + 5 * RS_MAX_A;
struct TileRows {
    int n_in, n_out;
    RowEnt in[kMaxRowsIn], out[kMaxRowsOut];
};

// warp-aggregated append of unit u to a shared-memory work list (whole warps call this)
__device__ __forceinline__ void list_push(bool flag, int u, uint16_t *list, int *count) {
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (flag) list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)u;
}

// Shortest path, first half, for unit u (-1: this thread has none; whole warps call this): seed + corner marking, then the
// marked corners are appended as (unit, corner) pairs to the tile-wide list (warp-aggregated reservation).  A warp whose
// pairs do not fit walks its corners itself.
__device__ __forceinline__ void seed_and_push(const RsState &S, const rs::Tile &T, int n0, int u, uint16_t *pairs,
                                              int pair_cap, int *pair_count) {
    const int lane = threadIdx.x & 31;
    uint32_t mask = 0u;
    int besti = -1;
    double best = 0.0;
    if (u >= 0) mask = rs::phase_path_seed(S, T, n0, u, best, besti);
    const int n = __popc(mask);
    int pre = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += v;
    }
    const int tot = __shfl_sync(0xffffffffu, pre, 31);
    int wbase = 0;
    if (lane == 31 && tot) wbase = atomicAdd(pair_count, tot);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    const bool fits = wbase + tot <= pair_cap;
    if (!fits)                                                // reserved but unused slots inside the list
        for (int i = wbase + lane; i < min(pair_cap, wbase + tot); i += 32) pairs[i] = 0xffffu;
    if (u >= 0) {
        if (fits) {
            rs::phase_path_finish(T, u, best, besti);
            int pos = wbase + pre - n;
            while (mask) {
                pairs[pos++] = (uint16_t)((u << 5) | (__ffs(mask) - 1));
                mask &= mask - 1;
            }
        } else {
            rs::phase_path_walk(S, T, n0, u, mask, best, besti);
        }
    }
}

template <bool kFast, int E, int kOcc, int TB>
__global__ void __launch_bounds__(TB, kOcc) step_kernel(const __grid_constant__ rs::Params P,
                                                         const __grid_constant__ RsState S,
                                                         const __grid_constant__ rs::StepArgs a,
                                                         const __grid_constant__ rs::TileLayout L,
                                                         const __grid_constant__ TileRows R, int bulk_ok,
                                                         uint32_t tx_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int A = P.n_agents, K = P.k_max, U = E * A;
    const rs::Tile T = rs::carve_tile(smem, L, E, A, K, a.actions != nullptr);
    uint16_t *lists = reinterpret_cast<uint16_t *>(smem + L.lists);     // [3][U]: B, D, P
    int *counters = reinterpret_cast<int *>(smem + L.counters);         // B, D, P, scheduled, pairs
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L.mbar);
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * E;
    const int valid = min(E, a.n_env - n0);
    const bool bulk = bulk_ok && valid == E;

    // ---- stage the tile ------------------------------------------------------------------------------------------
    if (tid < 8) counters[tid] = 0;
    if (bulk) {
        if (tid == 0) {
            mbar_init(mbar, 1);
            mbar_expect_tx(mbar, tx_bytes);
        }
        if (tid < 32) {
            __syncwarp();
            for (int i = tid; i < R.n_in; i += 32) {
                const RowEnt r = R.in[i];
                bulk_g2s(smem + r.off, reinterpret_cast<const void *>(r.g + (unsigned long long)n0 * r.bpe),
                         (uint32_t)E * r.bpe, mbar);
            }
        }
        __syncthreads();                                    // the barrier object is initialised for everybody
        mbar_wait(mbar, 0);
    } else {
        for (int i = 0; i < R.n_in; i++) {
            const RowEnt r = R.in[i];
            const uint32_t *g = reinterpret_cast<const uint32_t *>(r.g + (unsigned long long)n0 * r.bpe);
            uint32_t *s = reinterpret_cast<uint32_t *>(smem + r.off);
            for (uint32_t w = tid; w < (uint32_t)valid * r.bpe / 4u; w += TB) s[w] = g[w];
        }
        __syncthreads();
    }
    const uint64_t step_ctr = (a.flags & RS_F_DEVICE_CTR) ? *S.ctr_dev : a.step_ctr;

    // ---- phase_move: every unit; build the work lists ---------------------------------------------------------------
    for (int u0 = 0; u0 < U; u0 += TB) {
        const int u = u0 + tid;
        int uf = 0;
        if (u < U && (u % E) < valid) uf = rs::phase_move<kFast>(P, S, a, T, n0, u, step_ctr);
        list_push(uf & rs::UF_NEED_B, u, lists, counters + 0);
        list_push(uf & rs::UF_NEED_D, u, lists + U, counters + 1);
        list_push(uf & rs::UF_NEED_P, u, lists + 2 * U, counters + 2);
    }
    __syncthreads();
    // ---- compacted phases: list j starts at the first warp the previous lists left idle -----------------------------------
    {
        const int cb = counters[0], cd = counters[1], cp = counters[2];
        {
            // shortest path, flattened: (1) seed + marking pass per unit, the marked corners appended as (unit, corner)
            // pairs to a tile-wide list (warp-aggregated reservation; the reward / team rows serve as scratch until
            // phase_commit writes them); (2) one pair per thread: candidate, visibility, 64-bit atomicMin on the unit's
            // running minimum (positive doubles order like their bit patterns); (3) the winners record the hint.
            uint16_t *pairs = reinterpret_cast<uint16_t *>(T.reward);
            const int pair_cap = (L.done - L.reward) / 2;
            int *pair_count = counters + 4;
            for (int base = 0; base < cb; base += TB) {
                const int j = base + tid;
                seed_and_push(S, T, n0, j < cb ? (int)lists[j] : -1, pairs, pair_cap, pair_count);
            }
            __syncthreads();
            const int np = min(*pair_count, pair_cap);
            unsigned long long *spbits = reinterpret_cast<unsigned long long *>(T.sp);
            // One pool of work for the rest of the phase, fetched by the warps 32 items at a time: first the Poisson
            // retries (few units, the longest dependent chains), then the sensor units, then the (unit, corner) pairs
            // (many, short: they fill the gaps).  No warp waits at a barrier while another walks a list alone.
            const int cd_ = cd, cp_ = cp;
            const int chP = (cp_ + 31) >> 5, chD = (cd_ + 31) >> 5, chB = (np + 31) >> 5;
            const int lane = tid & 31;
            for (;;) {
                int c = 0;
                if (lane == 0) c = atomicAdd(counters + 5, 1);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= chP + chD + chB) break;
                if (c < chP) {
                    const int j = 32 * c + lane;
                    if (j < cp_) rs::phase_count<kFast>(P, S, a, T, n0, lists[2 * U + j], step_ctr);
                } else if (c < chP + chD) {
                    const int j = 32 * (c - chP) + lane;
                    if (j < cd_) rs::phase_sense(S, T, n0, lists[U + j]);
                } else {
                    const int j = 32 * (c - chP - chD) + lane;
                    const uint32_t pr = j < np ? pairs[j] : 0xffffu;
                    if (pr != 0xffffu) {
                        const int u = (int)(pr >> 5), cc = (int)(pr & 31u);
                        const double cur = *reinterpret_cast<volatile double *>(T.sp + u);
                        const double cand = rs::phase_path_pair(S, T, n0, u, cc, cur);
                        if (cand < cur) {
                            const unsigned long long bits = (unsigned long long)__double_as_longlong(cand);
                            atomicMin(spbits + u, bits);
                            // the corner of the smallest candidate (up to the last 5 mantissa bits: it only seeds the
                            // next search)
                            atomicMin(T.hkey + u, (bits & ~31ull) | (unsigned long long)cc);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    // ---- phase_commit: every environment; CTA-aggregated append to the reset work list ---------------------------------

    for (int t0 = 0; t0 < E; t0 += TB) {
        const int t = t0 + tid;
        bool sched = false;
        if (t < valid) sched = rs::phase_commit<kFast>(P, S, a, T, n0, t, step_ctr);
        // reuse list B's storage for the scheduled env slots
        list_push(sched, t, lists, counters + 3);
    }
    __syncthreads();
    {
        const int cs = counters[3];
        if (cs > 0) {
            __shared__ int s_base;
            if (tid == 0) s_base = atomicAdd(S.reset_count, cs);
            __syncthreads();
            for (int j = tid; j < cs; j += TB) S.reset_list[s_base + j] = n0 + lists[j];
        }
    }
    // ---- write the tile back ----------------------------------------------------------------------------------------
    if (bulk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the copy engine
        __syncthreads();
        if (tid < 32) {
            for (int i = tid; i < R.n_out; i += 32) {
                const RowEnt r = R.out[i];
                bulk_s2g(reinterpret_cast<void *>(r.g + (unsigned long long)n0 * r.bpe), smem + r.off, (uint32_t)E * r.bpe);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be released once it has been read
        }
    } else {
        for (int i = 0; i < R.n_out; i++) {
            const RowEnt r = R.out[i];
            const uint8_t *s = smem + r.off;
            uint8_t *g = reinterpret_cast<uint8_t *>(r.g + (unsigned long long)n0 * r.bpe);
            const uint32_t nb = (uint32_t)valid * r.bpe;
            if ((nb & 3u) == 0 && (reinterpret_cast<uintptr_t>(g) & 3u) == 0) {
                for (uint32_t w = tid; w < nb / 4u; w += TB)
                    reinterpret_cast<uint32_t *>(g)[w] = reinterpret_cast<const uint32_t *>(s)[w];
            } else {
                for (uint32_t w = tid; w < nb; w += TB) g[w] = s[w];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-agent step (rs_step1.cuh): one thread per environment, warp-autonomous.  step1_tile is what a warp does for the
// 32 consecutive environments [nw, nw + 32) once their rectangle rows and their block of the float source-distance
// table are in shared memory and the scalar state rows in registers: the front half (move, segment to the source,
// shortest path, Poisson draw) is straight-line per-lane code, the marked corners and the ray casts are (unit, corner) /
// (unit, direction) items of the warp, the reset work list is appended with one atomic per warp, state and scalar outputs
// leave as coalesced stores and the observation rows through a shared-memory staging block as 16-byte stores.
// (A pipelined caller -- 14 warps per SM walking two tiles each with the next tile's rows in flight by bulk copies, no
// spills -- was measured at 33.5 us against 32.1 us for step1_kernel below: a warp issues one instruction in ~8 cycles
// whatever it waits for, so 28 resident warps that all wait at the start beat 14 that never do.  Removed.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPairCap = 128;                               // (unit, corner) pairs of a warp: ~35 at 5 obstructions
struct WarpTile {
    const int4 *rects;      // rectangle k of lane l at rects[k * RST + l]
    const float *dsf;       // [32][4K] float source-distance rows
    float *obs;             // [32][RS_OBS_DIM] staging block of the observation rows (scratch until the commit phase)
    uint8_t *list;          // [32]
    uint16_t *pairs;        // [kPairCap]
    const int2 *rad;        // [32] intensity, background: first needed by the measurement
    const double *best;     // [32] running minimum of the shortest path: first needed by the commit phase
};

template <bool kFast, int KMAX, int RST>
__device__ __forceinline__ void step1_tile(const rs::Params &P, const RsState &S, const rs::StepArgs &a, const int K,
                                           const int nw, const int lane, const bool live, const int2 src,
                                           const int2 det, const int meta, const int action, const int af,
                                           const double ds_hint,
                                           const uint64_t step_ctr, const uint32_t (&x)[4], const WarpTile t,
                                           const int bulk_ok, uint64_t *dsf_bar) {
    const int n = nw + lane;
    const size_t N = (size_t)a.n_env;
    const int hint = (af >> 25) & 31;
    // An episode that reaches its step limit at this step is scheduled for reset whatever happens: whether its prefetched
    // successor is ready (nx_seq == episode number + 1) is read NOW, while the step waits for its tile anyway, instead of as
    // the first link of a chain of dependent loads in the commit phase.  0: not known, 1: not ready, 2: ready.
    // acquire: the tag is read before the rows it publishes (rs_prepare stores them, fences, then stores the tag)
    int pre_adopt = 0;
    if (live && (a.flags & RS_F_PREFETCH) && (a.flags & RS_F_AUTO_RESET) && !(a.flags & RS_F_EPOCH_END) && a.actions &&
        (meta >> 16) + 1 == P.max_ep_len) {
        uint32_t tag;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(tag) : "l"(S.nx_seq + n) : "memory");
        pre_adopt = tag == S.epi[n] + 1u ? 2 : 1;
    }
    // ---- take_action, segment to the source; for obstructed units the bound through last step's corner + marking pass -----
    const int num_obs = meta & 0xff;
    rs::Move1 mv;
    mv.det = det; mv.af = af; mv.uf = 0; mv.d2 = 0; mv.direct = true; mv.blocked_raw = false; mv.status = 0u;
    double best_sp = 0.0;
    int besti = -1;
    uint32_t marked = 0u;
    if (live) mv = rs::unit1_move<KMAX>(P, t.rects + lane, RST, src, meta, action, det, af);
    if (KMAX > 0 && dsf_bar) mbar_wait(dsf_bar, 0);         // the float table was requested after the rectangles had landed
    if (live) {
        if (!mv.direct)
            marked = rs::sp_hint_mark1<KMAX>(t.rects + lane, RST, num_obs, t.dsf + lane * 4 * K, mv.det.x, mv.det.y, hint, ds_hint,
                                             best_sp, besti);
    }
    // ---- the marked corners of the warp as (unit, corner) pairs, one per lane: exact candidate + visibility ---------------
    if (KMAX > 0 && __any_sync(0xffffffffu, marked != 0u)) {
        const int cnt = __popc(marked);
        int incl = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31), off = incl - cnt;
        uint16_t *pl = t.pairs;
        double *res = reinterpret_cast<double *>(t.obs);      // the warp's staging block, free until commit
        // the unit's corners, highest first, at list positions [off, off + fit); the loops below run as long as the unit with
        // the most corners needs, so their bodies are kept to a few instructions
        const int fit = max(0, min(cnt, kPairCap - off));
        {
            uint32_t m = marked;
            uint16_t *q = pl + off;
            const int tag = lane << 5;
            for (int i = 0; i < fit; i++) {
                const int c = 31 - __clz(m);
                m ^= 1u << c;
                q[i] = (uint16_t)(tag | c);
            }
        }
        __syncwarp();
        const int np = min(total, kPairCap);
        for (int base = 0; base < np; base += 32) {
            const int j = base + lane;
            const bool valid = j < np;
            const int e = valid ? (int)pl[j] : 0, owner = e >> 5, c = e & 31;
            const int px = __shfl_sync(0xffffffffu, mv.det.x, owner), py = __shfl_sync(0xffffffffu, mv.det.y, owner);
            const int nob = __shfl_sync(0xffffffffu, num_obs, owner);
            const double bo = __shfl_sync(0xffffffffu, best_sp, owner);
            double ds = __longlong_as_double(0x7ff0000000000000LL);
            if (valid) ds = S.dsrc[(size_t)(nw + owner) * 4 * K + c];
            const double v = rs::sp_pair1<KMAX>(t.rects + owner, RST, nob, px, py, c, ds, bo);
            if (valid) res[j] = v;
        }
        __syncwarp();
        {   // the unit's own pairs: keep the smallest (its list position; the corner is read back once)
            const double *r = res + off;
            int imin = -1;
            for (int i = 0; i < fit; i++) {
                const double v = r[i];
                if (v < best_sp) { best_sp = v; imin = i; }
            }
            if (imin >= 0) besti = (int)pl[off + imin] & 31;
        }
        if (__any_sync(0xffffffffu, fit < cnt)) {           // pairs beyond the warp's list (never seen so far): by their own thread
            uint32_t m = marked;
            for (int i = 0; i < cnt; i++) {
                const int c = 31 - __clz(m);
                m ^= 1u << c;
                if (i >= fit) {
                    const double v = rs::sp_pair1<KMAX>(t.rects + lane, RST, num_obs, mv.det.x, mv.det.y, c,
                                                        S.dsrc[(size_t)n * 4 * K + c], best_sp);
                    if (v < best_sp) { best_sp = v; besti = c; }
                }
            }
        }
        __syncwarp();
    }
    float *row = t.obs + lane * RS_OBS_DIM;
#pragma unroll
    for (int d = 0; d < 8; d++) row[3 + d] = 0.0f;
    rs::Unit1 o;
    o.det = det; o.af = af; o.uf = 0; o.sp = 0.0; o.blocked_los = false; o.count = 0.0f; o.status = 0u;
    if (live) o = rs::unit1_measure<kFast>(P, a, mv, n, t.rad[lane], best_sp, besti >= 0 ? besti : hint, step_ctr, x);
    __syncwarp();
#ifndef RS_S1_RESYNC
#define RS_S1_RESYNC 1
#endif
    // the CTA's warps enter the back half together: they drift apart over the pair phase, and warps at different places of
    // the 46 KB of hot code miss the instruction cache (measured: 85 % of the no-instruction stalls sat behind this point)
    if (RS_S1_RESYNC && RST > 32) __syncthreads();
    // ---- obstruction_sensors: (unit, direction) items of the warp ---------------------------------------------------------
    const unsigned need = __ballot_sync(0xffffffffu, (o.uf & rs::UF_NEED_D) != 0);
    if (need) {
        uint8_t *wl = t.list;
        if ((need >> lane) & 1u) wl[__popc(need & ((1u << lane) - 1u))] = (uint8_t)lane;
        __syncwarp();
        const int cnt = __popc(need);
        for (int base = 0; base < 8 * cnt; base += 32) {
            const int item = base + lane, j = item >> 3, d = item & 7;
            const bool valid = j < cnt;
            const int owner = valid ? (int)wl[j] : 0;
            const int px = __shfl_sync(0xffffffffu, o.det.x, owner), py = __shfl_sync(0xffffffffu, o.det.y, owner);
            const int ufo = __shfl_sync(0xffffffffu, o.uf, owner), nob = __shfl_sync(0xffffffffu, meta, owner) & 0xff;
            const int4 *col = t.rects + owner;
            unsigned long long hits = 0ull;
            int dmin = -1;
            if (valid) dmin = rs::sense_dir1(col, RST, (ufo >> 16) & 0xff, px, py, d, hits);
            float v = rs::sense_value(dmin);
            // the detector stands on an obstruction edge when more than three of its rays read exactly 1.0: R:1219-1226
            const unsigned zero = __ballot_sync(0xffffffffu, valid && dmin == 0);
            const bool fix = __popc((zero >> (lane & 24)) & 0xffu) > 3;
            if (__any_sync(0xffffffffu, fix)) {
#pragma unroll
                for (int s = 1; s < 8; s <<= 1) hits += __shfl_xor_sync(0xffffffffu, hits, s);
                if (fix) {
                    float out[8];
                    uint32_t st = 0u;
                    rs::correct_coords(px, py, col[rs::sense_correct_rect(col, RST, nob, hits) * RST], out, st);
                    v = out[0];
#pragma unroll
                    for (int i = 1; i < 8; i++) v = d == i ? out[i] : v;
                    if (d == 0) rs::raise_status(S, nw + owner, st);
                }
            }
            if (valid) t.obs[owner * RS_OBS_DIM + 3 + d] = v;
        }
        __syncwarp();
    }
    // ---- commit: reward, terminal, caller rules, state and scalar outputs (coalesced) --------------------------------------
    bool sched = false, adopt = false;
    uint32_t status = 0;
    float raw = 0.0f;
    double stm = 0.0, stq = 0.0;
    rs::Commit1 c;
    if (live) {
        status = o.status;
        if (P.standardize) { stm = S.st_mean[n]; stq = S.st_m2[n]; }    // (L2-prefetched when the step started)
        c = rs::unit1_commit(P, a, o, meta, action, t.best[lane], row, P.standardize ? &stm : nullptr, &stq, &raw, status);
        if (a.reward) a.reward[n] = c.reward;
        if (a.team_reward) a.team_reward[n] = c.reward;                 // one agent: the team reward is its reward R:661-665
        if (a.done) a.done[n] = (uint8_t)c.done;
        if (a.info) a.info[n] = (uint8_t)c.info;
        if (a.ended) a.ended[n] = (uint8_t)c.ended;
        sched = c.scheduled;
        if (sched && a.final_obs) {
#pragma unroll
            for (int i = 0; i < RS_OBS_DIM; i++) a.final_obs[(size_t)n * RS_OBS_DIM + i] = row[i];
        }
        // RS_F_PREFETCH: an env whose next episode rs_prepare has already computed (same seed, env, episode number and
        // obstructions: nx_seq carries the episode number) starts it right here and never reaches the reset kernel; the
        // others go to the reset work list as before
        if (sched && (a.flags & RS_F_PREFETCH) && !(a.flags & RS_F_EPOCH_END)) {
            if (pre_adopt) adopt = pre_adopt == 2;
            else {                                                      // ended before its step limit: the tag is read here
                uint32_t tag;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(tag) : "l"(S.nx_seq + n) : "memory");
                adopt = tag == S.epi[n] + 1u;
            }
        }
    }
    // Everything an adopting env reads is requested before anything is stored (one round trip to memory: the rows were
    // pulled into L2 when the step started), and its slot in the refill list is taken at the same time
    int2 s0 = make_int2(0, 0), r1 = s0, d0 = s0;
    double b0 = 0.0;
    uint32_t epi_n = 0u;
    int slot = 0;
    if (adopt) {
        s0 = reinterpret_cast<const int2 *>(S.nx_src)[n];
        r1 = reinterpret_cast<const int2 *>(S.nx_rad)[n];
        d0 = reinterpret_cast<const int2 *>(S.nx_det)[n];
        b0 = S.nx_best[n];
        epi_n = S.epi[n];
        slot = atomicAdd(S.refill_count + a.parity, 1);
    }
    // The copies of an adopted episode are shared by the warp: one lane per table entry / observation value, so that they
    // cost one round trip to memory instead of a chain of twenty by the one thread that owns the env
    unsigned am = __ballot_sync(0xffffffffu, adopt);
    if (am) {
        __syncwarp();
        while (am) {
            const int owner = __ffs(am) - 1;
            am &= am - 1;
            const int nn = nw + owner;
            const int nc = 4 * __shfl_sync(0xffffffffu, num_obs, owner);
            const size_t r0 = (size_t)nn * 4 * K;
            const bool t1 = lane < nc, t2 = lane < RS_OBS_DIM;
            double ds = 0.0;
            float df = 0.0f, ob = 0.0f;
            if (t1) { ds = S.nx_dsrc[r0 + lane]; df = S.nx_dsf[r0 + lane]; }      // source-distance rows of the new episode
            if (t2) ob = S.nx_obs[(size_t)nn * RS_OBS_DIM + lane];                 // its first observation
            if (t1) { S.dsrc[r0 + lane] = ds; S.dsf[r0 + lane] = df; }
            if (t2) t.obs[owner * RS_OBS_DIM + lane] = ob;                  // -> the env's staged row
        }
        __syncwarp();
    }
    if (live) {
        if (adopt) {
            if (P.standardize) {                                        // first reading of the episode: z = 0
                stm = (double)row[0]; stq = 0.0; raw = row[0]; row[0] = 0.0f;
            }
            reinterpret_cast<int2 *>(S.src)[n] = s0;
            reinterpret_cast<int2 *>(S.rad)[n] = make_int2(r1.x, r1.y & 0xff);
            reinterpret_cast<int2 *>(S.det)[n] = d0;
            S.best[n] = b0;
            S.aflags[n] = (r1.y >> 8) << 25;                            // the new episode's search seed (rs_prepare)
            S.meta[n] = meta & 0xff;                                    // done = 0, ep_len = 0 (a sampled source is in no rectangle)
            S.epi[n] = epi_n + 1u;
            if (slot < a.n_env) S.refill_list[(size_t)a.parity * N + slot] = n;
            else status |= RS_ST_REFILL_OVERFLOW;
            sched = false;
        } else {
            S.meta[n] = c.meta;
            reinterpret_cast<int2 *>(S.det)[n] = o.det;
            S.best[n] = c.best;
            S.aflags[n] = o.af;
        }
        if (P.standardize) {
            S.st_mean[n] = stm;
            S.st_m2[n] = stq;
            if (S.raw_count) S.raw_count[n] = raw;
        }
        rs::raise_status(S, n, status);
    }
    {   // reset work list: one atomic per warp
        const unsigned m = __ballot_sync(0xffffffffu, sched);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(S.reset_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sched) S.reset_list[base + __popc(m & ((1u << lane) - 1u))] = n;
        }
    }
    __syncwarp();
    // ---- observation rows: the warp's 32 rows are one contiguous run of 352 floats in shared memory and in HBM --------------
    {
        const float *sw = t.obs;
        float *g = a.obs + (size_t)nw * RS_OBS_DIM;
        if (bulk_ok && nw + 32 <= a.n_env) {
            for (int i = lane; i < 32 * RS_OBS_DIM / 4; i += 32)
                reinterpret_cast<float4 *>(g)[i] = reinterpret_cast<const float4 *>(sw)[i];
        } else {
            const int rows = min(32, a.n_env - nw);
            for (int i = lane; i < rows * RS_OBS_DIM; i += 32) g[i] = sw[i];
        }
    }
}

// an episode that times out at this step is known when its meta word arrives: its prefetched successor (adopted in the
// commit phase) is pulled into L2 while the step is computed
__device__ __forceinline__ void step1_successor_prefetch(const rs::Params &P, const RsState &S, const rs::StepArgs &a,
                                                         const int K, const int n, const int meta) {
    if ((a.flags & RS_F_PREFETCH) && a.actions && (meta >> 16) + 1 == P.max_ep_len) {
        const char *r = reinterpret_cast<const char *>(S.nx_dsrc + (size_t)n * 4 * K);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 128));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.nx_dsf + (size_t)n * 4 * K));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.nx_obs + (size_t)n * RS_OBS_DIM));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.nx_src + (size_t)n * 2));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.nx_det + (size_t)n * 2));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.nx_rad + (size_t)n * 2));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.nx_best + n));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.nx_seq + n));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.epi + n));
    }
    if (P.standardize) {                                    // read by the commit phase
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.st_mean + n));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.st_m2 + n));
    }
}

// A CTA of TB threads owns TB consecutive environments.  Thread 0 starts bulk-async copies (cp.async.bulk, one mbarrier)
// of the tile's rectangle rows [k][TB] and of its block of the float source-distance table [TB][4K] into shared memory --
// the two tables that lanes read for each other's environments or index dynamically; every other state row goes straight
// from HBM into the owning thread's registers with coalesced loads, and the unit's Philox block is computed while both
// are in flight.  After the one CTA barrier that publishes the mbarrier, warps never meet again.
template <bool kFast, int KMAX, int TB, int kOcc>
__global__ void __launch_bounds__(TB, kOcc) step1_kernel(const __grid_constant__ rs::Params P,
                                                          const __grid_constant__ RsState S,
                                                          const __grid_constant__ rs::StepArgs a, int bulk_ok) {
    constexpr int KS = KMAX > 0 ? KMAX : 1;
    __shared__ __align__(16) int4 s_rects[KS * TB];
    __shared__ __align__(16) float s_dsf[TB * 4 * KS];
    __shared__ __align__(16) float s_obs[TB * RS_OBS_DIM];
    __shared__ uint8_t s_list[TB];
    __shared__ uint16_t s_pairs[(TB / 32) * kPairCap];
    __shared__ __align__(16) int2 s_rad[TB];
    __shared__ __align__(16) double s_best[TB];
    __shared__ __align__(8) uint64_t s_mbar[2];
    const int tid = threadIdx.x, lane = tid & 31, w0 = tid & ~31;
    const int n0 = blockIdx.x * TB;
    const int n = n0 + tid;
    const bool live = n < a.n_env;
    const int K = KMAX > 0 ? P.k_max : 0;
    const size_t N = (size_t)a.n_env;
    const bool bulk = KMAX > 0 && bulk_ok && n0 + TB <= a.n_env;
    // The rectangles are needed first (move, segment to the source); the float source-distance block only by the marking
    // pass.  RS_STEP1_DSF_LATE: its copy is requested when the rectangles have landed, so that a launch whose CTAs all start
    // together asks HBM for a third less before anybody can compute.
#ifndef RS_STEP1_DSF_LATE
#define RS_STEP1_DSF_LATE 1
#endif
    if (bulk && tid == 0) {
        mbar_init(&s_mbar[0], 1);
        mbar_init(&s_mbar[1], 2);                           // two arrivals: the late rows here, the float table below
        mbar_expect_tx(&s_mbar[0], (uint32_t)(K * 16 * TB));
        for (int k = 0; k < K; k++)
            bulk_g2s(s_rects + k * TB, S.rects + ((size_t)k * N + n0) * 4, TB * 16, &s_mbar[0]);
        if (!RS_STEP1_DSF_LATE) {
            mbar_expect_tx(&s_mbar[1], (uint32_t)(16 * TB));
            bulk_g2s(s_rad, S.rad + (size_t)n0 * 2, TB * 8, &s_mbar[1]);
            bulk_g2s(s_best, S.best + n0, TB * 8, &s_mbar[1]);
            mbar_expect_tx(&s_mbar[1], (uint32_t)(K * 16 * TB));
            bulk_g2s(s_dsf, S.dsf + (size_t)n0 * 4 * K, (uint32_t)(TB * 16 * K), &s_mbar[1]);
        }
    }
    // scalar state rows: coalesced, straight into registers
    int2 src = make_int2(0, 0), det = make_int2(0, 0);
    int meta = 0, action = -1, af = 0;
    if (live) {
        src = reinterpret_cast<const int2 *>(S.src)[n];
        det = reinterpret_cast<const int2 *>(S.det)[n];
        meta = S.meta[n];
        af = S.aflags[n];
        if (a.actions) action = a.actions[n];
    }
    if (!bulk) {
        s_rad[tid] = live ? reinterpret_cast<const int2 *>(S.rad)[n] : make_int2(1, 10);
        s_best[tid] = live ? S.best[n] : 0.0;
    }
    if (live) step1_successor_prefetch(P, S, a, K, n, meta);
    // the source distance of the corner that was optimal at the previous step: requested now, used after the segment test
    const int hint = (af >> 25) & 31;
    double ds_hint = __longlong_as_double(0x7ff0000000000000LL);
    if (KMAX > 0 && live && hint < 4 * (meta & 0xff)) ds_hint = S.dsrc[(size_t)n * 4 * K + hint];
    const uint64_t step_ctr = (a.flags & RS_F_DEVICE_CTR) ? *S.ctr_dev : a.step_ctr;
    uint32_t x[4] = {0u, 0u, 0u, 0u};
    if (kFast)      // needs nothing from memory: runs while the tile is in flight
        rs::philox4x32_10(a.env_id0 + (uint32_t)n, 0u, (uint32_t)step_ctr, (uint32_t)(step_ctr >> 32), (uint32_t)a.seed,
                          (uint32_t)(a.seed >> 32), x);
    if (bulk) {
        __syncthreads();                                    // the barrier objects are initialised for everybody
        mbar_wait(&s_mbar[0], 0);
        if (RS_STEP1_DSF_LATE && tid == 0) {
            mbar_expect_tx(&s_mbar[1], (uint32_t)(K * 16 * TB));
            bulk_g2s(s_dsf, S.dsf + (size_t)n0 * 4 * K, (uint32_t)(TB * 16 * K), &s_mbar[1]);
            // rows that are first needed late in the step (intensities, running minimum) wait in shared memory instead of
            // in (spilled) registers; like the float table they are asked for once the rectangles are here
            mbar_expect_tx(&s_mbar[1], (uint32_t)(16 * TB));
            bulk_g2s(s_rad, S.rad + (size_t)n0 * 2, TB * 8, &s_mbar[1]);
            bulk_g2s(s_best, S.best + n0, TB * 8, &s_mbar[1]);
        }
    } else if (KMAX > 0) {
        if (live) {
            for (int k = 0; k < K; k++) s_rects[k * TB + tid] = reinterpret_cast<const int4 *>(S.rects)[(size_t)k * N + n];
            for (int k = 0; k < K; k++)
                reinterpret_cast<float4 *>(s_dsf)[tid * K + k] = reinterpret_cast<const float4 *>(S.dsf)[(size_t)n * K + k];
        }
        __syncwarp();
    }
    WarpTile t;
    t.rects = s_rects + w0; t.dsf = s_dsf + w0 * 4 * K; t.obs = s_obs + w0 * RS_OBS_DIM; t.list = s_list + w0;
    t.pairs = s_pairs + (w0 >> 5) * kPairCap; t.rad = s_rad + w0; t.best = s_best + w0;
    step1_tile<kFast, KMAX, TB>(P, S, a, K, n0 + w0, lane, live, src, det, meta, action, af, ds_hint, step_ctr, x, t, bulk_ok,
                                bulk ? &s_mbar[1] : nullptr);
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
    const int n = blockIdx.x * kBlock + threadIdx.x;
    if (n >= n_env) return;
    out[n] = rs::query_sp(S, n, n_env, P.k_max, pts[2 * n], pts[2 * n + 1], variant, rs::Col<int4>{srects + threadIdx.x, kBlock});
}

// Reset: a persistent grid whose threads team up in groups of `nl` lanes per environment.  Few envs to reset (the
// steady state: ~N/120 per step) -> a whole warp per env for low latency; a bulk reset (epoch end, synchronised
// timeouts) -> one thread per env, which wastes no lanes on the sequential rejection sampling.
// RS_RESET_MINB: CTAs per SM the register allocation of the reset / prepare kernel aims at.  rs_prepare shares the GPU with
// the step kernels of other env batches for its whole (latency-bound) life: what it costs them is registers x time.
#ifndef RS_RESET_MINB
#define RS_RESET_MINB 1
#endif
template <bool kFast, int TB, int MINB>
__global__ void __launch_bounds__(TB, MINB) reset_kernel(rs::Params P, RsState S, rs::ResetArgs a, const uint8_t *mask,
                                                    const uint8_t *new_mask, int flags, const int32_t *list,
                                                    const int32_t *count, int prepare_nl) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int total = list ? *count : a.n_env;
    // rs_prepare runs off the critical path: one thread per env wastes no lanes; a reset the step is waiting for
    // teams up lanes for latency
    const int nl = a.prepare ? prepare_nl : (total > 32768 ? 1 : (total > 4096 ? 8 : 32));
    const int G = TB / nl;                              // groups (environments in flight) per CTA
    const int g = threadIdx.x / nl, lane = threadIdx.x % nl;
    const uint32_t sync_mask = nl == 32 ? 0xffffffffu : (((1u << nl) - 1u) << ((threadIdx.x & 31) & ~(nl - 1)));
    // scratch columns, element i of group g at [i * G + g]: rects [K] int4 | dsrc [4K] f64 | vis [4K] u32
    int4 *srects = reinterpret_cast<int4 *>(smem);
    double *sdsrc = reinterpret_cast<double *>(srects + (size_t)P.k_max * TB);
    uint32_t *svis = reinterpret_cast<uint32_t *>(sdsrc + (size_t)4 * P.k_max * TB);
    const int stride = gridDim.x * G;
    for (int i = blockIdx.x * G + g; i < total; i += stride) {
        int n = i;
        if (list) n = list[i];
        else if (mask && !mask[n]) continue;
        const bool new_obs = !a.prepare && ((flags & RS_F_NEW_OBSTACLES) || (new_mask && new_mask[n]));
        rs::reset_env<kFast>(P, S, a, n, new_obs, lane, nl, sync_mask, rs::Col<int4>{srects + g, G},
                             rs::Col<double>{sdsrc + g, G}, rs::Col<uint32_t>{svis + g, G});
    }
    if (flags & RS_F_BUMP_CTR) {
        // tail of a captured step: the last CTA to finish advances the device step counter and empties the reset list
        // (every CTA has read *count by then) -- what rs_bump_ctr does, without its launch
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(S.ticket, 1u) == gridDim.x - 1) {
                *S.ticket = 0u;
                *reinterpret_cast<unsigned long long *>(S.ctr_dev) += 1ull;
                if (S.reset_count) *S.reset_count = 0;
            }
        }
    }
}

int check_prefetch(const RsConfig *cfg, const RsState *st) {
    if (!st->nx_src || !st->nx_det || !st->nx_rad || !st->nx_best || !st->nx_obs || !st->nx_seq || !st->refill_list ||
        !st->refill_count || (cfg->k_max > 0 && (!st->nx_dsrc || !st->nx_dsf)))
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
    if (cfg->k_max > 0 && (!st->rects || !st->dsrc || !st->vis || !st->dsf)) return fail("RsState obstruction tables are NULL");
    if (cfg->standardize < 0 || cfg->standardize > 2) return fail("standardize must be 0, 1 or 2");
    if (cfg->standardize && (!st->st_mean || !st->st_m2)) return fail("standardize needs RsState.st_mean / st_m2");
    return 0;
}

// envs per CTA of the multi-agent tile kernel: 128 threads cover 128 / 96 / 128 / ... / 256 (env, agent) units
int step_tile_envs(int n_agents) { return n_agents == 2 ? 64 : 32; }
size_t query_smem(const RsConfig *cfg) { return (size_t)cfg->k_max * kBlock * sizeof(int4); }
bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
size_t reset_smem(const RsConfig *cfg, int tb = kBlock) {
    return (size_t)tb * cfg->k_max * (sizeof(int4) + 4 * sizeof(double) + 4 * sizeof(uint32_t));
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
    if ((flags & RS_F_ZERO_REFILL) && !st->refill_count) return fail("RS_F_ZERO_REFILL needs RsState.refill_count");
    const int parity = (flags & RS_F_PARITY1) ? 1 : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if ((flags & RS_F_AUTO_RESET) && !(flags & RS_F_DEVICE_CTR)) {     // with RS_F_DEVICE_CTR rs_bump_ctr empties the list
        cudaError_t e = cudaMemsetAsync(st->reset_count, 0, sizeof(int32_t), s);
        if (e != cudaSuccess) return (int)e;
    }
    if (flags & RS_F_ZERO_REFILL) {         // a new refill list starts with this step (the step kernel itself appends to it)
        cudaError_t e = cudaMemsetAsync(st->refill_count + parity, 0, sizeof(int32_t), s);
        if (e != cudaSuccess) return (int)e;
    }
    if ((flags & RS_F_PREFETCH) && (flags & RS_F_AUTO_RESET)) {
        if (int rc = check_prefetch(cfg, st)) return rc;
    }
    rs::Params P = rs::make_params(*cfg);
    rs::StepArgs a;
    a.actions = actions; a.obs = obs; a.reward = reward; a.team_reward = team_reward; a.final_obs = final_obs;
    a.done = done; a.info = info; a.ended = ended; a.n_env = n_env; a.env_id0 = env_id0; a.seed = seed;
    a.step_ctr = step_ctr; a.uniforms = uniforms; a.n_uniforms = n_uniforms; a.flags = flags; a.parity = parity;
    const bool fast = (flags & RS_F_FAST_POISSON) && !uniforms;
    const int A = cfg->n_agents, K = cfg->k_max;
    if (A == 1) {
        // one thread per environment (rs_step1.cuh); KMAX = the unroll bound of the per-rectangle loops
#ifndef RS_STEP1_TB
#define RS_STEP1_TB 128
#endif
        constexpr int TB = RS_STEP1_TB;
        const int grid = (n_env + TB - 1) / TB;
        // 16-byte alignment of what the bulk copies and the float4 stores touch (tile offsets are multiples of 128 envs)
        const int bulk_ok = aligned16(st->rects) && aligned16(st->dsf) && aligned16(st->best) && aligned16(st->rad) &&
                            aligned16(obs) && n_env % 4 == 0;
#define RS_LAUNCH_STEP1(FAST, KM, OCC) step1_kernel<FAST, KM, TB, OCC><<<grid, TB, 0, s>>>(P, *st, a, bulk_ok)
#define RS_LAUNCH_STEP1_K(KM, OCC)                                     \
    do {                                                               \
        if (fast) RS_LAUNCH_STEP1(true, KM, OCC);                      \
        else RS_LAUNCH_STEP1(false, KM, OCC);                          \
    } while (0)
        // CTAs per SM the register allocation aims at: 7 x 128 threads hold the 886 environments per SM of the headline
        // workload in one wave (72 registers)
#ifndef RS_STEP1_OCC
#define RS_STEP1_OCC 7
#endif
        if (K == 0) RS_LAUNCH_STEP1_K(0, RS_STEP1_OCC);
        else if (K <= 3) RS_LAUNCH_STEP1_K(3, RS_STEP1_OCC);
        else if (K <= 5) {
            // up to 4 CTAs per SM there is room for 128 registers: no spills, 8 % faster per launch (65 536 envs: 26.6 -> 24.6 us)
            if ((long long)n_env <= 148LL * 4 * TB && RS_STEP1_OCC > 4) RS_LAUNCH_STEP1_K(5, 4);
            else RS_LAUNCH_STEP1_K(5, RS_STEP1_OCC);
        }
        else RS_LAUNCH_STEP1_K(8, 5);
#undef RS_LAUNCH_STEP1_K
#undef RS_LAUNCH_STEP1
        return (int)cudaGetLastError();
    }
    // several agents per environment: the tile program (rs_step_tiled.cuh), 64 / 32 environments per CTA of 128 threads
    const int E = step_tile_envs(cfg->n_agents);
    const int grid = (n_env + E - 1) / E;
    const rs::TileLayout L = rs::make_layout(E, A, K, kBlock, cfg->standardize);
    const size_t smem = (size_t)L.total;
    // the tile's rows: state rows read (src, rad, meta, actions, rects[k], then det / best / aflags / running count
    // statistics per agent) and rows written (meta, obs, reward, team_reward, done, info, ended, raw counts, then the
    // per-agent state rows)
    TileRows R;
    R.n_in = R.n_out = 0;
    const size_t Nn = (size_t)n_env;
    auto add = [](RowEnt *tab, int &cnt, const void *g, int off, int bpe) {
        if (g) tab[cnt++] = RowEnt{(unsigned long long)reinterpret_cast<uintptr_t>(g), (uint32_t)off, (uint32_t)bpe};
    };
    add(R.in, R.n_in, st->src, L.src, 8);
    add(R.in, R.n_in, st->rad, L.rad, 8);
    add(R.in, R.n_in, st->meta, L.meta, 4);
    add(R.in, R.n_in, actions, L.act, A * 4);
    for (int k = 0; k < K; k++) add(R.in, R.n_in, st->rects + (size_t)k * Nn * 4, L.rects + k * E * 16, 16);
    add(R.out, R.n_out, st->meta, L.meta, 4);
    add(R.out, R.n_out, obs, L.obs, A * RS_OBS_DIM * 4);
    add(R.out, R.n_out, reward, L.reward, A * 4);
    add(R.out, R.n_out, team_reward, L.team, 4);
    add(R.out, R.n_out, done, L.done, A);
    add(R.out, R.n_out, info, L.info, A);
    add(R.out, R.n_out, ended, L.ended, 1);
    if (cfg->standardize) add(R.out, R.n_out, st->raw_count, L.raw, A * 4);
    for (int ag = 0; ag < A; ag++) {
        for (int dir = 0; dir < 2; dir++) {
            RowEnt *tab = dir ? R.out : R.in;
            int &cnt = dir ? R.n_out : R.n_in;
            add(tab, cnt, st->det + ((size_t)ag * Nn) * 2, L.det + ag * E * 8, 8);
            add(tab, cnt, st->best + (size_t)ag * Nn, L.best + ag * E * 8, 8);
            add(tab, cnt, st->aflags + (size_t)ag * Nn, L.af + ag * E * 4, 4);
            if (cfg->standardize) {
                add(tab, cnt, st->st_mean + (size_t)ag * Nn, L.stm + ag * E * 8, 8);
                add(tab, cnt, st->st_m2 + (size_t)ag * Nn, L.stq + ag * E * 8, 8);
            }
        }
    }
    // bytes of one full tile's state rows (what the bulk copies of a CTA deliver to its mbarrier)
    uint32_t tx_bytes = 0;
    for (int i = 0; i < R.n_in; i++) tx_bytes += (uint32_t)E * R.in[i].bpe;
    // bulk-async tile copies need 16-byte aligned rows: every array base (the tile offsets are multiples of 32 elements)
    const int bulk_ok = aligned16(st->src) && aligned16(st->rad) && aligned16(st->rects) && aligned16(st->meta) &&
                        aligned16(st->det) && aligned16(st->best) && aligned16(st->aflags) && aligned16(actions) &&
                        aligned16(obs) && aligned16(reward) && aligned16(team_reward) && aligned16(done) &&
                        aligned16(info) && aligned16(ended) && aligned16(st->st_mean) && aligned16(st->st_m2) &&
                        aligned16(st->raw_count) && n_env % 4 == 0;     // per-agent rows start at multiples of N elements
#define RS_LAUNCH_STEP(FAST, TE)                                                                                          \
    do {                                                                                                                  \
        if (smem > 48 * 1024)                                                                                             \
            cudaFuncSetAttribute(step_kernel<FAST, TE, 6, kBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        step_kernel<FAST, TE, 6, kBlock><<<grid, kBlock, smem, s>>>(P, *st, a, L, R, bulk_ok, tx_bytes);                  \
    } while (0)
    if (E == 64) { if (fast) RS_LAUNCH_STEP(true, 64); else RS_LAUNCH_STEP(false, 64); }
    else { if (fast) RS_LAUNCH_STEP(true, 32); else RS_LAUNCH_STEP(false, 32); }
#undef RS_LAUNCH_STEP
    return (int)cudaGetLastError();
}

static int launch_reset(const RsConfig *cfg, const RsState *st, const rs::ResetArgs &a, const uint8_t *mask,
                        const uint8_t *new_mask, int flags, const int32_t *list, const int32_t *count, cudaStream_t s) {
    rs::Params P = rs::make_params(*cfg);
    int need = (a.n_env + 3) / 4;                       // one warp per env is the widest teaming
    int cap = kResetGrid;
#ifndef RS_PREPARE_NL
#define RS_PREPARE_NL 1
#endif
    const int prepare_nl = RS_PREPARE_NL;               // rs_prepare: lanes per environment (1: throughput, not latency)
    int tb = kBlock;
    if (a.prepare) {
        // rs_prepare shares the GPU with the step kernels of the other env batches for its whole (latency-bound) life.  Its
        // threads are packed into the largest CTAs the scratch columns allow (512 threads for k_max <= 6), i.e. onto as few
        // SMs as possible: a few SMs given over to it cost the step kernels less than one fat CTA on every fourth SM
        // (measured, 131 072 envs: 128 / 256 / 512 threads per CTA = 36.0 / 32.7 / 32.4 us per step)
        const size_t per_thread = reset_smem(cfg, 1);
        tb = 512 * per_thread <= 200 * 1024 ? 512 : (256 * per_thread <= 200 * 1024 ? 256 : 128);
#ifdef RS_PREPARE_TB
        tb = RS_PREPARE_TB;
#endif
        need = (a.n_env + tb / prepare_nl - 1) / (tb / prepare_nl);
        cap = 148 * (prepare_nl > 1 ? 4 : 1) * kBlock / tb;
    }
    // single-agent steps adopt prefetched episodes themselves (step1_kernel): their work list only holds the stragglers.
    // That launch follows every step, mostly to find an empty list (CTA size, register cap and grid are build switches:
    // 32-thread CTAs and grids of 16 / 37 measured within 2 % of this shape, a 72-register build 7 % slower)
#ifndef RS_STRAG_TB
#define RS_STRAG_TB 128
#endif
#ifndef RS_STRAG_MINB
#define RS_STRAG_MINB 1
#endif
#ifndef RS_STRAG_CAP
#define RS_STRAG_CAP 148
#endif
    const bool stragglers = list && (flags & RS_F_PREFETCH) && cfg->n_agents == 1 && !a.prepare;
    if (stragglers) { tb = RS_STRAG_TB; cap = RS_STRAG_CAP; }
    const int grid = need < cap ? need : cap;
    const size_t smem = reset_smem(cfg, tb);
    const bool fast = (flags & RS_F_FAST_POISSON) && !a.uniforms;
#define RS_LAUNCH_RESET(FAST, TBV, MINB)                                                                                    \
    do {                                                                                                                    \
        if (smem > 48 * 1024)                                                                                               \
            cudaFuncSetAttribute(reset_kernel<FAST, TBV, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        reset_kernel<FAST, TBV, MINB><<<grid, TBV, smem, s>>>(P, *st, a, mask, new_mask, flags, list, count, prepare_nl);   \
    } while (0)
#define RS_LAUNCH_RESET2(TBV, MINB) do { if (fast) RS_LAUNCH_RESET(true, TBV, MINB); else RS_LAUNCH_RESET(false, TBV, MINB); } while (0)
    if (stragglers) RS_LAUNCH_RESET2(RS_STRAG_TB, RS_STRAG_MINB);
    else if (a.prepare && tb == 512) RS_LAUNCH_RESET2(512, 1);
    else if (a.prepare && tb == 256) RS_LAUNCH_RESET2(256, 2);
    else if (tb == kBlock) RS_LAUNCH_RESET2(kBlock, RS_RESET_MINB);
    else return fail("rs_prepare: unsupported RS_PREPARE_TB");
#undef RS_LAUNCH_RESET2
#undef RS_LAUNCH_RESET
    return (int)cudaGetLastError();
}

int rs_reset(const RsConfig *cfg, const RsState *st, const uint8_t *reset_mask, const uint8_t *new_obstacles_mask,
             float *obs, int32_t n_env, uint32_t env_id0, uint64_t seed, uint64_t step_ctr, const double *uniforms,
             int32_t n_uniforms, int32_t flags, void *stream) {
    if (int rc = check_cfg(cfg, st, n_env)) return rc;
    if (!obs) return fail("obs is NULL");
    if (uniforms && n_uniforms < 2) return fail("n_uniforms must be >= 2 when uniforms are injected");
    if ((flags & RS_F_RESET_LIST) && (!st->reset_list || !st->reset_count)) return fail("list reset needs reset_list/reset_count");
    if ((flags & RS_F_BUMP_CTR) && (!st->ticket || !st->ctr_dev)) return fail("RS_F_BUMP_CTR needs RsState.ticket and ctr_dev");
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
    const size_t smem = query_smem(cfg);
    if (smem > 48 * 1024) cudaFuncSetAttribute(sp_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sp_query_kernel<<<grid, kBlock, smem, static_cast<cudaStream_t>(stream)>>>(P, *st, pts, out, n_env, variant);
    return (int)cudaGetLastError();
}

const char *rs_last_error(void) { return g_err; }
int rs_version(void) { return RS_VERSION; }
int rs_sizeof_config(void) { return (int)sizeof(RsConfig); }
int rs_sizeof_state(void) { return (int)sizeof(RsState); }

}  // extern "C"
