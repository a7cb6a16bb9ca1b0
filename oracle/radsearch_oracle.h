/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference hot path.
 *
 * Follows /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py (RadSearch.step / reset and their
 * helpers) and /root/reference/algos/multiagent/ppo.py (PPOBuffer.GAE_advantage_and_rewardsToGO) one scalar
 * environment at a time, the way the reference runs.  Geometry follows the exact-arithmetic reading of the
 * un-vendored `visilibity` dependency (see oracle/shims/visilibity.py): PARITY UNPINNED against the real library.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may link or call this.
 */
#ifndef RADSEARCH_ORACLE_H
#define RADSEARCH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_K 8   /* obstruction slots (env allows 0..7)             rad_search_env.py:317 */
#define ORC_MAX_A 8   /* agents per environment                          rad_search_env.py:353 */
#define ORC_OBS_DIM 11

/* status bits (soft failures; the reference would raise, assert or hang) */
#define ORC_ST_REJECT_CAP 1u     /* a rejection-sampling loop hit its cap                          */
#define ORC_ST_LAMBDA_INF 2u     /* detector exactly on the source: intensity/0 (reference raises)  */
#define ORC_ST_UNIFORMS_OUT 4u   /* injected uniform stream exhausted                              */
#define ORC_ST_CORRECT_MISS 8u   /* correct_coords would never terminate (reference hangs)         */
#define ORC_ST_WALL_ASSERT 16u   /* `assert dists[i] == 0.0` in obstruction_sensors would fire      */

typedef struct {
    int32_t bbox[4];        /* x0,y0,x1,y1 of the arena                      default 0,0,2700,2700 */
    int32_t obs_area[2];    /* observation_area                               default 200,500       */
    int32_t enforce;        /* enforce_grid_boundaries                                              */
    int32_t n_agents;
    int32_t obstruction_count; /* -1 random 1..5, 0..7 fixed                                       */
    int32_t count_law;      /* 0 = reference: intensity/dist + bkg ; 1 = inverse square            */
    int32_t max_ep_len;     /* caller rule: timeout (train.py:394)            default 120           */
} OrcConfig;

typedef struct {
    int32_t num_obs;
    int32_t rect[ORC_MAX_K][4];   /* x0,y0,x1,y1 */
    int32_t src[2];
    int32_t intensity, bkg;
    int32_t det[ORC_MAX_A][2];
    double best[ORC_MAX_A];       /* Agent.prev_det_dist (running minimum)    */
    double sp[ORC_MAX_A];         /* Agent.sp_dist                           */
    double euc[ORC_MAX_A];        /* Agent.euc_dist                          */
    int32_t oob[ORC_MAX_A];       /* Agent.out_of_bounds  (this step)        */
    int32_t oob_count[ORC_MAX_A]; /* Agent.out_of_bounds_count               */
    int32_t blocked[ORC_MAX_A];   /* Agent.obstacle_blocking (sticky)        */
    int32_t collision[ORC_MAX_A];
    int32_t los_blocked[ORC_MAX_A]; /* Agent.intersect                       */
    int32_t done;                 /* RadSearch.done (sticky, shared)         */
    int32_t iter_count;
    int32_t ep_len;               /* caller-side steps_in_episode            */
    uint32_t status;
    uint32_t episode;             /* resets so far: keys the reset draws     */
} OrcEnv;

typedef struct {
    double obs[ORC_MAX_A][ORC_OBS_DIM];
    double reward[ORC_MAX_A];
    double team_reward;           /* NaN when the reference leaves it None   */
    int32_t done[ORC_MAX_A];
    double lam[ORC_MAX_A];        /* Poisson mean that was sampled           */
} OrcStepOut;

/* uniform source: injected doubles (u != NULL) or the Philox4x32-10 stream (seed, env_id, step_ctr, domain) */
typedef struct {
    const double *u;
    int32_t n_u, pos;
    uint32_t key[2], ctr[4];
    uint32_t buf[4];
    int32_t have;
    uint32_t *status;
} OrcRng;

void orc_default_config(OrcConfig *c);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_rng_philox(OrcRng *r, uint64_t seed, uint32_t env_id, uint32_t domain, uint32_t agent, uint64_t step_ctr,
                    uint32_t *status);
void orc_rng_inject(OrcRng *r, const double *u, int32_t n, uint32_t *status);
uint32_t orc_rng_u32(OrcRng *r);
double orc_rng_double(OrcRng *r);
uint32_t orc_rng_below(OrcRng *r, uint32_t range);

int64_t orc_poisson(OrcRng *r, double lam);
double orc_loggam(double x);
double orc_round2(double x);

/* geometry predicates (exported for the unit tests) */
int orc_seg_hits_open_rect(const int32_t p[2], const int32_t q[2], const int32_t r[4]);
int orc_seg_touches_seg(const int32_t a[2], const int32_t b[2], const int32_t c[2], const int32_t d[2]);
int orc_los_blocked_rect(const int32_t p[2], const int32_t q[2], const int32_t r[4]);
int orc_in_rect_closed(const int32_t p[2], const int32_t r[4]);
double orc_shortest_path(const OrcEnv *e, const int32_t det[2]);
void orc_sensors(const OrcConfig *c, OrcEnv *e, int agent, double out[8]);

void orc_step(const OrcConfig *c, OrcEnv *e, const int32_t *actions /* NULL = step(None); per agent 0..8 */,
              OrcRng *rngs /* one per agent */, OrcStepOut *out);
/* the draws of a reset are keyed by (seed, env_id, ++e->episode): a scenario does not depend on when it is computed */
void orc_reset(const OrcConfig *c, OrcEnv *e, int new_obstacles, uint64_t seed, uint32_t env_id,
               const double *inj_u, int32_t n_inj, OrcStepOut *out);
void orc_load_scenario(const OrcConfig *c, OrcEnv *e, const int32_t src[2], const int32_t det[2], int32_t intensity,
                       int32_t bkg, const int32_t *rects, int32_t num_obs);

/* batched drivers (OpenMP over environments) used by the parity tests and the CPU baseline */
void orc_step_batch(const OrcConfig *c, OrcEnv *envs, int32_t n, const int32_t *actions, uint64_t seed,
                    uint32_t env_id0, uint64_t step_ctr, const double *inj_u, int32_t n_inj, OrcStepOut *outs,
                    int32_t threads);
void orc_reset_batch(const OrcConfig *c, OrcEnv *envs, int32_t n, const uint8_t *mask, const uint8_t *new_obs_mask,
                     uint64_t seed, uint32_t env_id0, const double *inj_u, int32_t n_inj, OrcStepOut *outs,
                     int32_t threads);
/* rollout with the caller rules of train.py:394-405,446-548 (auto-reset, epoch end); returns env-steps done */
int64_t orc_rollout(const OrcConfig *c, OrcEnv *envs, int32_t n, int32_t T, uint64_t seed, uint32_t env_id0,
                    uint64_t step_ctr0, int32_t epoch_end_last, int32_t threads, double *checksum);

void orc_gae(const float *rew, const float *val, const uint8_t *path_end, const float *boot, float *adv, float *ret,
             int32_t T, int32_t N, double gamma, double lam, int32_t threads);
/* per-episode running standardiser of the count channel (RADTEAM_core.py:188-277 / test_environment/core.py:55-79) */
typedef struct {
    double mean, m2, std;
    int32_t count;
} OrcStat;
double orc_stat_update_standardize(OrcStat *s, double x, int32_t mode /* 1 RAD-TEAM rule, 2 StatBuff + clip 8 */);
void orc_stat_reset(OrcStat *s);

/* RAD-TEAM map observation: one MapsBuffer (RADTEAM_core.py:394-932), see oracle/maps_oracle.c */
typedef struct {
    int32_t dim_x, dim_y, n_agents, base;
    double ra;                        /* resolution_accuracy                                   */
    float *maps;                      /* [7][dim_x][dim_y]: prediction, location, others, readings, visits, obstacles, combined */
    int32_t *shadow;                  /* [dim_x][dim_y] visit_counts_shadow                    */
    int32_t *log_cell;                /* sample table of the IntensityEstimator: cell of reading i */
    double *log_val;                  /*                                         value of reading i */
    int32_t log_len, log_cap;
    int32_t last_cell[ORC_MAX_A];     /* tools.last_coords (-1 = none)                         */
    int32_t last_pred;                /* tools.last_prediction (-1 = none)                     */
    double std_mean, std_m2, std_std; /* tools.standardizer                                    */
    int32_t std_count;
    uint32_t status;                  /* 1 agent cell outside the map, 2 sample table full, 4 prediction outside the map */
} OrcMaps;
OrcMaps *orc_maps_new(int32_t dim_x, int32_t dim_y, int32_t n_agents, int32_t steps_per_episode,
                      double resolution_accuracy);
void orc_maps_free(OrcMaps *m);
void orc_maps_reset(OrcMaps *m);
void orc_maps_observation_to_map(OrcMaps *m, const double *obs /* [A][11] */, int32_t id, const double pred[2]);
const float *orc_maps_data(const OrcMaps *m);
uint32_t orc_maps_status(const OrcMaps *m);

int32_t orc_sizeof_env(void);
int32_t orc_sizeof_out(void);

#ifdef __cplusplus
}
#endif
#endif
