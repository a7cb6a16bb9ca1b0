/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the RAD-TEAM map observation.
 *
 * Follows /root/reference/algos/multiagent/NeuralNetworkCores/RADTEAM_core.py (M: below), one MapsBuffer at a time,
 * the way the reference runs: MapsBuffer.observation_to_map M:532-616, _inflate_coordinates M:692-715, the _update_*
 * methods M:748-932, IntensityEstimator M:101-182 (median of the samples of a grid cell), StatisticStandardization
 * M:188-277, Normalizer.normalize_incremental_logscale M:322-365, MapsBuffer.reset / _clear_maps M:513-530, 618-667.
 * Module switches as shipped: PFGRU = True (M:39), SIMPLE_NORMALIZATION = False (M:58), NORMALIZE_RADIATION = False
 * (M:59).  Pinned against the reference class itself: tests/golden/ref_maps_*.npz (tests/golden/make_golden.py maps).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link or call this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "radsearch_oracle.h"

/* int(v * resolution_accuracy) M:704-713 (Python int() truncates toward zero) */
static int32_t inflate(double v, double ra) { return (int32_t)(v * ra); }

/* numpy index semantics of map[x][y]: negative indices wrap once, anything else out of range raises (-> skipped, flagged) */
static int cell_index(const OrcMaps *m, int32_t cx, int32_t cy, int32_t *out) {
    if (cx < 0) cx += m->dim_x;
    if (cy < 0) cy += m->dim_y;
    if (cx < 0 || cy < 0 || cx >= m->dim_x || cy >= m->dim_y) return 0;
    *out = cx * m->dim_y + cy;
    return 1;
}

void orc_maps_reset(OrcMaps *m) {                                          /* M:513-523, 618-667; tools.reset M:378-385 */
    size_t cells = (size_t)m->dim_x * m->dim_y;
    memset(m->maps, 0, sizeof(float) * 7 * cells);
    memset(m->shadow, 0, sizeof(int32_t) * cells);
    m->log_len = 0;
    for (int a = 0; a < ORC_MAX_A; a++) m->last_cell[a] = -1;
    m->last_pred = -1;
    m->std_mean = 0.0; m->std_m2 = 0.0; m->std_std = 1.0; m->std_count = 0;
    m->status = 0;
}

OrcMaps *orc_maps_new(int32_t dim_x, int32_t dim_y, int32_t n_agents, int32_t steps_per_episode,
                      double resolution_accuracy) {
    OrcMaps *m = (OrcMaps *)calloc(1, sizeof(OrcMaps));
    size_t cells = (size_t)dim_x * dim_y;
    m->dim_x = dim_x; m->dim_y = dim_y; m->n_agents = n_agents;
    m->base = (steps_per_episode + 1) * n_agents;                          /* M:500 */
    m->ra = resolution_accuracy;
    m->log_cap = 4 * m->base + 64;
    m->maps = (float *)calloc(7 * cells, sizeof(float));
    m->shadow = (int32_t *)calloc(cells, sizeof(int32_t));
    m->log_cell = (int32_t *)calloc((size_t)m->log_cap, sizeof(int32_t));
    m->log_val = (double *)calloc((size_t)m->log_cap, sizeof(double));
    orc_maps_reset(m);
    return m;
}

void orc_maps_free(OrcMaps *m) {
    if (!m) return;
    free(m->maps); free(m->shadow); free(m->log_cell); free(m->log_val); free(m);
}

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* statistics.median of the readings logged at a cell (IntensityEstimator.get_estimate M:160-167) */
static double cell_median(const OrcMaps *m, int32_t cell) {
    double *buf = (double *)malloc(sizeof(double) * (size_t)(m->log_len > 0 ? m->log_len : 1));
    int n = 0;
    for (int i = 0; i < m->log_len; i++)
        if (m->log_cell[i] == cell) buf[n++] = m->log_val[i];
    qsort(buf, (size_t)n, sizeof(double), cmp_double);
    double r = (n & 1) ? buf[n / 2] : (buf[n / 2 - 1] + buf[n / 2]) / 2.0;
    free(buf);
    return r;
}

/* MapsBuffer.observation_to_map(observation, id, loc_prediction) M:532-616.
 * obs[a][11] = the agents' observations in dict order; pred = the (deflated) source location predicted for agent `id`.
 * Map order in m->maps: 0 prediction, 1 location, 2 others, 3 readings, 4 visits, 5 obstacles, 6 combined (M:606-616). */
void orc_maps_observation_to_map(OrcMaps *m, const double *obs, int32_t id, const double pred[2]) {
    const int A = m->n_agents;
    const size_t cells = (size_t)m->dim_x * m->dim_y;
    float *prediction = m->maps, *location = m->maps + cells, *others = m->maps + 2 * cells,
          *readings = m->maps + 3 * cells, *visits = m->maps + 4 * cells, *obstacles = m->maps + 5 * cells,
          *combined = m->maps + 6 * cells;
    /* M:541-545: every agent's reading goes to the sample table of its cell first */
    for (int a = 0; a < A; a++) {
        const double *o = obs + (size_t)a * ORC_OBS_DIM;
        int32_t c;
        if (!cell_index(m, inflate(o[1], m->ra), inflate(o[2], m->ra), &c)) { m->status |= 1u; c = -1; }
        if (m->log_len < m->log_cap) {
            m->log_cell[m->log_len] = c;
            m->log_val[m->log_len] = o[0];
            m->log_len++;
        } else m->status |= 2u;
    }
    for (int a = 0; a < A; a++) {                                          /* M:547-604 */
        const double *o = obs + (size_t)a * ORC_OBS_DIM;
        int32_t c, pc;
        const int have_c = cell_index(m, inflate(o[1], m->ra), inflate(o[2], m->ra), &c);
        const int have_p = cell_index(m, inflate(pred[0], m->ra), inflate(pred[1], m->ra), &pc);
        if (!have_p) m->status |= 4u;
        const int32_t last = m->last_cell[a];
        /* prediction map (PFGRU) M:564-568, 748-766 */
        if (m->last_pred >= 0) prediction[m->last_pred] -= 1.0f;
        if (have_p) prediction[pc] = 1.0f;
        if (!have_c) { m->last_pred = have_p ? pc : -1; continue; }         /* the reference raises IndexError here */
        /* location maps M:570-588, 768-848 */
        if (a == id) {
            if (last >= 0) location[last] -= 1.0f;
            location[c] = 1.0f;
        } else {
            if (last >= 0) others[last] -= 1.0f;
            others[c] += 1.0f;
        }
        if (last >= 0) combined[last] -= 1.0f;
        combined[c] += 1.0f;
        /* readings map M:850-884: median estimate -> running standardisation -> map */
        {
            const double est = cell_median(m, c);
            m->std_count += 1;                                             /* StatisticStandardization.update M:215-252 */
            if (m->std_count == 1) m->std_mean = est;
            else {
                const double mean_new = m->std_mean + (est - m->std_mean) / (double)m->std_count;
                const double m2_new = m->std_m2 + (est - m->std_mean) * (est - mean_new);
                m->std_mean = mean_new;
                m->std_m2 = m2_new;
                const double sd = sqrt(m2_new / (double)(m->std_count - 1));
                m->std_std = sd > 1.0 ? sd : 1.0;
            }
            readings[c] = (float)((est - m->std_mean) / m->std_std);       /* standardize M:265; Map dtype float32 */
        }
        /* visit counts map M:886-916: shadow counter steps by 2, log-scale normalisation M:356-360 */
        {
            const int32_t current = m->shadow[c];
            m->shadow[c] = current + 2;
            const double v = (log((double)(2 + current)) / log((double)m->base)) * 1.0 /
                             (log((double)(2 * m->base)) / log((double)m->base));
            visits[c] = (float)v;
        }
        /* obstacles map M:590-599, 918-932: the last non-zero detection wins */
        for (int d = 3; d < ORC_OBS_DIM; d++)
            if (o[d] != 0.0) obstacles[c] = (float)o[d];
        m->last_cell[a] = c;                                               /* M:602-603 */
        m->last_pred = have_p ? pc : -1;
    }
}

const float *orc_maps_data(const OrcMaps *m) { return m->maps; }
uint32_t orc_maps_status(const OrcMaps *m) { return m->status; }
