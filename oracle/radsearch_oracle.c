/* TEST INFRASTRUCTURE ONLY -- see radsearch_oracle.h.  Plain C99 + OpenMP; build: oracle/build_oracle.py.
 *
 * Reference citations use R: = /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py and
 * P: = /root/reference/algos/multiagent/ppo.py.
 */
#include "radsearch_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------------------ */
/* configuration                                                                                                 */
/* ------------------------------------------------------------------------------------------------------------ */
void orc_default_config(OrcConfig *c) {
    c->bbox[0] = 0; c->bbox[1] = 0; c->bbox[2] = 2700; c->bbox[3] = 2700;   /* R:321-325 */
    c->obs_area[0] = 200; c->obs_area[1] = 500;                             /* R:326     */
    c->enforce = 1;                                                         /* main.py:314 default */
    c->n_agents = 1;
    c->obstruction_count = 5;
    c->count_law = 0;
    c->max_ep_len = 120;                                                    /* main.py:243 */
}

/* search area corners R:393-420: [bx0+lo, by0+lo] .. [bx1-hi, by1-hi] */
static inline int sa_x0(const OrcConfig *c) { return c->bbox[0] + c->obs_area[0]; }
static inline int sa_y0(const OrcConfig *c) { return c->bbox[1] + c->obs_area[0]; }
static inline int sa_x1(const OrcConfig *c) { return c->bbox[2] - c->obs_area[1]; }
static inline int sa_y1(const OrcConfig *c) { return c->bbox[3] - c->obs_area[1]; }
/* max_dist = dist(search_area[2], search_area[1]) R:423-425 */
static inline double max_dist(const OrcConfig *c) { return sqrt((double)(sa_y1(c) - sa_y0(c)) * (double)(sa_y1(c) - sa_y0(c))); }

/* get_step R:205-224: 0 left, 1 up-left, 2 up, 3 up-right, 4 right, 5 down-right, 6 down, 7 down-left, 8 idle */
static const int STEP[9][2] = {{-100, 0}, {-71, 71}, {0, 100}, {71, 71}, {100, 0}, {71, -71}, {0, -100}, {-71, -71}, {0, 0}};
/* unit coefficients get_x_step_coeff / get_y_step_coeff R:187-202 */
static const int COEF[8][2] = {{-1, 0}, {-1, 1}, {0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}};

/* ------------------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) and the uniform stream                                                   */
/* ------------------------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_rng_philox(OrcRng *r, uint64_t seed, uint32_t env_id, uint32_t domain, uint32_t agent, uint64_t step_ctr,
                    uint32_t *status) {
    r->u = NULL; r->n_u = 0; r->pos = 0;
    r->key[0] = (uint32_t)seed; r->key[1] = (uint32_t)(seed >> 32);
    r->ctr[0] = env_id;
    r->ctr[1] = (domain << 24) | (agent << 16);
    r->ctr[2] = (uint32_t)step_ctr; r->ctr[3] = (uint32_t)(step_ctr >> 32);
    r->have = 0;
    r->status = status;
}

void orc_rng_inject(OrcRng *r, const double *u, int32_t n, uint32_t *status) {
    memset(r, 0, sizeof(*r));
    r->u = u; r->n_u = n; r->pos = 0; r->status = status;
}

uint32_t orc_rng_u32(OrcRng *r) {
    if (r->have == 0) {
        orc_philox4x32_10(r->ctr, r->key, r->buf);
        r->ctr[1] += 1;          /* block index lives in the low 16 bits */
        r->have = 4;
    }
    uint32_t v = r->buf[4 - r->have];
    r->have -= 1;
    return v;
}

/* numpy next_double: (next_uint64 >> 11) * 2^-53 */
double orc_rng_double(OrcRng *r) {
    if (r->u) {
        if (r->pos >= r->n_u) {
            if (r->status) *r->status |= ORC_ST_UNIFORMS_OUT;
            return 0.5;
        }
        return r->u[r->pos++];
    }
    uint32_t lo = orc_rng_u32(r);
    uint32_t hi = orc_rng_u32(r);
    uint64_t w = ((uint64_t)hi << 32) | lo;
    return (double)(w >> 11) * (1.0 / 9007199254740992.0);
}

/* unbiased integer in [0, range) (Lemire 2019, the method numpy's Generator.integers uses for 32-bit ranges) */
uint32_t orc_rng_below(OrcRng *r, uint32_t range) {
    uint64_t m = (uint64_t)orc_rng_u32(r) * range;
    uint32_t l = (uint32_t)m;
    if (l < range) {
        uint32_t t = (uint32_t)(-range) % range;
        while (l < t) {
            m = (uint64_t)orc_rng_u32(r) * range;
            l = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* numpy Generator.poisson restated (numpy/random/src/distributions/distributions.c: random_poisson,             */
/* random_poisson_ptrs [Hoermann 1993 PTRS], random_poisson_mult, random_loggam).  R:498 is the call site.        */
/* ------------------------------------------------------------------------------------------------------------ */
double orc_loggam(double x) {
    static const double a[10] = {8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04,
                                 -5.952380952380952e-04, 8.417508417508418e-04, -1.917526917526918e-03,
                                 6.410256410256410e-03, -2.955065359477124e-02, 1.796443723688307e-01,
                                 -1.39243221690590e+00};
    double x0, x2, gl, gl0;
    int64_t k, n;
    if (x == 1.0 || x == 2.0) return 0.0;
    n = (x < 7.0) ? (int64_t)(7 - x) : 0;
    x0 = x + n;
    x2 = (1.0 / x0) * (1.0 / x0);
    gl0 = a[9];
    for (k = 8; k >= 0; k--) {
        gl0 *= x2;
        gl0 += a[k];
    }
    gl = gl0 / x0 + 0.5 * 1.8378770664093453e+00 + (x0 - 0.5) * log(x0) - x0;
    if (x < 7.0) {
        for (k = 1; k <= n; k++) {
            gl -= log(x0 - 1.0);
            x0 -= 1.0;
        }
    }
    return gl;
}

int64_t orc_poisson(OrcRng *r, double lam) {
    if (lam >= 10) {
        double slam = sqrt(lam), loglam = log(lam);
        double b = 0.931 + 2.53 * slam;
        double a = -0.059 + 0.02483 * b;
        double invalpha = 1.1239 + 1.1328 / (b - 3.4);
        double vr = 0.9277 - 3.6224 / (b - 2);
        for (int it = 0; it < 1000; it++) {
            double U = orc_rng_double(r) - 0.5;
            double V = orc_rng_double(r);
            double us = 0.5 - fabs(U);
            int64_t k = (int64_t)floor((2 * a / us + b) * U + lam + 0.43);
            if (us >= 0.07 && V <= vr) return k;
            if (k < 0 || (us < 0.013 && V > us)) continue;
            if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + k * loglam - orc_loggam((double)(k + 1))))
                return k;
            if (r->u && r->pos >= r->n_u) break;
        }
        if (r->status) *r->status |= ORC_ST_UNIFORMS_OUT;
        return (int64_t)floor(lam);
    } else if (lam == 0) {
        return 0;
    } else {
        double enlam = exp(-lam), prod = 1.0;
        int64_t X = 0;
        for (;;) {
            prod *= orc_rng_double(r);
            if (prod > enlam) X += 1; else return X;
            if (r->u && r->pos >= r->n_u) { if (r->status) *r->status |= ORC_ST_UNIFORMS_OUT; return X; }
        }
    }
}

/* Python round(x, 2): correctly rounded decimal rounding of the exact binary value, ties to even.  R:613 */
double orc_round2(double x) {
    double p = x * 100.0;
    double e = fma(x, 100.0, -p);          /* exact: x*100 = p + e */
    double n = nearbyint(p);               /* ties-to-even on p */
    double diff = p - n;                   /* exact */
    if (diff == 0.5 || diff == -0.5) {
        /* p is exactly half way; the residual e decides, else keep ties-to-even */
        if (e > 0) n = floor(p) + 1.0;
        else if (e < 0) n = floor(p);
    } else if (diff == 0.0 && e != 0.0) {
        /* nothing to do: |e| < 0.5 */
    }
    return n / 100.0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* exact integer geometry on the lattice (every coordinate is an integer)                              */
/* ------------------------------------------------------------------------------------------------------------ */
int orc_in_rect_closed(const int32_t p[2], const int32_t r[4]) {
    /* vis Point._in(poly, 1e-7) for a lattice point and a lattice rectangle: closed containment  R:1061,1108,1155 */
    return r[0] <= p[0] && p[0] <= r[2] && r[1] <= p[1] && p[1] <= r[3];
}

static int in_rect_open(const int32_t p[2], const int32_t r[4]) {
    return r[0] < p[0] && p[0] < r[2] && r[1] < p[1] && p[1] < r[3];
}

/* open segment pq meets the open rectangle: max(enter) < min(exit) with rational parameters */
int orc_seg_hits_open_rect(const int32_t p[2], const int32_t q[2], const int32_t r[4]) {
    int64_t lo_n = 0, lo_d = 1, hi_n = 1, hi_d = 1; /* t in (lo, hi), denominators > 0 */
    for (int ax = 0; ax < 2; ax++) {
        int64_t s = p[ax], d = (int64_t)q[ax] - p[ax], a = r[ax], b = r[ax + 2];
        if (d == 0) {
            if (!(a < s && s < b)) return 0;
        } else {
            int64_t en, ex, den;
            if (d > 0) { en = a - s; ex = b - s; den = d; } else { en = s - b; ex = s - a; den = -d; }
            if (en * lo_d > lo_n * den) { lo_n = en; lo_d = den; }
            if (ex * hi_d < hi_n * den) { hi_n = ex; hi_d = den; }
        }
    }
    return lo_n * hi_d < hi_n * lo_d;
}

/* closed segment pq meets the closed rectangle */
static int seg_hits_closed_rect(const int32_t p[2], const int32_t q[2], const int32_t r[4]) {
    int64_t lo_n = 0, lo_d = 1, hi_n = 1, hi_d = 1;
    for (int ax = 0; ax < 2; ax++) {
        int64_t s = p[ax], d = (int64_t)q[ax] - p[ax], a = r[ax], b = r[ax + 2];
        if (d == 0) {
            if (!(a <= s && s <= b)) return 0;
        } else {
            int64_t en, ex, den;
            if (d > 0) { en = a - s; ex = b - s; den = d; } else { en = s - b; ex = s - a; den = -d; }
            if (en * lo_d > lo_n * den) { lo_n = en; lo_d = den; }
            if (ex * hi_d < hi_n * den) { hi_n = ex; hi_d = den; }
        }
    }
    return lo_n * hi_d <= hi_n * lo_d;
}

static inline int64_t cross64(int64_t ax, int64_t ay, int64_t bx, int64_t by) { return ax * by - ay * bx; }
static inline int sgn64(int64_t v) { return (v > 0) - (v < 0); }
static int on_seg(const int32_t a[2], const int32_t b[2], const int32_t p[2]) {
    /* p collinear with ab assumed */
    return (p[0] >= (a[0] < b[0] ? a[0] : b[0])) && (p[0] <= (a[0] > b[0] ? a[0] : b[0])) &&
           (p[1] >= (a[1] < b[1] ? a[1] : b[1])) && (p[1] <= (a[1] > b[1] ? a[1] : b[1]));
}

/* vis.intersect(seg, seg, 1e-7) on the lattice == the closed segments share a point (R:1205): the smallest non-zero
 * distance between a lattice ray of get_step() and a lattice edge is 1/(71*sqrt(2)) >> 1e-7. */
int orc_seg_touches_seg(const int32_t a[2], const int32_t b[2], const int32_t c[2], const int32_t d[2]) {
    int o1 = sgn64(cross64(b[0] - a[0], b[1] - a[1], c[0] - a[0], c[1] - a[1]));
    int o2 = sgn64(cross64(b[0] - a[0], b[1] - a[1], d[0] - a[0], d[1] - a[1]));
    int o3 = sgn64(cross64(d[0] - c[0], d[1] - c[1], a[0] - c[0], a[1] - c[1]));
    int o4 = sgn64(cross64(d[0] - c[0], d[1] - c[1], b[0] - c[0], b[1] - c[1]));
    if (o1 != o2 && o3 != o4) return 1;
    if (o1 == 0 && on_seg(a, b, c)) return 1;
    if (o2 == 0 && on_seg(a, b, d)) return 1;
    if (o3 == 0 && on_seg(c, d, a)) return 1;
    if (o4 == 0 && on_seg(c, d, b)) return 1;
    return 0;
}

/* vis.boundary_distance(Line_Segment(p,q), rect) < 0.001   (R:1110, 1141)
 * = the segment touches or crosses the rectangle's boundary, or a rectangle corner whose projection falls inside the
 *   segment lies within 0.001 of it (cross^2 * 1e6 < |pq|^2).  Endpoint-to-edge distances on the lattice are 0 or >= 1. */
int orc_los_blocked_rect(const int32_t p[2], const int32_t q[2], const int32_t r[4]) {
    if (seg_hits_closed_rect(p, q, r) && !(in_rect_open(p, r) && in_rect_open(q, r))) return 1;
    int64_t dx = (int64_t)q[0] - p[0], dy = (int64_t)q[1] - p[1];
    int64_t l2 = dx * dx + dy * dy;
    if (l2 == 0) return 0;
    const int32_t cx[4] = {r[0], r[0], r[2], r[2]}, cy[4] = {r[1], r[3], r[3], r[1]};
    for (int i = 0; i < 4; i++) {
        int64_t wx = (int64_t)cx[i] - p[0], wy = (int64_t)cy[i] - p[1];
        int64_t t = wx * dx + wy * dy;
        if (t < 0 || t > l2) continue;
        int64_t cr = cross64(wx, wy, dx, dy);
        if (cr < 0) cr = -cr;
        if (cr <= 3 && cr * cr * 1000000 < l2) return 1;
    }
    return 0;
}

/* Euclidean shortest path source -> detector around the rectangles (R:491-493, 774-776).  Dijkstra over
 * {source, 4*num_obs corners, detector}; lengths are accumulated from the source outwards, the order
 * Polyline::length() sums them in.  Two points see each other iff the open segment misses every open rectangle. */
static int visible_pts(const OrcEnv *e, const int32_t a[2], const int32_t b[2]) {
    for (int k = 0; k < e->num_obs; k++)
        if (orc_seg_hits_open_rect(a, b, e->rect[k])) return 0;
    return 1;
}

double orc_shortest_path(const OrcEnv *e, const int32_t det[2]) {
    int32_t node[4 * ORC_MAX_K + 2][2];
    int m = 0;
    node[m][0] = e->src[0]; node[m][1] = e->src[1]; m++;
    for (int k = 0; k < e->num_obs; k++) {
        const int32_t *r = e->rect[k];
        node[m][0] = r[0]; node[m][1] = r[1]; m++;
        node[m][0] = r[0]; node[m][1] = r[3]; m++;
        node[m][0] = r[2]; node[m][1] = r[3]; m++;
        node[m][0] = r[2]; node[m][1] = r[1]; m++;
    }
    node[m][0] = det[0]; node[m][1] = det[1]; m++;
    if (node[0][0] == det[0] && node[0][1] == det[1]) return 0.0;
    /* mutually visible (grazing corners allowed): VisiLibity returns the direct segment, whose length can differ by an
     * ulp from the sum over a corner that lies exactly on it */
    if (visible_pts(e, node[0], det)) {
        double dx = (double)node[0][0] - det[0], dy = (double)node[0][1] - det[1];
        return sqrt(dx * dx + dy * dy);
    }
    double dist[4 * ORC_MAX_K + 2];
    int fin[4 * ORC_MAX_K + 2];
    for (int i = 0; i < m; i++) { dist[i] = INFINITY; fin[i] = 0; }
    dist[0] = 0.0;
    for (;;) {
        int u = -1;
        for (int i = 0; i < m; i++)
            if (!fin[i] && dist[i] < INFINITY && (u < 0 || dist[i] < dist[u])) u = i;
        if (u < 0) return INFINITY;
        if (u == m - 1) return dist[u];
        fin[u] = 1;
        for (int w = 1; w < m; w++) {
            if (fin[w]) continue;
            if (!visible_pts(e, node[u], node[w])) continue;
            double dx = (double)node[u][0] - node[w][0], dy = (double)node[u][1] - node[w][1];
            double nd = dist[u] + sqrt(dx * dx + dy * dy);
            if (nd < dist[w]) dist[w] = nd;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* environment pieces                                                                                            */
/* ------------------------------------------------------------------------------------------------------------ */
/* in_obstruction R:1148-1170: first rectangle (index order) that contains the point (closed); blocked iff the point
 * is strictly inside that rectangle's bounding box. */
static int in_obstruction(const OrcEnv *e, const int32_t p[2]) {
    for (int k = 0; k < e->num_obs; k++)
        if (orc_in_rect_closed(p, e->rect[k])) return in_rect_open(p, e->rect[k]);
    return 0;
}

/* is_intersect R:1133-1146 (incl. the leftover isclose(sqrt(euc_dist), sp_dist, abs_tol=0.1) clause) */
static int is_intersect(const OrcEnv *e, int ag) {
    double a = sqrt(e->euc[ag]), b = e->sp[ag];
    double diff = fabs(a - b), big = fabs(a) > fabs(b) ? fabs(a) : fabs(b);
    double tol = 1e-09 * big > 0.1 ? 1e-09 * big : 0.1;
    int close = (a == b) || (isfinite(a) && isfinite(b) && diff <= tol);
    if (close) return 0;
    for (int k = 0; k < e->num_obs; k++)
        if (orc_los_blocked_rect(e->det[ag], e->src, e->rect[k])) return 1;
    return 0;
}

/* take_action R:876-946; returns 1 when the detector moved */
static int take_action(const OrcConfig *c, OrcEnv *e, int ag, int action, const int32_t (*proposed)[2], int n_prop) {
    if (action < 0) return 0;                                   /* action is None R:900 */
    int32_t tent[2] = {e->det[ag][0] + STEP[action][0], e->det[ag][1] + STEP[action][1]};
    int cnt = 0;
    for (int i = 0; i < n_prop; i++)
        if (proposed[i][0] == tent[0] && proposed[i][1] == tent[1]) cnt++;
    if (cnt > 1) { e->collision[ag] = 1; return 0; }            /* R:908-910 */
    int roll = 0;
    if (c->enforce) {                                           /* R:917-925 */
        if (tent[0] < c->bbox[0] || tent[1] < c->bbox[1] || c->bbox[2] <= tent[0] || c->bbox[3] <= tent[1]) {
            e->oob[ag] = 1; e->oob_count[ag] += 1; roll = 1;
        }
    } else {                                                    /* R:928-933: tested on the pre-move coordinates */
        int lower = e->det[ag][0] < sa_x0(c) || e->det[ag][1] < sa_y0(c);
        int upper = sa_x1(c) < e->det[ag][0] || sa_y1(c) < e->det[ag][1];
        if (lower || upper) { e->oob[ag] = 1; e->oob_count[ag] += 1; }
    }
    if (in_obstruction(e, tent)) { roll = 1; e->blocked[ag] = 1; }  /* R:935-937 */
    if (!roll) { e->det[ag][0] = tent[0]; e->det[ag][1] = tent[1]; }
    return !roll;
}

static double pt_edge_dist(const int32_t p[2], const int32_t a[2], const int32_t b[2]) {
    /* vis.distance(Point, Line_Segment) for an axis-aligned lattice edge: distance to the clamped projection R:1207 */
    int64_t dx, dy;
    if (a[0] == b[0]) {
        int32_t lo = a[1] < b[1] ? a[1] : b[1], hi = a[1] > b[1] ? a[1] : b[1];
        dx = (int64_t)p[0] - a[0];
        dy = p[1] < lo ? lo - p[1] : (p[1] > hi ? p[1] - hi : 0);
    } else {
        int32_t lo = a[0] < b[0] ? a[0] : b[0], hi = a[0] > b[0] ? a[0] : b[0];
        dy = (int64_t)p[1] - a[1];
        dx = p[0] < lo ? lo - p[0] : (p[0] > hi ? p[0] - hi : 0);
    }
    return sqrt((double)(dx * dx + dy * dy));
}

/* correct_coords R:1263-1306.  The nudged points are det + n*0.1*coef; inside-ness (closed, eps 1e-7) is decided in
 * integer tenths.  The loop runs until any direction is inside. */
static void correct_coords(OrcEnv *e, const int32_t det[2], const int32_t r[4], double dists[8]) {
    int xc[8] = {0};
    int any = 0;
    for (int n = 1; n <= 400000 && !any; n++) {
        for (int d = 0; d < 8; d++) {
            int64_t qx = 10 * (int64_t)det[0] + (int64_t)n * COEF[d][0], qy = 10 * (int64_t)det[1] + (int64_t)n * COEF[d][1];
            if (10 * (int64_t)r[0] <= qx && qx <= 10 * (int64_t)r[2] && 10 * (int64_t)r[1] <= qy && qy <= 10 * (int64_t)r[3]) {
                xc[d] = 1; any = 1;
            }
        }
    }
    for (int d = 0; d < 8; d++) dists[d] = 0.0;
    if (!any) { e->status |= ORC_ST_CORRECT_MISS; return; }
    int s = 0;
    for (int d = 0; d < 8; d++) s += xc[d];
    if (s >= 4) {
        for (int ii = 0; ii <= 6; ii += 2) {
            int lo = (ii + 7) % 8, hi = ii + 1;
            if (xc[lo] && xc[hi]) { dists[ii] = 1.0; dists[lo] = 1.0; dists[hi] = 1.0; }
        }
    }
}

/* obstruction_sensors R:1172-1261 */
void orc_sensors(const OrcConfig *c, OrcEnv *e, int ag, double dists[8]) {
    const int32_t *det = e->det[ag];
    int hits[ORC_MAX_K] = {0};
    for (int d = 0; d < 8; d++) dists[d] = 0.0;
    if (e->num_obs > 0) {
        for (int d = 0; d < 8; d++) {
            int32_t end[2] = {det[0] + STEP[d][0], det[1] + STEP[d][1]};
            int inter = 0;
            for (int k = 0; k < e->num_obs; k++) {
                const int32_t *r = e->rect[k];
                const int32_t p0[2] = {r[0], r[1]}, p1[2] = {r[0], r[3]}, p2[2] = {r[2], r[3]}, p3[2] = {r[2], r[1]};
                const int32_t *ea[4] = {p0, p0, p2, p2}, *eb[4] = {p1, p3, p1, p3};   /* R:1000-1006 */
                double seg_dist[4] = {0, 0, 0, 0};
                for (int s = 0; s < 4; s++) {
                    if (inter < 2 && orc_seg_touches_seg(ea[s], eb[s], det, end)) {
                        seg_dist[s] = (110.0 - pt_edge_dist(det, ea[s], eb[s])) / 110.0;
                        inter++;
                        hits[k]++;
                    }
                }
                if (inter > 0) {
                    double m = seg_dist[0];
                    for (int s = 1; s < 4; s++) if (seg_dist[s] > m) m = seg_dist[s];
                    if (m > dists[d]) dists[d] = m;
                }
            }
        }
        double ones = 0.0;
        for (int d = 0; d < 8; d++) if (dists[d] == 1.0) ones += 1.0;
        if (ones > 3) {
            /* max(zip(obs_idx_ls, self.poly)): most hits, ties -> lexicographically largest vertex list R:1222-1226 */
            int best = 0;
            for (int k = 1; k < e->num_obs; k++) {
                if (hits[k] > hits[best]) best = k;
                else if (hits[k] == hits[best]) {
                    const int32_t *a = e->rect[k], *b = e->rect[best];
                    /* vertex list (x0,y0),(x0,y1),(x1,y1),(x1,y0) */
                    int32_t ka[8] = {a[0], a[1], a[0], a[3], a[2], a[3], a[2], a[1]};
                    int32_t kb[8] = {b[0], b[1], b[0], b[3], b[2], b[3], b[2], b[1]};
                    int cmp = 0;
                    for (int i = 0; i < 8 && !cmp; i++) cmp = (ka[i] > kb[i]) - (ka[i] < kb[i]);
                    if (cmp > 0) best = k;
                }
            }
            correct_coords(e, det, e->rect[best], dists);
        }
    }
    if (c->enforce) {                                                     /* R:1232-1259 */
        if (det[0] - 110 < c->bbox[0]) {
            if (dists[0] != 0.0) e->status |= ORC_ST_WALL_ASSERT;
            dists[0] = (110.0 - fabs((double)det[0] - c->bbox[0])) / 110.0;
        }
        if (det[1] - 110 < c->bbox[1]) {
            if (dists[6] != 0.0) e->status |= ORC_ST_WALL_ASSERT;
            dists[6] = (110.0 - fabs((double)det[1] - c->bbox[1])) / 110.0;
        }
        if (c->bbox[2] <= det[0] + 110) {
            if (dists[4] != 0.0) e->status |= ORC_ST_WALL_ASSERT;
            dists[4] = (110.0 - fabs((double)c->bbox[2] - det[0])) / 110.0;
        }
        if (c->bbox[3] <= det[1] + 110) {
            if (dists[2] != 0.0) e->status |= ORC_ST_WALL_ASSERT;
            dists[2] = (110.0 - fabs((double)c->bbox[3] - det[1])) / 110.0;
        }
    }
}

static double expected_counts(const OrcConfig *c, OrcEnv *e, int ag) {
    if (e->los_blocked[ag]) return (double)e->bkg;                        /* R:499-501 */
    double d = e->euc[ag];
    if (d == 0.0) { e->status |= ORC_ST_LAMBDA_INF; d = 1.0; }          /* reference: inf -> poisson raises */
    if (c->count_law == 1) return (double)e->intensity / (d * d) + (double)e->bkg;
    return (double)e->intensity / d + (double)e->bkg;
}

/* RadSearch.step R:443-728 (agent_step R:458-613 inlined).  actions == NULL is step(None). */
void orc_step(const OrcConfig *c, OrcEnv *e, const int32_t *actions, OrcRng *rngs, OrcStepOut *out) {
    int A = c->n_agents;
    int32_t proposed[ORC_MAX_A][2];
    int n_prop = 0;
    if (actions) {                                                        /* R:645-648 */
        for (int i = 0; i < A; i++) {
            proposed[i][0] = e->det[i][0] + STEP[actions[i]][0];
            proposed[i][1] = e->det[i][1] + STEP[actions[i]][1];
        }
        n_prop = A;
    }
    double md = max_dist(c);
    int have_max = 0;
    double max_reward = 0.0;
    for (int ag = 0; ag < A; ag++) {
        int action = actions ? actions[ag] : -1;
        double reward, measurement;
        e->oob[ag] = 0; e->collision[ag] = 0;                              /* R:479-480 */
        if (take_action(c, e, ag, action, (const int32_t(*)[2])proposed, n_prop)) {
            e->sp[ag] = orc_shortest_path(e, e->det[ag]);                  /* R:491-493 */
            double dx = (double)e->det[ag][0] - e->src[0], dy = (double)e->det[ag][1] - e->src[1];
            e->euc[ag] = sqrt(dx * dx + dy * dy);                          /* R:494 */
            e->los_blocked[ag] = is_intersect(e, ag);                      /* R:495 */
            out->lam[ag] = expected_counts(c, e, ag);
            measurement = (double)orc_poisson(&rngs[ag], out->lam[ag]);    /* R:498-502 */
            if (e->sp[ag] < 110) { reward = 0.1; e->done = 1; }            /* R:507-510 */
            else if (e->sp[ag] < e->best[ag]) { reward = 0.1; e->best[ag] = e->sp[ag]; }
            else if (action == 8) reward = -1.0 * e->sp[ag] / md;          /* R:516-520 */
            else reward = -0.5 * e->sp[ag] / md;
        } else {
            if (e->iter_count > 0) {                                       /* R:528-549: stale sp/euc */
                e->los_blocked[ag] = is_intersect(e, ag);
            } else {                                                       /* R:551-567 */
                e->sp[ag] = e->best[ag];
                double dx = (double)e->det[ag][0] - e->src[0], dy = (double)e->det[ag][1] - e->src[1];
                e->euc[ag] = sqrt(dx * dx + dy * dy);
                e->los_blocked[ag] = is_intersect(e, ag);
            }
            out->lam[ag] = expected_counts(c, e, ag);
            measurement = (double)orc_poisson(&rngs[ag], out->lam[ag]);
            reward = -0.5 * e->sp[ag] / md;
        }
        double sens[8];
        if (e->num_obs > 0 || c->enforce) orc_sensors(c, e, ag, sens);     /* R:584-588 */
        else for (int d = 0; d < 8; d++) sens[d] = 0.0;
        double inv = 1.0 / (double)sa_y1(c);                               /* 1 / search_area[2][1]  R:577-579 */
        out->obs[ag][0] = measurement;
        out->obs[ag][1] = (double)e->det[ag][0] * inv;
        out->obs[ag][2] = (double)e->det[ag][1] * inv;
        for (int d = 0; d < 8; d++) out->obs[ag][3 + d] = sens[d];
        out->reward[ag] = orc_round2(reward);                              /* R:613 */
        out->done[ag] = e->done;
        /* team reward R:661-665 (`if not max_reward` treats 0.0 like None) */
        if (!have_max || max_reward == 0.0) { max_reward = out->reward[ag]; have_max = 1; }
        else if (max_reward < out->reward[ag]) max_reward = out->reward[ag];
    }
    out->team_reward = have_max ? max_reward : NAN;
    e->iter_count += 1;                                                    /* R:674 */
}

/* ------------------------------------------------------------------------------------------------------------ */
/* reset R:730-797, create_obs R:948-1011, sample_source_loc_pos R:1013-1131                                       */
/* Draws come from the Philox stream (domain 1); numpy's PCG64 sequence is not reproduced (distributional parity). */
/* ------------------------------------------------------------------------------------------------------------ */
static int rects_touch(const int32_t a[4], const int32_t b[4]) {
    /* isclose(boundary_distance(poly_a, poly_b), 0, abs_tol=1e-7) on the lattice: the two boundaries share a point */
    int closed = a[0] <= b[2] && b[0] <= a[2] && a[1] <= b[3] && b[1] <= a[3];
    int a_in_b = b[0] < a[0] && a[2] < b[2] && b[1] < a[1] && a[3] < b[3];
    int b_in_a = a[0] < b[0] && b[2] < a[2] && a[1] < b[1] && b[3] < a[3];
    return closed && !a_in_b && !b_in_a;
}
static int rects_nested(const int32_t a[4], const int32_t b[4]) {
    int a_in_b = b[0] <= a[0] && a[2] <= b[2] && b[1] <= a[1] && a[3] <= b[3];
    int b_in_a = a[0] <= b[0] && b[2] <= a[2] && a[1] <= b[1] && b[3] <= a[3];
    return a_in_b || b_in_a;
}

static void rand_point(const OrcConfig *c, OrcRng *g, int32_t p[2]) {
    uint32_t span = (uint32_t)(sa_x1(c) - sa_x0(c));                       /* R:1030-1036: x range for both */
    p[0] = sa_x0(c) + (int32_t)orc_rng_below(g, span);
    p[1] = sa_x0(c) + (int32_t)orc_rng_below(g, span);
}

static void create_obstructions(const OrcConfig *c, OrcEnv *e, OrcRng *g) {
    int hx = (int)((double)sa_x1(c) * 0.9), hy = (int)((double)sa_y1(c) * 0.9);   /* R:961-966 */
    for (int attempt = 0; attempt < 256; attempt++) {
        if (c->obstruction_count == -1) e->num_obs = 1 + (int)orc_rng_below(g, 5);     /* R:745-750 */
        else e->num_obs = c->obstruction_count;
        int ii = 0, tries = 0;
        while (ii < e->num_obs && tries < 4096) {
            tries++;
            int32_t sx = sa_x0(c) + (int32_t)orc_rng_below(g, (uint32_t)(hx - sa_x0(c)));
            int32_t sy = sa_y0(c) + (int32_t)orc_rng_below(g, (uint32_t)(hy - sa_y0(c)));
            int32_t ex = c->obs_area[0] + (int32_t)orc_rng_below(g, (uint32_t)(c->obs_area[1] - c->obs_area[0]));
            int32_t ey = c->obs_area[0] + (int32_t)orc_rng_below(g, (uint32_t)(c->obs_area[1] - c->obs_area[0]));
            int32_t r[4] = {sx, sy, sx + ex, sy + ey};
            int touch = 0;
            for (int kk = 0; kk < ii && !touch; kk++) touch = rects_touch(e->rect[kk], r);   /* R:985-992 */
            if (!touch) { memcpy(e->rect[ii], r, sizeof(r)); ii++; }
        }
        if (ii < e->num_obs) { e->status |= ORC_ST_REJECT_CAP; e->num_obs = ii; }
        /* world.is_valid R:788-791: a rectangle nested in another one invalidates the world -> everything is redrawn */
        int nested = 0;
        for (int i = 0; i < e->num_obs && !nested; i++)
            for (int j = i + 1; j < e->num_obs && !nested; j++) nested = rects_nested(e->rect[i], e->rect[j]);
        if (!nested) return;
    }
    e->status |= ORC_ST_REJECT_CAP;
}

static void sample_source_loc_pos(const OrcConfig *c, OrcEnv *e, OrcRng *g, int32_t det[2]) {
    int32_t src[2];
    rand_point(c, g, src);
    rand_point(c, g, det);
    int tries = 0;
    for (;;) {                                                             /* R:1057-1076 */
        int inside = 0;
        for (int k = 0; k < e->num_obs && !inside; k++) inside = orc_in_rect_closed(det, e->rect[k]);
        if (!inside) break;
        if (++tries > 100000) { e->status |= ORC_ST_REJECT_CAP; break; }
        rand_point(c, g, det);
    }
    int num_retry = 0;
    tries = 0;
    for (;;) {                                                             /* R:1091-1129 */
        for (;;) {
            int64_t dx = (int64_t)det[0] - src[0], dy = (int64_t)det[1] - src[1];
            if (dx * dx + dy * dy >= 1000000) break;                       /* dist_p < MIN_STARTING_DISTANCE */
            if (++tries > 100000) { e->status |= ORC_ST_REJECT_CAP; break; }
            rand_point(c, g, src);
        }
        int resamp = 0, inter = 0;
        for (int k = 0; k < e->num_obs && !resamp; k++) {
            if (orc_in_rect_closed(src, e->rect[k])) resamp = 1;
            if (!resamp && orc_los_blocked_rect(det, src, e->rect[k])) inter = 1;
        }
        if (e->num_obs == 0 || (num_retry > 20 && !resamp)) break;
        else if (resamp || !inter) { rand_point(c, g, src); num_retry++; }
        else break;
        if (++tries > 100000) { e->status |= ORC_ST_REJECT_CAP; break; }
    }
    e->src[0] = src[0]; e->src[1] = src[1];
}

static void finish_reset(const OrcConfig *c, OrcEnv *e, const int32_t det[2], uint64_t seed, uint32_t env_id,
                         const double *inj_u, int32_t n_inj, OrcStepOut *out) {
    const uint64_t step_ctr = (uint64_t)e->episode;
    for (int ag = 0; ag < c->n_agents; ag++) {
        e->det[ag][0] = det[0]; e->det[ag][1] = det[1];
        e->oob[ag] = 0; e->oob_count[ag] = 0; e->blocked[ag] = 0; e->collision[ag] = 0;   /* Agent.reset R:289-300 */
        e->best[ag] = orc_shortest_path(e, det);                           /* R:771-776 */
    }
    e->done = 0; e->iter_count = 0; e->ep_len = 0;                         /* R:739-740 */
    OrcRng rngs[ORC_MAX_A];
    for (int ag = 0; ag < c->n_agents; ag++) {
        if (inj_u) orc_rng_inject(&rngs[ag], inj_u + (size_t)ag * n_inj, n_inj, &e->status);
        else orc_rng_philox(&rngs[ag], seed, env_id, 2, (uint32_t)ag, step_ctr, &e->status);
    }
    OrcStepOut tmp;
    orc_step(c, e, NULL, rngs, out ? out : &tmp);                          /* R:794 */
    e->iter_count = 0;                                                     /* R:796 */
}

void orc_reset(const OrcConfig *c, OrcEnv *e, int new_obstacles, uint64_t seed, uint32_t env_id,
               const double *inj_u, int32_t n_inj, OrcStepOut *out) {
    OrcRng g;
    e->episode += 1;
    orc_rng_philox(&g, seed, env_id, 1, 0, (uint64_t)e->episode, &e->status);
    if (new_obstacles) create_obstructions(c, e, &g);                      /* R:744-762 */
    int32_t det[2];
    sample_source_loc_pos(c, e, &g, det);                                  /* R:764-769 */
    e->intensity = 1000000 + (int32_t)orc_rng_below(&g, 9000000);          /* R:778 integers(1e6, 10e6) */
    e->bkg = 10 + (int32_t)orc_rng_below(&g, 41);                          /* R:779 integers(10, 51)    */
    finish_reset(c, e, det, seed, env_id, inj_u, n_inj, out);
}

/* refresh_environment R:799-874 (scenario injection); prev_det_dist is set after the probe step, as in R:864-868 */
void orc_load_scenario(const OrcConfig *c, OrcEnv *e, const int32_t src[2], const int32_t det[2], int32_t intensity,
                       int32_t bkg, const int32_t *rects, int32_t num_obs) {
    const uint32_t episode = e->episode;
    memset(e, 0, sizeof(*e));
    e->episode = episode;
    e->num_obs = num_obs;
    for (int k = 0; k < num_obs; k++) memcpy(e->rect[k], rects + 4 * k, 4 * sizeof(int32_t));
    e->src[0] = src[0]; e->src[1] = src[1];
    e->intensity = intensity; e->bkg = bkg;
    for (int ag = 0; ag < c->n_agents; ag++) {
        e->det[ag][0] = det[0]; e->det[ag][1] = det[1];
        e->best[ag] = orc_shortest_path(e, det);
        e->sp[ag] = e->best[ag];
        double dx = (double)det[0] - src[0], dy = (double)det[1] - src[1];
        e->euc[ag] = sqrt(dx * dx + dy * dy);
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* batched drivers                                                                                               */
/* ------------------------------------------------------------------------------------------------------------ */
static void set_threads(int32_t threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
}

void orc_step_batch(const OrcConfig *c, OrcEnv *envs, int32_t n, const int32_t *actions, uint64_t seed,
                    uint32_t env_id0, uint64_t step_ctr, const double *inj_u, int32_t n_inj, OrcStepOut *outs,
                    int32_t threads) {
    set_threads(threads);
    int A = c->n_agents;
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; i++) {
        OrcRng rngs[ORC_MAX_A];
        for (int ag = 0; ag < A; ag++) {
            if (inj_u) orc_rng_inject(&rngs[ag], inj_u + ((size_t)i * A + ag) * n_inj, n_inj, &envs[i].status);
            else orc_rng_philox(&rngs[ag], seed, env_id0 + (uint32_t)i, 0, (uint32_t)ag, step_ctr, &envs[i].status);
        }
        orc_step(c, &envs[i], actions ? actions + (size_t)i * A : NULL, rngs, &outs[i]);
        envs[i].ep_len += 1;
    }
}

void orc_reset_batch(const OrcConfig *c, OrcEnv *envs, int32_t n, const uint8_t *mask, const uint8_t *new_obs_mask,
                     uint64_t seed, uint32_t env_id0, const double *inj_u, int32_t n_inj, OrcStepOut *outs,
                     int32_t threads) {
    set_threads(threads);
    int A = c->n_agents;
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t i = 0; i < n; i++) {
        if (mask && !mask[i]) continue;
        orc_reset(c, &envs[i], new_obs_mask ? new_obs_mask[i] : 1, seed, env_id0 + (uint32_t)i,
                  inj_u ? inj_u + (size_t)i * A * n_inj : NULL, n_inj, &outs[i]);
    }
}

/* the per-env Python loop of train.py:321-549 / test_environment/ppo.py:495-573 without the policy: uniform random
 * actions 0..7 (Philox domain 3), timeout at max_ep_len, reset on done/timeout, new obstructions at the epoch end. */
int64_t orc_rollout(const OrcConfig *c, OrcEnv *envs, int32_t n, int32_t T, uint64_t seed, uint32_t env_id0,
                    uint64_t step_ctr0, int32_t epoch_end_last, int32_t threads, double *checksum) {
    set_threads(threads);
    int A = c->n_agents;
    double total = 0.0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int32_t i = 0; i < n; i++) {
        OrcEnv *e = &envs[i];
        OrcStepOut out;
        double acc = 0.0;
        for (int32_t t = 0; t < T; t++) {
            uint64_t ctr = step_ctr0 + (uint64_t)t;
            OrcRng rngs[ORC_MAX_A], ga;
            int32_t act[ORC_MAX_A];
            orc_rng_philox(&ga, seed, env_id0 + (uint32_t)i, 3, 0, ctr, &e->status);
            for (int ag = 0; ag < A; ag++) {
                act[ag] = (int32_t)(orc_rng_u32(&ga) & 7u);
                orc_rng_philox(&rngs[ag], seed, env_id0 + (uint32_t)i, 0, (uint32_t)ag, ctr, &e->status);
            }
            orc_step(c, e, act, rngs, &out);
            e->ep_len += 1;
            acc += out.obs[0][0] + out.reward[0];
            int timeout = e->ep_len == c->max_ep_len;                      /* train.py:394-405 */
            int over = e->done || timeout;
            int epoch_ended = epoch_end_last && t == T - 1;
            if (over || epoch_ended) orc_reset(c, e, epoch_ended, seed, env_id0 + (uint32_t)i, NULL, 0, &out);
        }
        total += acc;
    }
    if (checksum) *checksum = total;
    return (int64_t)n * T;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* GAE-lambda + rewards-to-go, P:391-423 with discount_cumsum P:62-85 (scipy.signal.lfilter([1],[1,-d]) on the      */
/* reversed vector == y[t] = x[t] + d*y[t+1] in float64), batched over columns as SURVEY Appendix C.               */
/* ------------------------------------------------------------------------------------------------------------ */
void orc_gae(const float *rew, const float *val, const uint8_t *path_end, const float *boot, float *adv, float *ret,
             int32_t T, int32_t N, double gamma, double lam, int32_t threads) {
    set_threads(threads);
    const double gl = gamma * lam;
#pragma omp parallel for schedule(static)
    for (int32_t n = 0; n < N; n++) {
        double nv = 0.0, na = 0.0, nr = 0.0;
        for (int32_t t = T - 1; t >= 0; t--) {
            size_t i = (size_t)t * N + n;
            if (path_end[i] || t == T - 1) { nv = (double)boot[i]; na = 0.0; nr = (double)boot[i]; }
            double r = (double)rew[i], v = (double)val[i];
            double delta = r + gamma * nv - v;                             /* P:415 */
            double a = (t == T - 1 || path_end[i]) ? delta : delta + gl * na;   /* lfilter start: y = x */
            double g = r + gamma * nr;                                     /* P:420 */
            adv[i] = (float)a;
            ret[i] = (float)g;
            nv = v; na = a; nr = g;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Per-episode running standardisation of the count channel.                                                       */
/* mode 1: StatisticStandardization.update then .standardize (algos/multiagent/NeuralNetworkCores/RADTEAM_core.py   */
/*         :215-265): Welford mean / M2, sample variance M2/(count-1), std = max(sqrt(var), 1).                      */
/* mode 2: StatBuff.update (algos/test_environment/core.py:62-73): same recurrence, sig_obs = sqrt(var) with 0 -> 1, */
/*         and the caller's np.clip((o - mu)/sig_obs, -8, 8) (algos/test_environment/ppo.py:502).                    */
/* s = {mean, M2, std, count}; returns the z-score of x after the update.                                            */
/* ------------------------------------------------------------------------------------------------------------ */
double orc_stat_update_standardize(OrcStat *s, double x, int32_t mode) {
    s->count += 1;
    if (s->count == 1) {
        s->mean = x;                                                       /* :236-238 (M2, std keep their defaults) */
    } else {
        double mean_new = s->mean + (x - s->mean) / (double)s->count;      /* :240 */
        double m2_new = s->m2 + (x - s->mean) * (x - mean_new);            /* :241-243 */
        s->mean = mean_new;
        s->m2 = m2_new;
        double sd = sqrt(m2_new / (double)(s->count - 1));                 /* :246-247 */
        if (mode == 2) s->std = (sd == 0.0) ? 1.0 : sd;                    /* core.py:70-72 */
        else s->std = sd > 1.0 ? sd : 1.0;                                 /* :247 max(sqrt(var), 1) */
    }
    double z = (x - s->mean) / s->std;                                     /* :265 */
    if (mode == 2) z = z < -8.0 ? -8.0 : (z > 8.0 ? 8.0 : z);
    return z;
}
void orc_stat_reset(OrcStat *s) { s->mean = 0.0; s->m2 = 0.0; s->std = 1.0; s->count = 0; }   /* :274-276 */

int32_t orc_sizeof_env(void) { return (int32_t)sizeof(OrcEnv); }
int32_t orc_sizeof_out(void) { return (int32_t)sizeof(OrcStepOut); }
