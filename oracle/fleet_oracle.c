/*
 * fleet_oracle.c — CPU restatement of the FleetRL environment step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (fleetrl_b200/) never calls into it.
 *
 * It restates, in scalar C with the reference's per-env / per-car loop structure, float64 arithmetic and
 * operation order (compile with -ffp-contract=off):
 *   FleetEnv.reset / step                    fleetrl/fleet_env/fleet_environment.py:330-434, 436-702
 *   EvCharger.charge                         fleetrl/utils/ev_charging/ev_charger.py:39-231
 *   LoadCalculation.check_violation          fleetrl/utils/load_calculation/load_calculation.py:83-94
 *   ScoreConfig penalties                    fleetrl/fleet_env/config/score_config.py:26-41
 *   Observer*.get_obs                        fleetrl/utils/observation/observer_bl_pv.py:12-136 (+ siblings)
 *   Unit/OracleNormalization.normalize_obs   fleetrl/utils/normalization/unit_normalization.py:15-43,
 *                                            oracle_normalization.py:56-162
 *   RainflowSeiDegradation                   fleetrl/utils/battery_degradation/rainflow_sei_degradation.py:37-212
 *   EmpiricalDegradation                     fleetrl/utils/battery_degradation/empirical_degradation.py:29-99
 *   rainflow 3.2.0 reversals/extract_cycles  third-party, PyPI `rainflow==3.2.0` (requirements.txt), not vendored:
 *                                            restated from the published ASTM E1049-85 three-point algorithm.
 *
 * Pinning: tests/test_oracle_golden.py checks this file against trajectories produced by the UNMODIFIED
 * reference package run in the build container (oracle/gen_golden.py -> tests/golden/ npz files) and the rainflow
 * part against the ASTM E1049-85 example of the rainflow README (tests/test_rainflow_kat.py).
 *
 * Numerical note: the reference is pinned to NumPy 1.26, where `python_float * np.float32` is float64, so the
 * float32 action is widened once and everything else is float64.  That is what is restated here.
 *
 * Batch semantics added on top of the reference (they mirror what SB3's SubprocVecEnv worker does around a
 * FleetEnv): E independent envs, optional auto-reset on done with the terminal observation kept aside.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/fleetstep.h"

#ifndef M_E
#define M_E 2.718281828459045
#endif

typedef struct OracleEnv {
    int32_t t, t_start, t_fin, ep_count, done_sticky;
    double *soc, *hl, *soc_deg, *soh, *cap, *target;      /* [N] */
    double *rf_len, *fd_cyc, *life, *sei_soh;             /* [N] RainflowSeiDegradation members           */
    double *last_deg;                                      /* [N]                                          */
    double *charge_log;                                    /* [N] EvCharger's charge_log of the last step, ev_charger.py:212 */
    int32_t *n_cycles;                                     /* [N] len(rainflow_result) at last evaluation  */
    double *hist;                                          /* [hist_cap][N] LogDataDeg.soc_log             */
    int32_t hist_len, hist_cap;
    double ep_return, last_ep_return, last_reward, last_cashflow, last_overload, last_soc_viol;
} OracleEnv;

typedef struct Oracle {
    FleetConsts c;
    FleetTables tb;        /* deep copies */
    int32_t E, D;
    int64_t env_id_offset;
    OracleEnv* envs;
    const int32_t* next_start; /* optional injected start indices for auto-reset */
    double stats[FLEET_S__COUNT];
    uint32_t err_flags;
    struct Pool* pool;     /* persistent worker threads of oracle_step_mt (CPU-baseline timing) */
} Oracle;

/* ------------------------------------------------------------------------------------------------ helpers */

static void* dupmem(const void* src, size_t bytes) {
    if (!src) return NULL;
    void* p = malloc(bytes);
    memcpy(p, src, bytes);
    return p;
}

/* detect_dim_and_bounds, fleet_environment.py:854-949 */
static int32_t obs_dim(const FleetConsts* c) {
    int32_t N = c->num_evs, dim = 2 * N;
    if (c->include_price) dim += 2 * (c->price_lookahead + 1);
    if (c->include_price && c->include_building) dim += c->bl_pv_lookahead + 1;
    if (c->include_price && c->include_pv) dim += c->bl_pv_lookahead + 1;
    if (c->aux) {
        dim += 5 * N + 1 + 6;
        if (c->include_price && c->include_building) dim += 3;
    }
    return dim;
}

/* Counter-based start-index draw used on auto-reset when no start index is injected (integer work: the
 * product must match it bit for bit).  splitmix64 finaliser over (seed, global env id, episode number). */
static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static int32_t draw_start(const FleetConsts* c, int64_t env_id, int32_t episode_no) {
    uint64_t h = mix64(c->seed ^ mix64((uint64_t)env_id * 0xD1B54A32D192ED03ull + (uint64_t)(uint32_t)episode_no));
    uint64_t span = (uint64_t)(c->start_hi - c->start_lo) + 1ull;
    return c->start_lo + (int32_t)(h % span);
}

/* ------------------------------------------------------------------------------------------- observation */

/* Index of look-ahead element k at time index t: element 0 is the row at t itself, element k >= 1 the first
 * row of the k-th following clock hour (resample("H").first() on the slice starting at t,
 * observer_bl_pv.py:50-80).  Clamped to the table end (the reference's arrays would come out short there). */
static int32_t look_idx(const Oracle* o, int32_t t, int32_t k) {
    int32_t sph = o->c.steps_per_hour;
    int32_t pos = (int32_t)o->tb.minute[t] * sph / 60;
    int32_t i = (k == 0) ? t : t - pos + k * sph;
    if (i > o->c.table_len - 1) i = o->c.table_len - 1;
    return i;
}

/* Observer*.get_obs + normalizer.normalize_obs for time index t with the simulated soc / hours_left already
 * substituted (fleet_environment.py:406-414, 645-652).  `target` is FleetEnv.target_soc. */
static void build_obs(const Oracle* o, int32_t t, const double* soc, const double* hl, const double* target,
                      float* out) {
    const FleetConsts* c = &o->c;
    const FleetTables* tb = &o->tb;
    const int32_t N = c->num_evs, T = c->table_len;
    const int norm = c->normalize;
    int32_t p = 0;

    for (int n = 0; n < N; n++) out[p++] = (float)soc[n];                      /* soc: "already normalized" :65 */
    for (int n = 0; n < N; n++) out[p++] = (float)(norm ? hl[n] / c->max_time_left : hl[n]);   /* :66 */

    if (c->include_price) {
        for (int k = 0; k <= c->price_lookahead; k++) {                         /* observer_bl_pv.py:63 */
            double v = (tb->delu[look_idx(o, t, k)] + c->fixed_markup) * c->variable_multiplier;
            if (norm) v = (v - c->min_price) / (c->max_price - c->min_price);   /* oracle_normalization.py:70 */
            out[p++] = (float)v;
        }
        for (int k = 0; k <= c->price_lookahead; k++) {                         /* observer_bl_pv.py:64 */
            double v = tb->tariff[look_idx(o, t, k)] * (1 - c->feed_in_deduction);
            if (norm) v = (v - c->min_tariff) / (c->max_tariff - c->min_tariff);
            out[p++] = (float)v;
        }
        if (c->include_building)
            for (int k = 0; k <= c->bl_pv_lookahead; k++) {
                double v = tb->load[look_idx(o, t, k)];
                if (norm) v = v / c->max_building;                              /* :101,138 */
                out[p++] = (float)v;
            }
        if (c->include_pv)
            for (int k = 0; k <= c->bl_pv_lookahead; k++) {
                double v = tb->pv[look_idx(o, t, k)];
                if (norm) v = v / c->max_pv;                                    /* :139 */
                out[p++] = (float)v;
            }
    }
    if (!c->aux) return;

    /* auxiliary block, observer_bl_pv.py:85-107: computed from the SCHEDULE columns at t, not from the
     * simulated state (SURVEY B-4). */
    const double max_soc = c->target_soc;                                                   /* :49 */
    const double max_hours_needed = (c->target_soc * c->init_battery_cap) / (c->evse_max_power * c->charging_eff); /* :50 */
    float* o_there = out + p;
    float* o_tgt = o_there + N;
    float* o_cl = o_tgt + N;
    float* o_hn = o_cl + N;
    float* o_lax = o_hn + N;
    for (int n = 0; n < N; n++) {
        double th = (double)tb->there[(size_t)n * T + t];
        double s_ret = tb->soc_on_return[(size_t)n * T + t];
        double t_left = tb->time_left[(size_t)n * T + t];
        double tt = target[n] * th;                                                          /* :86 */
        double cl = tt - s_ret;                                                              /* :88 */
        double hn = cl * c->lc_batt_cap / (c->evse_max_power * c->charging_eff);             /* :89 */
        double lax = (t_left / (hn + 0.001) - 1) * th;                                       /* :90 */
        lax = lax < 0 ? 0 : (lax > 5 ? 5 : lax);                                             /* :91 */
        if (norm) {                                                                          /* oracle_normalization.py:146-151 */
            th = th / 1; tt = tt / max_soc; cl = cl / max_soc; hn = hn / max_hours_needed; lax = lax / 5;
        }
        o_there[n] = (float)th; o_tgt[n] = (float)tt; o_cl[n] = (float)cl; o_hn[n] = (float)hn; o_lax[n] = (float)lax;
    }
    p += 5 * N;
    double evse = c->evse_max_power;
    out[p++] = (float)(norm ? evse / c->evse_max_power : evse);                              /* :93, norm :151 */
    if (c->include_price && c->include_building) {
        double grid = c->grid_connection;                                                    /* :95 */
        double avail = grid - tb->load[t];                                                   /* :96 / bl-only :88 */
        if (c->include_pv) avail = avail + tb->pv[t];
        double poss = avail / ((double)N * evse);                                            /* :98 */
        if (poss > 1) poss = 1;
        if (norm) { grid = grid / c->grid_connection; avail = avail / c->grid_connection; poss = poss / 1; }
        out[p++] = (float)grid; out[p++] = (float)avail; out[p++] = (float)poss;
    }
    for (int k = 0; k < 6; k++) out[p++] = (float)tb->cal_sincos[(size_t)t * 6 + k];          /* :100-107 */
}

/* ------------------------------------------------------------------------------------------ degradation */

/* rainflow.extract_cycles over x[0..n) with stride (see oracle/refshim/rainflow.py for the Python restatement).
 * Emits cycles in generation order through cb.  Returns the number of cycles. */
typedef struct { double range, mean, count; int32_t i_start, i_end; } Cycle;

typedef struct { double* v; int32_t* i; int32_t lo, hi; } Deque; /* points[lo..hi) */

static int32_t rainflow_cycles(const double* x, int32_t n, int32_t stride, Cycle* out /* cap >= n */) {
    int32_t ncyc = 0;
    if (n < 2) return 0;
    Deque q;
    q.v = (double*)malloc(sizeof(double) * (size_t)(n + 2));
    q.i = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 2));
    q.lo = q.hi = 0;

#define EMIT(a, b, cnt)                                                                   \
    do {                                                                                  \
        out[ncyc].range = fabs(q.v[a] - q.v[b]); out[ncyc].mean = 0.5 * (q.v[a] + q.v[b]); \
        out[ncyc].count = (cnt); out[ncyc].i_start = q.i[a]; out[ncyc].i_end = q.i[b]; ncyc++; \
    } while (0)
#define PUSH(idx, val)                                                                    \
    do {                                                                                  \
        q.v[q.hi] = (val); q.i[q.hi] = (idx); q.hi++;                                     \
        while (q.hi - q.lo >= 3) {                                                        \
            double X = fabs(q.v[q.hi - 1] - q.v[q.hi - 2]);                               \
            double Y = fabs(q.v[q.hi - 2] - q.v[q.hi - 3]);                               \
            if (X < Y) break;                                                             \
            if (q.hi - q.lo == 3) { EMIT(q.lo, q.lo + 1, 0.5); q.lo++; }                  \
            else { EMIT(q.hi - 3, q.hi - 2, 1.0);                                         \
                   q.v[q.hi - 3] = q.v[q.hi - 1]; q.i[q.hi - 3] = q.i[q.hi - 1]; q.hi -= 2; } \
        }                                                                                 \
    } while (0)

    /* reversals() */
    double x_last = x[0], xc = x[(size_t)stride];
    double d_last = xc - x_last;
    PUSH(0, x_last);
    int32_t index = -1;
    double x_next = 0;
    for (int32_t p = 2; p < n; p++) {
        index = p - 1;
        x_next = x[(size_t)p * stride];
        if (x_next == xc) continue;
        double d_next = x_next - xc;
        if (d_last * d_next < 0) PUSH(index, xc);
        x_last = xc; xc = x_next; d_last = d_next;
    }
    if (index >= 0) PUSH(index + 1, x_next);
    while (q.hi - q.lo > 1) { EMIT(q.lo, q.lo + 1, 0.5); q.lo++; }
#undef EMIT
#undef PUSH
    free(q.v); free(q.i);
    return ncyc;
}

/* Exported for the known-answer test: cycles as 5 doubles each (range, mean, count, i_start, i_end). */
int32_t oracle_rainflow(const double* x, int32_t n, double* out5 /* [n][5] */) {
    Cycle* cyc = (Cycle*)malloc(sizeof(Cycle) * (size_t)(n + 2));
    int32_t m = rainflow_cycles(x, n, 1, cyc);
    for (int32_t i = 0; i < m; i++) {
        out5[i * 5 + 0] = cyc[i].range; out5[i * 5 + 1] = cyc[i].mean; out5[i * 5 + 2] = cyc[i].count;
        out5[i * 5 + 3] = cyc[i].i_start; out5[i * 5 + 4] = cyc[i].i_end;
    }
    free(cyc);
    return m;
}

/* RainflowSeiDegradation.calculate_degradation for one env; deg[N] out. */
static void sei_degradation(Oracle* o, OracleEnv* e, double* deg) {
    const FleetConsts* c = &o->c;
    const int32_t N = c->num_evs, n = e->hist_len;
    const double alpha_sei = 5.75E-2, beta_sei = 121, kd1 = 1.4E5, kd2 = -5.01E-1, kd3 = -1.23E5;
    const double k_sigma = 1.04, sigma_ref = 0.5, k_temp = 6.93E-2, temp_ref = 25, k_dt = 4.14E-10;
    const double temp = c->temperature;
    const double s_temp = pow(M_E, k_temp * (temp - temp_ref) * ((temp_ref + 273.15) / (temp + 273.15)));  /* :72-73 */
    Cycle* cyc = (Cycle*)malloc(sizeof(Cycle) * (size_t)(n + 2));
    for (int i = 0; i < N; i++) {
        int32_t m = rainflow_cycles(e->hist + i, n, N, cyc);                                  /* :132 */
        e->n_cycles[i] = m;
        deg[i] = 0;
        if ((double)m > e->rf_len[i]) {                                                        /* :143 */
            int32_t max_end = 0; double mean_sum = 0;
            for (int j = 0; j < m; j++) { if (cyc[j].i_end > max_end) max_end = cyc[j].i_end; mean_sum += cyc[j].mean; }
            double battery_age = (double)max_end * c->dt * 3600;                               /* :138 */
            double mean_soc_cal = mean_sum / (double)m;                                        /* :140 */
            int32_t a = (int32_t)(e->rf_len[i] - 1), b = m - 1;                                /* :146 */
            double fsum = 0; double max_dod = 0;
            for (int j = a; j < b; j++) {
                double dod = cyc[j].range;
                if (dod > max_dod) max_dod = dod;
                double eff = dod * cyc[j].count;                                               /* :170 */
                eff = eff < 0 ? 0 : (eff > 1 ? 1 : eff);
                double s_dod = 1.0 / (kd1 * pow(eff, kd2) + kd3);                              /* :68  (x ** -1) */
                double s_soc = pow(M_E, k_sigma * (cyc[j].mean - sigma_ref));                  /* :70 */
                fsum += s_dod * s_soc * s_temp;                                                /* :77-79, np.sum :174 */
            }
            if (max_dod > 5) o->err_flags |= 4u;                                               /* :164-167 */
            double fd_cal = (k_dt * battery_age) * pow(M_E, k_sigma * (mean_soc_cal - sigma_ref)) * s_temp; /* :81-83 */
            double new_l;
            e->fd_cyc[i] += fsum;                                                              /* :174 / :184 */
            double fd = e->fd_cyc[i] + fd_cal;
            if (c->init_soh == 1.0) {
                new_l = 1 - alpha_sei * pow(M_E, -beta_sei * fd) - (1 - alpha_sei) * pow(M_E, -fd);  /* :85-86 */
                if (new_l < 0) o->err_flags |= 2u;                                             /* :179-180 */
            } else {
                new_l = 1 - (1 - e->life[i]) * pow(M_E, -fd);                                  /* :89,186 */
            }
            deg[i] = new_l - e->life[i];                                                       /* :189 */
            e->life[i] = new_l;                                                                /* :192 */
            e->rf_len[i] = (double)m;                                                          /* :195 */
        }
        e->sei_soh[i] -= deg[i];                                                               /* :206 */
        e->last_deg[i] = deg[i];
    }
    free(cyc);
}

/* EmpiricalDegradation.calculate_degradation for one env (last two history entries only). */
static void empirical_degradation(Oracle* o, OracleEnv* e, double* deg) {
    const FleetConsts* c = &o->c;
    const int32_t N = c->num_evs;
    const double cal_soc[3] = {0, 40, 90};
    const double cal_aging[3] = {0.0065, 0.0293, 0.065};
    for (int i = 0; i < N; i++) {
        double old_soc = e->hist[(size_t)(e->hist_len - 2) * N + i];                           /* :63 */
        double new_soc = e->hist[(size_t)(e->hist_len - 1) * N + i];                           /* :64 */
        double avg_soc = (old_soc + new_soc) / 2;                                              /* :67 */
        int best = 0; double bd = fabs(cal_soc[0] - avg_soc);                                  /* :70-72 argmin, first wins */
        for (int k = 1; k < 3; k++) { double d = fabs(cal_soc[k] - avg_soc); if (d < bd) { bd = d; best = k; } }
        double cal = cal_aging[best] * c->dt / 8760;                                           /* :75-79 */
        double dod = fabs(new_soc - old_soc);                                                  /* :85 */
        double cyc = (c->evse_max_power <= 22.0) ? dod * 0.000125 / 2 : dod * 0.000167 / 2;     /* :88-91 */
        deg[i] = cal + cyc;                                                                    /* :94 */
        e->last_deg[i] = deg[i];
        e->n_cycles[i] = 0;
    }
}

/* --------------------------------------------------------------------------------------------- reset/step */

static void hist_append(Oracle* o, OracleEnv* e) {
    const int32_t N = o->c.num_evs;
    if (e->hist_len == e->hist_cap) {
        e->hist_cap = e->hist_cap * 2 + 8;
        e->hist = (double*)realloc(e->hist, sizeof(double) * (size_t)e->hist_cap * N);
    }
    memcpy(e->hist + (size_t)e->hist_len * N, e->soc_deg, sizeof(double) * N);                 /* log_data_deg.py:14-15 */
    e->hist_len++;
}

static void env_reset(Oracle* o, OracleEnv* e, int32_t t0, float* obs) {
    const FleetConsts* c = &o->c;
    const FleetTables* tb = &o->tb;
    const int32_t N = c->num_evs, T = c->table_len;
    e->hist_len = 0;                                                                           /* :338-339 */
    e->done_sticky = 0;                                                                        /* :342 */
    if (!c->carry_degradation_state) {           /* fresh-object semantics (documented deviation, SURVEY B-3) */
        for (int n = 0; n < N; n++) {
            e->rf_len[n] = 1; e->fd_cyc[n] = 0; e->life[n] = 1 - c->init_soh; e->sei_soh[n] = c->init_soh;
            e->target[n] = c->target_soc;
        }
    }
    for (int n = 0; n < N; n++) { e->soh[n] = 1.0 * c->init_soh; e->cap[n] = e->soh[n] * c->init_battery_cap; } /* :345-348 */
    e->t_start = t0; e->t = t0; e->t_fin = t0 + c->episode_steps;                              /* :351-358 */
    for (int n = 0; n < N; n++) {                                                              /* :371-372 */
        e->soc[n] = tb->soc_on_return[(size_t)n * T + t0];
        e->hl[n] = tb->time_left[(size_t)n * T + t0];
    }
    for (int n = 0; n < N; n++) {                                                              /* :382-392 */
        double p_avail = fmin(c->obc_max_power, c->evse_max_power);
        double time_needed = (e->target[n] - e->soc[n]) * e->cap[n] / p_avail;
        if (e->hl[n] > 0 && c->min_laxity * time_needed > e->hl[n])
            e->soc[n] = e->target[n] - (time_needed * p_avail / e->cap[n]) / c->min_laxity;
    }
    for (int n = 0; n < N; n++) e->soc_deg[n] = (e->soc[n] == 0) ? c->def_soc : e->soc[n];      /* :395-399 */
    e->ep_return = 0;                                                                          /* :402-403 */
    e->ep_count++;
    if (obs) build_obs(o, e->t, e->soc, e->hl, e->target, obs);                                /* :406-414 */
    if (c->calc_degradation) hist_append(o, e);                                                /* :417-418 */
}

typedef struct StepOut { double reward, cashflow, overload, soc_viol, penalty, degradation; int32_t n_viol, done; } StepOut;

static void env_step(Oracle* o, OracleEnv* e, const float* act, float* obs, StepOut* so, double* deg_scratch,
                     double* next_soc) {
    const FleetConsts* c = &o->c;
    const FleetTables* tb = &o->tb;
    const int32_t N = c->num_evs, T = c->table_len;
    const int32_t t = e->t;
    const double dt = c->dt;

    /* ---- EvCharger.charge, ev_charger.py:69-231 ---- */
    double charging_cost = 0, discharging_revenue = 0, invalid_action_penalty = 0, overcharging_penalty = 0;
    double charging_reward = 0.0, discharging_reward = 0.0;
    const double spot_offset = c->fixed_markup / 1000;                                         /* :35 */
    double connected = 0;                                                                      /* :138 */
    for (int n = 0; n < N; n++) connected += (double)tb->there[(size_t)n * T + t];
    connected = connected > 1 ? connected : 1;                                                 /* :140 */
    for (int n = 0; n < N; n++) {
        const int there = tb->there[(size_t)n * T + t];                                        /* :92 */
        const double a = (double)act[n];
        if (a != a) o->err_flags |= 1u;                                                        /* NaN: TypeError :209 */
        const double possible_power = fmin(c->obc_max_power, c->evse_max_power);              /* :95 */
        if (a >= 0) {
            double demand = (e->target[n] - e->soc[n]) * e->cap[n];                            /* :100 */
            double demanded = possible_power * a * dt;                                         /* :101 */
            if (demanded * c->charging_eff > demand) {                                         /* :104 */
                double pen = c->penalty_overcharging * ((demanded - demand) * (demanded - demand)); /* :105 */
                pen = pen > c->clip_overcharging ? pen : c->clip_overcharging;                 /* :106 */
                overcharging_penalty += pen;
            }
            double energy;
            if (there == 1) energy = fmin(demand / c->charging_eff, demanded);                 /* :114 */
            else {
                energy = 0;
                if (fabs(a) > 0.05) invalid_action_penalty += c->penalty_invalid_action * (a * a);  /* :120-122 */
            }
            next_soc[n] = e->soc[n] + energy * c->charging_eff / e->cap[n];                    /* :128 */
            e->charge_log[n] = energy;                                                         /* :212 */
            double pv_energy = tb->pv ? tb->pv[t] * dt : 0.0;                                  /* :133-136 */
            double grid_energy = energy - (pv_energy / connected);                             /* :142 */
            grid_energy = grid_energy > 0 ? grid_energy : 0;
            double spot = tb->delu[t] / 1000.0;                                                /* :145 */
            charging_cost += (grid_energy * (spot + spot_offset) * c->variable_multiplier);    /* :149 */
            charging_reward += (-1 * c->price_multiplier * tb->price_reward_curve[t] / 1000 * grid_energy); /* :154-156 */
        } else if (a < 0) {
            double left = -1 * e->soc[n] * e->cap[n];                                          /* :161 */
            double demanded = possible_power * a * dt;                                         /* :162 */
            if (demanded * c->discharging_eff < left && there != 0) {                          /* :165 */
                double pen = c->penalty_overcharging * ((left - demanded) * (left - demanded)); /* :166 */
                overcharging_penalty += pen;
            }
            double energy;
            if (there == 1) energy = fmax(left, demanded);                                     /* :174 */
            else {
                energy = 0.0;
                if (fabs(a) > 0.05) invalid_action_penalty += c->penalty_invalid_action * (a * a);  /* :180-182 */
            }
            next_soc[n] = e->soc[n] + energy / e->cap[n];                                      /* :189 */
            e->charge_log[n] = energy;                                                         /* :212 */
            discharging_revenue += (-1 * energy * c->discharging_eff * tb->tariff[t] / 1000
                                    * (1 - c->feed_in_deduction));                             /* :196-199 */
            discharging_reward += (-1 * c->price_multiplier * tb->tariff_reward_curve[t] / 1000 * energy); /* :204-206 */
        } else {
            next_soc[n] = e->soc[n];   /* NaN action: the reference raises; flagged above */
            e->charge_log[n] = 0;
        }
    }
    double cashflow = -1 * charging_cost + discharging_revenue;                                /* :225 */
    double reward = charging_reward + discharging_reward + invalid_action_penalty + overcharging_penalty; /* :228 */
    for (int n = 0; n < N; n++) e->soc[n] = next_soc[n];                                       /* fleet_environment.py:470 */

    /* ---- overload, fleet_environment.py:480-502 + load_calculation.py:93 + score_config.py:33-41 ---- */
    double current_load = (c->include_building && tb->load) ? tb->load[t] : 0;
    double current_pv = (c->include_pv && tb->pv) ? tb->pv[t] : 0;
    double sum_a = 0;                                                                          /* python sum(): 0 + ... */
    for (int n = 0; n < N; n++) sum_a += (double)act[n] * (double)tb->there[(size_t)n * T + t];   /* :491 */
    double margin = c->grid_connection - current_load - sum_a * c->evse_max_power + current_pv;
    double overload = fabs(margin < 0.0 ? margin : 0.0);
    if (overload > 0) {
        double rel = overload / c->grid_connection + 1;                                        /* :496 */
        double pen = (rel < 1.1) ? 0.0 : -700 / (1 + exp(-15.77350877 * (rel - 1.33298382)));
        reward += pen * c->penalty_overloading;                                                /* :501 */
    }

    /* ---- advance time, departures / arrivals, fleet_environment.py:508-623 ---- */
    e->t = t + 1;
    const int32_t tn = e->t;
    const int hour = tb->hour[tn];
    double cum_soc_missing = 0; int32_t n_viol = 0;
    for (int n = 0; n < N; n++) {
        const double ntl = tb->time_left[(size_t)n * T + tn];
        const double nsr = tb->soc_on_return[(size_t)n * T + tn];
        if (e->hl[n] != 0 && ntl == 0) {                                                       /* :531 departure */
            double tg = (c->is_caretaker && hour > 11 && hour < 15) ? c->target_soc_lunch : e->target[n]; /* :536-557 */
            if (tg - e->soc[n] > c->soc_eps) {
                double missing = tg - e->soc[n];
                cum_soc_missing += missing; n_viol++;
                reward += -500 / (1 + exp(-16.48461585 * (missing - 0.29229767))) + 1;         /* score_config.py:26-30 */
            } else {
                reward += c->fully_charged_reward;
            }
        }
        if (ntl != 0 && e->hl[n] != 0) e->hl[n] -= dt;                                         /* :593-594 */
        else if (ntl == 0) { e->hl[n] = ntl; e->soc[n] = nsr; }                                /* :597-599 */
        else { e->hl[n] = ntl; e->soc[n] = nsr; }                                              /* :602-606 arrival */
        if (e->soh[n] <= 0.9) e->target[n] = 0.9;                                              /* :613-614 */
    }
    /* note: the observer ran BEFORE the loop above (:511), so this step's aux block uses the pre-flip targets;
     * a flip only changes charging_left/hours_needed/laxity from the next step on.  build_obs below is called
     * with a snapshot taken here to reproduce that. */
    for (int n = 0; n < N; n++) if (e->hl[n] != 0) e->soc_deg[n] = e->soc[n];                   /* :621-623 */
    int done = (e->t == e->t_fin) || e->done_sticky;                                           /* :627-628 */
    if (done) e->done_sticky = 1;
    e->ep_return += reward;                                                                    /* :637 */

    so->reward = reward; so->cashflow = cashflow; so->overload = overload; so->soc_viol = fabs(cum_soc_missing);
    so->n_viol = n_viol; so->done = done;
    so->penalty = reward - (cashflow * c->price_multiplier);                                   /* :659 */

    if (c->calc_degradation) hist_append(o, e);                                                /* :655-656 */
    so->degradation = 0;
    if (c->calc_degradation && hour == 14 && tb->minute[tn] == 45) {                           /* :665 */
        if (c->deg_mode == FLEET_DEG_EMPIRICAL) empirical_degradation(o, e, deg_scratch);
        else sei_degradation(o, e, deg_scratch);
        for (int n = 0; n < N; n++) {
            e->soh[n] = e->soh[n] - deg_scratch[n];                                            /* :671 */
            e->cap[n] = e->soh[n] * c->init_battery_cap;                                       /* :673 */
            so->degradation += deg_scratch[n];
        }
    }
    (void)obs;
}

/* ---------------------------------------------------------------------------------------------- public API */

Oracle* oracle_create(const FleetConsts* consts, const FleetTables* tb, int32_t E, int64_t env_id_offset) {
    if (!consts || !tb || consts->abi_version != FLEETSTEP_ABI_VERSION || !consts->include_price) return NULL;
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    o->c = *consts; o->E = E; o->env_id_offset = env_id_offset; o->D = obs_dim(consts);
    const size_t N = (size_t)consts->num_evs, T = (size_t)consts->table_len;
    o->tb.there = (const uint8_t*)dupmem(tb->there, N * T);
    o->tb.time_left = (const double*)dupmem(tb->time_left, N * T * 8);
    o->tb.soc_on_return = (const double*)dupmem(tb->soc_on_return, N * T * 8);
    o->tb.delu = (const double*)dupmem(tb->delu, T * 8);
    o->tb.tariff = (const double*)dupmem(tb->tariff, T * 8);
    o->tb.load = (const double*)dupmem(tb->load, T * 8);
    o->tb.pv = (const double*)dupmem(tb->pv, T * 8);
    o->tb.price_reward_curve = (const double*)dupmem(tb->price_reward_curve, T * 8);
    o->tb.tariff_reward_curve = (const double*)dupmem(tb->tariff_reward_curve, T * 8);
    o->tb.cal_sincos = (const double*)dupmem(tb->cal_sincos, T * 6 * 8);
    o->tb.hour = (const uint8_t*)dupmem(tb->hour, T);
    o->tb.minute = (const uint8_t*)dupmem(tb->minute, T);
    o->envs = (OracleEnv*)calloc((size_t)E, sizeof(OracleEnv));
    for (int32_t i = 0; i < E; i++) {
        OracleEnv* e = &o->envs[i];
        double* blk = (double*)calloc(12 * N, sizeof(double));
        e->soc = blk; e->hl = blk + N; e->soc_deg = blk + 2 * N; e->soh = blk + 3 * N; e->cap = blk + 4 * N;
        e->target = blk + 5 * N; e->rf_len = blk + 6 * N; e->fd_cyc = blk + 7 * N; e->life = blk + 8 * N;
        e->sei_soh = blk + 9 * N; e->last_deg = blk + 10 * N; e->charge_log = blk + 11 * N;
        e->n_cycles = (int32_t*)calloc(N, sizeof(int32_t));
        for (size_t n = 0; n < N; n++) {
            e->target[n] = 1.0 * consts->target_soc;                    /* fleet_environment.py:263 */
            e->rf_len[n] = 1; e->fd_cyc[n] = 0;                          /* rainflow_sei_degradation.py:57-60 */
            e->sei_soh[n] = 1.0 * consts->init_soh; e->life[n] = 1 - e->sei_soh[n];   /* :31-34 */
            e->soh[n] = consts->init_soh; e->cap[n] = e->soh[n] * consts->init_battery_cap;
        }
        e->hist_cap = consts->episode_steps + 2;
        e->hist = (double*)malloc(sizeof(double) * (size_t)e->hist_cap * N);
    }
    return o;
}

static void pool_destroy(struct Pool* p);

void oracle_destroy(Oracle* o) {
    if (!o) return;
    pool_destroy(o->pool);
    for (int32_t i = 0; i < o->E; i++) { free(o->envs[i].soc); free(o->envs[i].n_cycles); free(o->envs[i].hist); }
    free(o->envs);
    free((void*)o->tb.there); free((void*)o->tb.time_left); free((void*)o->tb.soc_on_return);
    free((void*)o->tb.delu); free((void*)o->tb.tariff); free((void*)o->tb.load); free((void*)o->tb.pv);
    free((void*)o->tb.price_reward_curve); free((void*)o->tb.tariff_reward_curve);
    free((void*)o->tb.cal_sincos); free((void*)o->tb.hour); free((void*)o->tb.minute);
    free(o);
}

int32_t oracle_obs_dim(const Oracle* o) { return o->D; }

void oracle_set_next_start(Oracle* o, const int32_t* next_start) { o->next_start = next_start; }

void oracle_reset(Oracle* o, const int32_t* start_idx, const uint8_t* mask, float* obs) {
    for (int32_t i = 0; i < o->E; i++) {
        if (mask && !mask[i]) continue;
        OracleEnv* e = &o->envs[i];
        int32_t t0 = start_idx ? start_idx[i] : draw_start(&o->c, o->env_id_offset + i, e->ep_count);
        env_reset(o, e, t0, obs ? obs + (size_t)i * o->D : NULL);
    }
}

/* One contiguous range of envs; the body of oracle_step. */
typedef struct StepJob {
    Oracle* o; int32_t lo, hi;
    const float* actions; float* obs; double* reward64; double* cashflow; uint8_t* done; float* terminal_obs;
    double st[FLEET_S__COUNT];
} StepJob;

static void* step_range(void* arg) {
    StepJob* j = (StepJob*)arg;
    Oracle* o = j->o;
    const int32_t N = o->c.num_evs, D = o->D;
    double* scratch = (double*)malloc(sizeof(double) * 2 * (size_t)N);
    float* obs_tmp = (float*)malloc(sizeof(float) * (size_t)D);
    double* lst = j->st;
    for (int k = 0; k < FLEET_S__COUNT; k++) lst[k] = 0;
    for (int32_t i = j->lo; i < j->hi; i++) {
        OracleEnv* e = &o->envs[i];
        StepOut so;
        float* dst = j->obs ? j->obs + (size_t)i * D : obs_tmp;
        if (!o->c.auto_reset && e->done_sticky) {
            /* frozen after the episode end when the caller has not reset (the reference would keep
             * simulating past finish_time; that use is undefined and not reproduced) */
            build_obs(o, e->t, e->soc, e->hl, e->target, dst);
            if (j->reward64) j->reward64[i] = 0;
            if (j->cashflow) j->cashflow[i] = 0;
            if (j->done) j->done[i] = 1;
            e->last_reward = 0; e->last_cashflow = 0; e->last_overload = 0; e->last_soc_viol = 0;
            memset(e->charge_log, 0, sizeof(double) * N);
            continue;
        }
        double* target_before = scratch + N;      /* aux block uses the targets as of the observer call (:511) */
        memcpy(target_before, e->target, sizeof(double) * N);
        /* next_soc and the degradation scratch share scratch[0..N): they are live at disjoint times */
        env_step(o, e, j->actions + (size_t)i * N, NULL, &so, scratch, scratch);
        build_obs(o, e->t, e->soc, e->hl, target_before, dst);                                 /* :645-652 */
        e->last_reward = so.reward; e->last_cashflow = so.cashflow;
        e->last_overload = so.overload; e->last_soc_viol = so.soc_viol;
        if (j->reward64) j->reward64[i] = so.reward;
        if (j->cashflow) j->cashflow[i] = so.cashflow;
        if (j->done) j->done[i] = (uint8_t)so.done;
        lst[FLEET_S_STEPS] += 1; lst[FLEET_S_REWARD] += so.reward; lst[FLEET_S_CASHFLOW] += so.cashflow;
        lst[FLEET_S_PENALTY] += so.penalty; lst[FLEET_S_OVERLOAD_KW] += so.overload;
        lst[FLEET_S_SOC_VIOL] += so.soc_viol; lst[FLEET_S_N_VIOL] += so.n_viol;
        lst[FLEET_S_DEGRADATION] += so.degradation;
        if (so.done) {
            lst[FLEET_S_EPISODES] += 1; lst[FLEET_S_EP_RETURN] += e->ep_return;
            e->last_ep_return = e->ep_return;
            if (o->c.auto_reset) {
                if (j->terminal_obs) memcpy(j->terminal_obs + (size_t)i * D, dst, sizeof(float) * D);
                int32_t t0 = o->next_start ? o->next_start[i]
                                           : draw_start(&o->c, o->env_id_offset + i, e->ep_count);
                env_reset(o, e, t0, dst);
            }
        }
    }
    free(scratch); free(obs_tmp);
    return NULL;
}

/* Persistent worker pool: the threads are created once per Oracle and woken for every step (creating and joining
 * them every step made the CPU-baseline numbers depend on the scheduler's mood). */
typedef struct Pool {
    pthread_mutex_t mu;
    pthread_cond_t go, done_cv;
    pthread_t* th;
    StepJob* jobs;
    int32_t n, generation, pending, quit;
} Pool;

typedef struct { Pool* p; int32_t k; } WorkerArg;

static void* pool_worker(void* arg) {
    WorkerArg* wa = (WorkerArg*)arg;
    Pool* p = wa->p;
    const int32_t k = wa->k;
    free(wa);
    int32_t seen = 0;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (p->generation == seen && !p->quit) pthread_cond_wait(&p->go, &p->mu);
        if (p->quit) { pthread_mutex_unlock(&p->mu); return NULL; }
        seen = p->generation;
        pthread_mutex_unlock(&p->mu);
        step_range(&p->jobs[k]);
        pthread_mutex_lock(&p->mu);
        if (--p->pending == 0) pthread_cond_signal(&p->done_cv);
        pthread_mutex_unlock(&p->mu);
    }
}

static void pool_destroy(Pool* p) {
    if (!p) return;
    pthread_mutex_lock(&p->mu);
    p->quit = 1;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->mu);
    for (int32_t k = 1; k < p->n; k++) pthread_join(p->th[k], NULL);
    pthread_mutex_destroy(&p->mu); pthread_cond_destroy(&p->go); pthread_cond_destroy(&p->done_cv);
    free(p->th); free(p->jobs); free(p);
}

static Pool* pool_create(int32_t n) {
    Pool* p = (Pool*)calloc(1, sizeof(Pool));
    pthread_mutex_init(&p->mu, NULL); pthread_cond_init(&p->go, NULL); pthread_cond_init(&p->done_cv, NULL);
    p->n = n;
    p->th = (pthread_t*)calloc((size_t)n, sizeof(pthread_t));
    p->jobs = (StepJob*)calloc((size_t)n, sizeof(StepJob));
    for (int32_t k = 1; k < n; k++) {
        WorkerArg* wa = (WorkerArg*)malloc(sizeof(WorkerArg));
        wa->p = p; wa->k = k;
        pthread_create(&p->th[k], NULL, pool_worker, wa);
    }
    return p;
}

/* reward64/cashflow/terminal_obs may be NULL.  n_threads > 1 splits the envs over POSIX threads (used by the
 * CPU-baseline timing only; envs are independent). */
void oracle_step_mt(Oracle* o, const float* actions, float* obs, double* reward64, double* cashflow, uint8_t* done,
                    float* terminal_obs, int32_t n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > o->E) n_threads = o->E > 0 ? o->E : 1;
    if (o->pool && o->pool->n != n_threads) { pool_destroy(o->pool); o->pool = NULL; }
    if (!o->pool) o->pool = pool_create(n_threads);
    Pool* p = o->pool;
    for (int32_t k = 0; k < n_threads; k++) {
        StepJob* j = &p->jobs[k];
        j->o = o; j->lo = (int32_t)((int64_t)o->E * k / n_threads); j->hi = (int32_t)((int64_t)o->E * (k + 1) / n_threads);
        j->actions = actions; j->obs = obs; j->reward64 = reward64; j->cashflow = cashflow; j->done = done;
        j->terminal_obs = terminal_obs;
    }
    if (n_threads > 1) {
        pthread_mutex_lock(&p->mu);
        p->pending = n_threads - 1;
        p->generation++;
        pthread_cond_broadcast(&p->go);
        pthread_mutex_unlock(&p->mu);
    }
    step_range(&p->jobs[0]);
    if (n_threads > 1) {
        pthread_mutex_lock(&p->mu);
        while (p->pending > 0) pthread_cond_wait(&p->done_cv, &p->mu);
        pthread_mutex_unlock(&p->mu);
    }
    for (int32_t k = 0; k < n_threads; k++)
        for (int q = 0; q < FLEET_S__COUNT; q++) o->stats[q] += p->jobs[k].st[q];
}

void oracle_step(Oracle* o, const float* actions, float* obs, double* reward64, double* cashflow, uint8_t* done,
                 float* terminal_obs) {
    oracle_step_mt(o, actions, obs, reward64, cashflow, done, terminal_obs, 1);
}

void oracle_get_stats(const Oracle* o, double* dst) { memcpy(dst, o->stats, sizeof(o->stats)); }
void oracle_reset_stats(Oracle* o) { memset(o->stats, 0, sizeof(o->stats)); }
uint32_t oracle_err_flags(const Oracle* o) { return o->err_flags; }

/* Field getter mirroring fleet_get_state (host destination). hours_left is returned as float32 like the
 * product stores it; everything else in the product's dtype too. */
int32_t oracle_get_state(const Oracle* o, int32_t field, void* dst) {
    const int32_t N = o->c.num_evs;
    for (int32_t i = 0; i < o->E; i++) {
        const OracleEnv* e = &o->envs[i];
        switch (field) {
            case FLEET_F_SOC: memcpy((double*)dst + (size_t)i * N, e->soc, 8 * (size_t)N); break;
            case FLEET_F_HOURS_LEFT: for (int n = 0; n < N; n++) ((float*)dst)[(size_t)i * N + n] = (float)e->hl[n]; break;
            case FLEET_F_SOC_DEG: memcpy((double*)dst + (size_t)i * N, e->soc_deg, 8 * (size_t)N); break;
            case FLEET_F_SOH: memcpy((double*)dst + (size_t)i * N, e->soh, 8 * (size_t)N); break;
            case FLEET_F_TARGET_SOC: memcpy((double*)dst + (size_t)i * N, e->target, 8 * (size_t)N); break;
            case FLEET_F_TIME_IDX: ((int32_t*)dst)[i] = e->t; break;
            case FLEET_F_FINISH_IDX: ((int32_t*)dst)[i] = e->t_fin; break;
            case FLEET_F_REWARD64: ((double*)dst)[i] = e->last_reward; break;
            case FLEET_F_CASHFLOW: ((double*)dst)[i] = e->last_cashflow; break;
            case FLEET_F_RF_LEN: for (int n = 0; n < N; n++) ((int32_t*)dst)[(size_t)i * N + n] = (int32_t)e->rf_len[n]; break;
            case FLEET_F_FD_CYC: memcpy((double*)dst + (size_t)i * N, e->fd_cyc, 8 * (size_t)N); break;
            case FLEET_F_LIFE: memcpy((double*)dst + (size_t)i * N, e->life, 8 * (size_t)N); break;
            case FLEET_F_EP_RETURN: ((double*)dst)[i] = e->ep_return; break;
            case FLEET_F_EP_COUNT: ((int32_t*)dst)[i] = e->ep_count; break;
            case FLEET_F_LAST_EP_RETURN: ((double*)dst)[i] = e->last_ep_return; break;
            case FLEET_F_N_CYCLES: memcpy((int32_t*)dst + (size_t)i * N, e->n_cycles, 4 * (size_t)N); break;
            case FLEET_F_LAST_DEG: memcpy((double*)dst + (size_t)i * N, e->last_deg, 8 * (size_t)N); break;
            case FLEET_F_CHARGE_LOG: memcpy((double*)dst + (size_t)i * N, e->charge_log, 8 * (size_t)N); break;
            case FLEET_F_OVERLOAD: ((double*)dst)[i] = e->last_overload; break;
            case FLEET_F_SOC_VIOL: ((double*)dst)[i] = e->last_soc_viol; break;
            default: return -1;
        }
    }
    return 0;
}

int32_t oracle_set_state(Oracle* o, int32_t field, const void* src) {
    const int32_t N = o->c.num_evs;
    for (int32_t i = 0; i < o->E; i++) {
        OracleEnv* e = &o->envs[i];
        switch (field) {
            case FLEET_F_SOC: memcpy(e->soc, (const double*)src + (size_t)i * N, 8 * (size_t)N); break;
            case FLEET_F_SOH:
                memcpy(e->soh, (const double*)src + (size_t)i * N, 8 * (size_t)N);
                for (int n = 0; n < N; n++) e->cap[n] = e->soh[n] * o->c.init_battery_cap;
                break;
            case FLEET_F_TARGET_SOC: memcpy(e->target, (const double*)src + (size_t)i * N, 8 * (size_t)N); break;
            default: return -1;
        }
    }
    return 0;
}
