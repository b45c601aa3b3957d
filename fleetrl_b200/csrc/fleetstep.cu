// fleetstep.cu — B200 (sm_100a) implementation of the FleetRL environment step behind include/fleetstep.h.
//
// One fused kernel per call:
//   fleet_step_kernel   EvCharger.charge + LoadCalculation.check_violation + ScoreConfig penalties + time advance +
//                       departure/arrival logic + Observer/Normalization observation assembly + (at the daily 14:45
//                       trigger) rainflow/SEI or linear degradation + SB3-style auto-reset, for E envs x N EVs.
//   fleet_reset_kernel  FleetEnv.reset for the masked envs.
// Reference: fleetrl/fleet_env/fleet_environment.py:330-702, utils/ev_charging/ev_charger.py:39-231,
// utils/load_calculation/load_calculation.py:83-94, fleet_env/config/score_config.py:26-41,
// utils/observation/observer_bl_pv.py:12-136, utils/normalization/*.py,
// utils/battery_degradation/rainflow_sei_degradation.py:91-212, empirical_degradation.py:29-99.
//
// Mapping: a CTA owns a tile of B = floor(256/N) consecutive envs; thread j of the tile owns the
// (env, EV) pair ("slot") j = b*N + n.  All [E][N] state is env-major, so slot j of the tile touches element
// e0*N + j of every array: perfectly coalesced scalar loads/stores with no padding.  Per-env reductions over the
// EVs go through shared memory and are summed SEQUENTIALLY IN CAR ORDER by one thread per (env, quantity), which
// reproduces the reference's Python accumulation order bit for bit.  Everything that depends only on the time
// index (price/tariff factors, PV share, grid margin, the observation header with its look-ahead windows and
// calendar features) is precomputed once on the host in float64 with the reference's operation order and staged in
// HBM (L2-resident, ~35 MB at N=50,T=35040); the kernel gathers it by time index.
//
// Arithmetic: float64 with -fmad=false (no FMA contraction), operation order of the reference; float32 only where
// the reference casts (observation, reward output) and for hours_left, whose values are exact multiples of dt
// (validated at create time).
//
// The product path has NO CPU fallback: every entry point fails with FLEET_E_CUDA if the device is unusable.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "fleetstep.h"

namespace {

constexpr int kThreads = 256;        // threads per CTA
constexpr int kStatStripes = 64;     // atomic striping of the statistics vector
constexpr int kNQ = 9;               // per-slot contributions: cost, rev, cr, dr, inv, oc, a*there, missing, dep
enum { Q_COST = 0, Q_REV, Q_CR, Q_DR, Q_INV, Q_OC, Q_ATH, Q_MISS, Q_DEP };
constexpr int kNSum = 8;             // quantities 0..7 are summed in phase 2; Q_DEP is chained onto the reward

// per-time flags
constexpr uint32_t TF_TRIGGER = 1u;  // hour == 14 && minute == 45     fleet_environment.py:665
constexpr uint32_t TF_LUNCH = 2u;    // 11 < hour < 15                 fleet_environment.py:538

// per-env tile flags
constexpr int EF_FROZEN = 1, EF_DONE = 2, EF_TRIGGER = 4, EF_LUNCH = 8, EF_RESET = 16;

// One (time, vehicle) schedule record: SOC_on_return, time_left, There at this row, There at the previous row.
struct __align__(16) EvRec {
    double sr;
    float tl;
    uint8_t there, there_prev;
    uint16_t pad;
};
static_assert(sizeof(EvRec) == 16, "EvRec must be 16 bytes");

// Everything EvCharger.charge / check_violation need that depends only on the time index t (host-precomputed).
struct __align__(16) StepRow {
    double S;         // delu[t]/1000.0 + fixed_markup/1000                       ev_charger.py:145,35,149
    double F_cr;      // ((-1*price_multiplier)*price_reward_curve[t])/1000       ev_charger.py:154-156
    double F_dr;      // ((-1*price_multiplier)*tariff_reward_curve[t])/1000      ev_charger.py:204-206
    double Rfac;      // discharging_eff*tariff[t]/1000*(1-feed_in_deduction)     ev_charger.py:196-199
    double pv_share;  // pv[t]*dt / max(sum(there[:,t]),1)                        ev_charger.py:134-142
    double gml;       // grid_connection - load[t]                                load_calculation.py:93
    double pvv;       // pv[t]
    uint32_t flags;       // TF_* of row t
    uint32_t flags_next;  // TF_* of row t+1
};
static_assert(sizeof(StepRow) == 64, "StepRow must be 64 bytes");

struct StepParams {
    // sizes
    int E, N, T, R, L, D, Ha, Hb, hdr_stride, B;
    // flags
    int is_ct, calc_deg, deg_mode, carry, auto_reset, stack_u16;
    int start_lo, start_hi;
    unsigned long long seed;
    long long env_id_offset;
    // constants
    double dt, P, eta_c, eta_d, mult, cap0, target, target_lunch, eps, def_soc, min_lax;
    double pen_inv, pen_oc, clip_oc, pen_ovl, full_reward, evse, grid, init_soh, lc_batt_cap, hn_den, price_mult;
    double temperature, max_tl, max_soc, max_hn;
    float dt_f;
    // tables
    const EvRec* ev_rec;      // [T][N]
    const StepRow* step_row;  // [T]
    const float* hdr;         // [T][hdr_stride]
    // state
    int4* env4;               // [E] {t, t_start, ep_count, unused}
    double* soc;              // [E][N]
    float* hl;                // [E][N]
    double* soh;              // [E][N]
    double* hist;             // [E][R][N] soc_deg history (row k = soc_deg after k steps of the episode)
    uint8_t* tflip;           // [E][N] target_soc raised to 0.9 (fleet_environment.py:613-614)
    int* n_flips;             // [1] number of set tflip bytes (0 => the array is never read)
    double* env_f64;          // [6][E] ep_return, last_ep_return, reward64, cashflow, overload, soc_viol
    int* rf_len;              // [E][N]
    double* fd_cyc;           // [E][N]
    double* life;             // [E][N]
    int* n_cycles;            // [E][N]
    double* last_deg;         // [E][N]
    double* stats;            // [kStatStripes][FLEET_S__COUNT]
    unsigned int* err_flags;  // [1]
    const int* next_start;    // [E] or nullptr
    // I/O
    const float* actions;
    float* obs;
    float* reward;
    uint8_t* done;
    float* terminal_obs;
    // reset kernel only
    const int* start_idx;
    const uint8_t* mask;
};

enum { EF_EP_RETURN = 0, EF_LAST_EP_RETURN, EF_REWARD64, EF_CASHFLOW, EF_OVERLOAD, EF_SOC_VIOL, EF__COUNT };

struct EnvS {  // per-env scratch of a tile, shared memory
    int t, t_start, ep_count, flags;
    double S, F_cr, F_dr, Rfac, pv_share, gml, pvv;
    float* obs_dst;
    double deg_sum;
    int t0_new;
    int pad;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Counter-based start-index draw keyed by (seed, global env id, episode number): integer work, bit-exact with
// the oracle's draw_start.  Replaces the reference's unseeded random.choice (random_time_picker.py:31).
__device__ __forceinline__ int draw_start(const StepParams& p, long long env_id, int episode_no) {
    unsigned long long h =
        mix64(p.seed ^ mix64((unsigned long long)env_id * 0xD1B54A32D192ED03ull + (unsigned long long)(unsigned)episode_no));
    unsigned long long span = (unsigned long long)(p.start_hi - p.start_lo) + 1ull;
    return p.start_lo + (int)(h % span);
}

__device__ __forceinline__ EvRec load_rec(const EvRec* ptr) {
    // one 128-bit read-only load
    const int4 v = __ldg(reinterpret_cast<const int4*>(ptr));
    EvRec r;
    r.sr = __hiloint2double(v.y, v.x);
    r.tl = __int_as_float(v.z);
    r.there = (uint8_t)(v.w & 0xff);
    r.there_prev = (uint8_t)((v.w >> 8) & 0xff);
    r.pad = 0;
    return r;
}

// Per-EV part of Observer.get_obs + normalize_obs (observer_bl_pv.py:85-98, oracle_normalization.py:65-66,146-150):
// simulated soc / hours_left, and the auxiliary block computed from the SCHEDULE columns of the row (SURVEY B-4).
template <bool kNorm, bool kAux>
__device__ __forceinline__ void write_ev_obs(const StepParams& p, float* __restrict__ o, int n, double soc, float hl,
                                             const EvRec& rec, double tgt) {
    const int N = p.N;
    o[n] = (float)soc;
    o[N + n] = kNorm ? (float)((double)hl / p.max_tl) : hl;
    if (kAux) {
        double th = (double)rec.there;
        double tt = tgt * th;
        double cl = tt - rec.sr;
        double hn = cl * p.lc_batt_cap / p.hn_den;
        double lax = ((double)rec.tl / (hn + 0.001) - 1) * th;
        lax = lax < 0 ? 0 : (lax > 5 ? 5 : lax);
        if (kNorm) {
            tt = tt / p.max_soc; cl = cl / p.max_soc; hn = hn / p.max_hn; lax = lax / 5;
        }
        float* a = o + 2 * N + p.Ha;
        a[n] = (float)th;
        a[N + n] = (float)tt;
        a[2 * N + n] = (float)cl;
        a[3 * N + n] = (float)hn;
        a[4 * N + n] = (float)lax;
    }
}

// Time-only part of the observation (price/tariff/load/pv windows, evse/grid terms, calendar sin/cos): a host-
// precomputed float32 row per time index, copied by the env's slot threads.
template <bool kAux>
__device__ __forceinline__ void copy_hdr(const StepParams& p, float* __restrict__ o, int n, int t) {
    const float* __restrict__ h = p.hdr + (size_t)t * p.hdr_stride;
    const int N = p.N;
    for (int q = n; q < p.Ha; q += N) o[2 * N + q] = __ldg(h + q);
    if (kAux)
        for (int q = n; q < p.Hb; q += N) o[2 * N + p.Ha + 5 * N + q] = __ldg(h + p.Ha + q);
}

// FleetEnv.reset for one (env, EV) slot: fleet_environment.py:345-348, 371-372, 382-399.
template <bool kNorm, bool kAux>
__device__ __forceinline__ void reset_slot(const StepParams& p, int e, int n, int t0, float* obs_row, bool reinit_deg) {
    const size_t i = (size_t)e * p.N + n;
    const EvRec rec = load_rec(&p.ev_rec[(size_t)t0 * p.N + n]);
    bool flip = false;
    if (reinit_deg) {
        // fresh-object semantics (carry_degradation_state == 0): re-initialise what lives on the reference's
        // env / degradation objects (rainflow_sei_degradation.py:31-34,57-60; fleet_environment.py:263)
        p.rf_len[i] = 1; p.fd_cyc[i] = 0; p.life[i] = 1 - p.init_soh;
        p.tflip[i] = 0;
    } else if (*p.n_flips != 0) {
        flip = p.tflip[i] != 0;
    }
    const double tgt = flip ? 0.9 : p.target;
    const double soh = 1.0 * p.init_soh;
    const double cap = soh * p.cap0;
    double soc = rec.sr;
    const float hl = rec.tl;
    const double time_needed = (tgt - soc) * cap / p.P;
    if (hl > 0.f && p.min_lax * time_needed > (double)hl) soc = tgt - (time_needed * p.P / cap) / p.min_lax;
    const double sdeg = (soc == 0) ? p.def_soc : soc;
    p.soc[i] = soc;
    p.hl[i] = hl;
    p.soh[i] = soh;
    p.hist[((size_t)e * p.R + 0) * p.N + n] = sdeg;
    if (obs_row) {
        write_ev_obs<kNorm, kAux>(p, obs_row, n, soc, hl, rec, tgt);
        copy_hdr<kAux>(p, obs_row, n, t0);
    }
}

// ------------------------------------------------------------------------------------------------ degradation
// rainflow 3.2.0 extract_cycles (ASTM E1049-85 three-point method) streamed over one vehicle's SOC history with
// an index stack in shared memory, feeding RainflowSeiDegradation.calculate_degradation
// (rainflow_sei_degradation.py:128-206) without materialising the cycle list.
//
// The reference recomputes the full cycle list at every daily call and then takes the POSITIONAL slice
// [rainflow_length-1 : len-1] (:146).  Streaming equivalent: cycles are numbered in generation order; cycle j
// contributes to fd_cyc iff rf_len-1 <= j < m-1 where m is the final count, so the newest cycle is held back as
// "pending" until the next one is emitted.  The mean over ALL cycles (:140) and max(End) = len-1 (:138) are
// accumulated on the way.
template <typename IdxT>
struct RfStack {
    IdxT* s;      // s[k * stride]
    int stride;
    __device__ __forceinline__ int get(int k) const { return (int)s[(size_t)k * stride]; }
    __device__ __forceinline__ void set(int k, int v) { s[(size_t)k * stride] = (IdxT)v; }
};

struct SeiAcc {
    int m;              // cycles emitted so far
    int a;              // first selected list position (rf_len - 1)
    double mean_sum;    // sum of cycle means, all cycles
    double fsum;        // sum of stress over selected cycles
    double max_dod;     // max range over selected cycles
    bool have_pending;
    double pend_eff, pend_mean, pend_range;
    double s_temp;
    bool stress;
};

__device__ __forceinline__ void sei_commit_pending(SeiAcc& acc) {
    // pending is list position m-1 at the time of the call (before the new cycle is counted)
    if (acc.have_pending && acc.stress && (acc.m - 1) >= acc.a) {
        const double kd1 = 1.4E5, kd2 = -5.01E-1, kd3 = -1.23E5, k_sigma = 1.04, sigma_ref = 0.5;
        // (kd1 * dod**kd2 + kd3) ** -1 ; dod == 0 -> inf -> 0                  rainflow_sei_degradation.py:68
        const double s_dod = 1.0 / (kd1 * pow(acc.pend_eff, kd2) + kd3);
        const double s_soc = exp(k_sigma * (acc.pend_mean - sigma_ref));        // :70
        acc.fsum += s_dod * s_soc * acc.s_temp;                                 // :77-79,174
        if (acc.pend_range > acc.max_dod) acc.max_dod = acc.pend_range;
    }
}

__device__ __forceinline__ void sei_emit(SeiAcc& acc, double xa, double xb, double count) {
    sei_commit_pending(acc);
    const double range = fabs(xa - xb);
    const double mean = 0.5 * (xa + xb);
    double eff = range * count;                                                  // :170
    eff = eff < 0 ? 0 : (eff > 1 ? 1 : eff);
    acc.pend_eff = eff; acc.pend_mean = mean; acc.pend_range = range; acc.have_pending = true;
    acc.mean_sum += mean;
    acc.m++;
}

template <typename IdxT>
__device__ __noinline__ void rainflow_stream(const double* __restrict__ x, int stride, int len, RfStack<IdxT> st,
                                             SeiAcc& acc) {
    if (len < 2) return;
    int lo = 0, hi = 0;
    double v1 = 0, v2 = 0, v3 = 0;  // values of the top three stack entries (v3 = top)
#define XVAL(idx) (x[(size_t)(idx) * stride])
#define RF_PUSH(idx, val)                                                                        \
    do {                                                                                         \
        st.set(hi, (idx)); hi++;                                                                 \
        v1 = v2; v2 = v3; v3 = (val);                                                            \
        while (hi - lo >= 3) {                                                                   \
            const double X = fabs(v3 - v2), Y = fabs(v2 - v1);                                   \
            if (X < Y) break;                                                                    \
            if (hi - lo == 3) { sei_emit(acc, v1, v2, 0.5); lo++; }                              \
            else {                                                                               \
                sei_emit(acc, v1, v2, 1.0);                                                      \
                st.set(hi - 3, st.get(hi - 1)); hi -= 2;                                         \
                v2 = XVAL(st.get(hi - 2));                                                       \
                if (hi - lo >= 3) v1 = XVAL(st.get(hi - 3));                                     \
            }                                                                                    \
        }                                                                                        \
    } while (0)

    double x_last = XVAL(0), xc = XVAL(1);
    double d_last = xc - x_last;
    RF_PUSH(0, x_last);
    int index = -1;
    double x_next = 0;
    for (int pos = 2; pos < len; pos++) {
        index = pos - 1;
        x_next = XVAL(pos);
        if (x_next == xc) continue;
        const double d_next = x_next - xc;
        if (d_last * d_next < 0) RF_PUSH(index, xc);
        x_last = xc; xc = x_next; d_last = d_next;
    }
    if (index >= 0) RF_PUSH(index + 1, x_next);
    while (hi - lo > 1) {
        sei_emit(acc, XVAL(st.get(lo)), XVAL(st.get(lo + 1)), 0.5);
        lo++;
    }
#undef RF_PUSH
#undef XVAL
}

// RainflowSeiDegradation.calculate_degradation for one vehicle.  Returns the degradation (SOH loss).
template <typename IdxT>
__device__ __noinline__ double sei_eval(const StepParams& p, size_t i, const double* __restrict__ hcol, int len,
                                        RfStack<IdxT> st) {
    const double alpha_sei = 5.75E-2, beta_sei = 121, k_sigma = 1.04, sigma_ref = 0.5, k_temp = 6.93E-2,
                 temp_ref = 25, k_dt = 4.14E-10;
    const int rf_len = p.rf_len[i];
    SeiAcc acc;
    acc.m = 0; acc.a = rf_len - 1; acc.mean_sum = 0; acc.fsum = 0; acc.max_dod = 0; acc.have_pending = false;
    acc.pend_eff = acc.pend_mean = acc.pend_range = 0;
    acc.s_temp = exp(k_temp * (p.temperature - temp_ref) * ((temp_ref + 273.15) / (p.temperature + 273.15)));  // :72-73
    // With a carried-over rainflow_length the slice is usually empty: count first, evaluate stress only if needed.
    acc.stress = (rf_len <= 1);
    rainflow_stream<IdxT>(hcol, p.N, len, st, acc);
    if (!acc.stress && acc.m > rf_len) {
        acc.m = 0; acc.mean_sum = 0; acc.fsum = 0; acc.max_dod = 0; acc.have_pending = false; acc.stress = true;
        rainflow_stream<IdxT>(hcol, p.N, len, st, acc);
    }
    const int m = acc.m;
    p.n_cycles[i] = m;
    double deg = 0;
    if (m > rf_len) {                                                                  // :143
        const double battery_age = (double)(len - 1) * p.dt * 3600;                    // :138  max(End) == len-1
        const double mean_soc_cal = acc.mean_sum / (double)m;                          // :140
        if (acc.max_dod > 5) atomicOr(p.err_flags, 4u);                                // :164-167
        const double fd_cal = (k_dt * battery_age) * exp(k_sigma * (mean_soc_cal - sigma_ref)) * acc.s_temp;  // :81-83
        const double fd_cyc = p.fd_cyc[i] + acc.fsum;                                  // :174
        p.fd_cyc[i] = fd_cyc;
        const double fd = fd_cyc + fd_cal;
        const double l_old = p.life[i];
        double new_l;
        if (p.init_soh == 1.0) {
            new_l = 1 - alpha_sei * exp(-beta_sei * fd) - (1 - alpha_sei) * exp(-fd);  // :85-86
            if (new_l < 0) atomicOr(p.err_flags, 2u);                                  // :179-180
        } else {
            new_l = 1 - (1 - l_old) * exp(-fd);                                        // :89,186
        }
        deg = new_l - l_old;                                                           // :189
        p.life[i] = new_l;                                                             // :192
        p.rf_len[i] = m;                                                               // :195
    }
    return deg;
}

// EmpiricalDegradation.calculate_degradation for one vehicle (empirical_degradation.py:60-94).
__device__ __forceinline__ double empirical_eval(const StepParams& p, const double* __restrict__ hcol, int len) {
    const double old_soc = hcol[(size_t)(len - 2) * p.N];
    const double new_soc = hcol[(size_t)(len - 1) * p.N];
    const double avg_soc = (old_soc + new_soc) / 2;
    const double cs[3] = {0, 40, 90};
    const double ca[3] = {0.0065, 0.0293, 0.065};
    int best = 0;
    double bd = fabs(cs[0] - avg_soc);
#pragma unroll
    for (int k = 1; k < 3; k++) {
        const double d = fabs(cs[k] - avg_soc);
        if (d < bd) { bd = d; best = k; }
    }
    const double cal = ca[best] * p.dt / 8760;
    const double dod = fabs(new_soc - old_soc);
    const double cyc = (p.evse <= 22.0) ? dod * 0.000125 / 2 : dod * 0.000167 / 2;
    return cal + cyc;
}

// ------------------------------------------------------------------------------------------------ step kernel
// Shared-memory layout (dynamic): EnvS envs[B] | double contrib[kNQ][slots] | double sums[kNSum][B] | index stack
__host__ __device__ inline size_t smem_envs_bytes(int B) { return ((size_t)B * sizeof(EnvS) + 15) & ~(size_t)15; }

template <bool kNorm, bool kAux, typename IdxT>
__global__ void __launch_bounds__(kThreads) fleet_step_kernel(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, B = p.B;
    EnvS* envs = reinterpret_cast<EnvS*>(smem_raw);
    double* contrib = reinterpret_cast<double*>(smem_raw + smem_envs_bytes(B));
    const int cstride = B * N;  // contrib[q][slot]
    double* sums = contrib + (size_t)kNQ * cstride;
    IdxT* stack_base = reinterpret_cast<IdxT*>(sums + (size_t)kNSum * B);
    __shared__ int s_any;  // bit0: some env triggered degradation, bit1: some env auto-resets

    const int tid = threadIdx.x;
    const int e0 = blockIdx.x * B;
    const int nb = min(B, p.E - e0);
    const int nslots = nb * N;

    // ---- P0: one thread per env: time index, per-time factors, destination of the observation row
    if (tid == 0) s_any = 0;
    if (tid < nb) {
        const int e = e0 + tid;
        const int4 ev = p.env4[e];
        EnvS& es = envs[tid];
        es.t = ev.x; es.t_start = ev.y; es.ep_count = ev.z;
        const int t_fin = ev.y + p.L;
        int fl = 0;
        if (!p.auto_reset && ev.x >= t_fin) fl |= EF_FROZEN;   // episode over and not reset by the caller
        const int t = min(ev.x, p.T - 2);
        const StepRow* r = p.step_row + t;
        const double2 ra = __ldg(reinterpret_cast<const double2*>(r));
        const double2 rb = __ldg(reinterpret_cast<const double2*>(r) + 1);
        const double2 r1 = __ldg(reinterpret_cast<const double2*>(r) + 2);
        const double2 r2 = __ldg(reinterpret_cast<const double2*>(r) + 3);
        es.S = ra.x; es.F_cr = ra.y; es.F_dr = rb.x; es.Rfac = rb.y; es.pv_share = r1.x; es.gml = r1.y; es.pvv = r2.x;
        const uint32_t fn = (uint32_t)(__double_as_longlong(r2.y) >> 32);  // flags_next (little endian: second u32)
        if (!(fl & EF_FROZEN)) {
            if (ev.x + 1 == t_fin) fl |= EF_DONE;
            if ((fn & TF_TRIGGER) && p.calc_deg) fl |= EF_TRIGGER;
            if (fn & TF_LUNCH) fl |= EF_LUNCH;
            if ((fl & EF_DONE) && p.auto_reset) fl |= EF_RESET;
        }
        es.flags = fl;
        // SB3 semantics: a finished env returns the first observation of its next episode in obs and the last
        // observation of the finished one in infos["terminal_observation"].
        if (fl & EF_RESET) es.obs_dst = p.terminal_obs ? p.terminal_obs + (size_t)e * p.D : nullptr;
        else es.obs_dst = p.obs ? p.obs + (size_t)e * p.D : nullptr;
        es.deg_sum = 0;
        es.t0_new = 0;
        int any = 0;
        if (fl & EF_TRIGGER) any |= 1;
        if (fl & EF_RESET) any |= 2;
        if (any) atomicOr(&s_any, any);
    }
    __syncthreads();

    const bool have_flips = (*p.n_flips != 0);

    // ---- P1: one thread per (env, EV): charge / discharge, transition, per-EV observation parts
    for (int j = tid; j < nslots; j += kThreads) {
        const int b = j / N, n = j - b * N;
        const EnvS& es = envs[b];
        const int e = e0 + b;
        const size_t i = (size_t)e * N + n;
        double c_cost = 0, c_rev = 0, c_cr = 0, c_dr = 0, c_inv = 0, c_oc = 0, c_ath = 0, c_miss = 0, c_dep = 0;
        if (es.flags & EF_FROZEN) {
            if (es.obs_dst) {
                const EvRec rec = load_rec(&p.ev_rec[(size_t)min(es.t, p.T - 1) * N + n]);
                const bool flip = have_flips && p.tflip[i] != 0;
                write_ev_obs<kNorm, kAux>(p, es.obs_dst, n, p.soc[i], p.hl[i], rec, flip ? 0.9 : p.target);
                copy_hdr<kAux>(p, es.obs_dst, n, min(es.t, p.T - 1));
            }
        } else {
            const int t = es.t;
            const int k = t - es.t_start;
            const float a32 = p.actions[i];
            double soc = p.soc[i];
            float hl = p.hl[i];
            const double soh = p.soh[i];
            double sdeg = p.hist[((size_t)e * p.R + (k % p.R)) * N + n];
            const EvRec rec = load_rec(&p.ev_rec[(size_t)min(t + 1, p.T - 1) * N + n]);
            const bool flip = have_flips && p.tflip[i] != 0;
            const double tgt = flip ? 0.9 : p.target;                       // FleetEnv.target_soc[car]
            const double cap = soh * p.cap0;                                // episode.battery_cap[car]
            const int there = rec.there_prev;                               // db.There at t
            const double a = (double)a32;
            double nsoc = soc;
            if (a >= 0) {                                                   // ev_charger.py:98-156
                const double dem = (tgt - soc) * cap;
                const double req = p.P * a * p.dt;
                if (req * p.eta_c > dem) {
                    const double d = req - dem;
                    const double pen = p.pen_oc * (d * d);
                    c_oc = pen > p.clip_oc ? pen : p.clip_oc;
                }
                double en;
                if (there == 1) en = fmin(dem / p.eta_c, req);
                else {
                    en = 0;
                    if (fabs(a) > 0.05) c_inv = p.pen_inv * (a * a);
                }
                nsoc = soc + en * p.eta_c / cap;
                double ge = en - es.pv_share;
                ge = ge > 0 ? ge : 0;
                c_cost = ge * es.S * p.mult;
                c_cr = es.F_cr * ge;
            } else if (a < 0) {                                             // ev_charger.py:159-206
                const double left = -1 * soc * cap;
                const double req = p.P * a * p.dt;
                if (req * p.eta_d < left && there != 0) {
                    const double d = left - req;
                    c_oc = p.pen_oc * (d * d);
                }
                double en;
                if (there == 1) en = fmax(left, req);
                else {
                    en = 0.0;
                    if (fabs(a) > 0.05) c_inv = p.pen_inv * (a * a);
                }
                nsoc = soc + en / cap;
                c_rev = -1 * en * es.Rfac;
                c_dr = es.F_dr * en;
            } else {
                atomicOr(p.err_flags, 1u);                                  // NaN action: TypeError ev_charger.py:209
            }
            c_ath = a * (double)there;                                      // fleet_environment.py:491
            soc = nsoc;                                                     // :470

            // time has advanced to t+1: departure / still there / gone / arrival   :528-618
            const float ntl = rec.tl;
            if (hl != 0.f && ntl == 0.f) {
                const double tg = (p.is_ct && (es.flags & EF_LUNCH)) ? p.target_lunch : tgt;
                const double diff = tg - soc;
                if (diff > p.eps) {
                    c_miss = diff;
                    c_dep = -500 / (1 + exp(-16.48461585 * (diff - 0.29229767))) + 1;   // score_config.py:26-30
                } else {
                    c_dep = p.full_reward;
                }
            }
            if (ntl != 0.f && hl != 0.f) hl -= p.dt_f;
            else { hl = ntl; soc = rec.sr; }
            if (soh <= 0.9 && !flip) {                                      // :613-614 (visible from the next step on)
                p.tflip[i] = 1;
                atomicAdd(p.n_flips, 1);
            }
            if (hl != 0.f) sdeg = soc;                                      // :621-623

            p.soc[i] = soc;
            p.hl[i] = hl;
            p.hist[((size_t)e * p.R + ((k + 1) % p.R)) * N + n] = sdeg;     // log_soc, :655-656
            if (es.obs_dst) {
                write_ev_obs<kNorm, kAux>(p, es.obs_dst, n, soc, hl, rec, tgt);
                copy_hdr<kAux>(p, es.obs_dst, n, min(t + 1, p.T - 1));
            }
        }
        contrib[Q_COST * cstride + j] = c_cost; contrib[Q_REV * cstride + j] = c_rev;
        contrib[Q_CR * cstride + j] = c_cr;     contrib[Q_DR * cstride + j] = c_dr;
        contrib[Q_INV * cstride + j] = c_inv;   contrib[Q_OC * cstride + j] = c_oc;
        contrib[Q_ATH * cstride + j] = c_ath;   contrib[Q_MISS * cstride + j] = c_miss;
        contrib[Q_DEP * cstride + j] = c_dep;
    }
    __syncthreads();

    // ---- P2: one thread per (quantity, env): sequential sum over the env's EVs in car order
    for (int w = tid; w < kNSum * nb; w += kThreads) {
        const int q = w / nb, b = w - q * nb;
        const double* c = contrib + (size_t)q * cstride + (size_t)b * N;
        double s = 0;
        for (int n = 0; n < N; n++) s += c[n];
        sums[q * B + b] = s;
    }
    __syncthreads();

    // ---- P3: one thread per env: cashflow, reward, overload penalty, departure terms, done, statistics
    double st_loc[FLEET_S__COUNT];
#pragma unroll
    for (int q = 0; q < FLEET_S__COUNT; q++) st_loc[q] = 0;
    if (tid < nb) {
        const int b = tid, e = e0 + b;
        EnvS& es = envs[b];
        double reward = 0, cashflow = 0, overload = 0, soc_viol = 0;
        int done = 0;
        if (es.flags & EF_FROZEN) {
            done = 1;
        } else {
            cashflow = -1 * sums[Q_COST * B + b] + sums[Q_REV * B + b];                        // ev_charger.py:225
            reward = sums[Q_CR * B + b] + sums[Q_DR * B + b] + sums[Q_INV * B + b] + sums[Q_OC * B + b];  // :228
            const double margin = es.gml - sums[Q_ATH * B + b] * p.evse + es.pvv;              // load_calculation.py:93
            overload = fabs(margin < 0.0 ? margin : 0.0);
            if (overload > 0) {
                const double rel = overload / p.grid + 1;                                      // fleet_environment.py:496
                const double pen = (rel < 1.1) ? 0.0 : -700 / (1 + exp(-15.77350877 * (rel - 1.33298382)));
                reward += pen * p.pen_ovl;                                                     // score_config.py:33-41
            }
            const double* cd = contrib + (size_t)Q_DEP * cstride + (size_t)b * N;
            const double* cm = contrib + (size_t)Q_MISS * cstride + (size_t)b * N;
            int n_viol = 0;
            for (int n = 0; n < N; n++) { reward += cd[n]; n_viol += (cm[n] > 0) ? 1 : 0; }  // :548,554,583,590
            soc_viol = fabs(sums[Q_MISS * B + b]);                                             // :661
            done = (es.flags & EF_DONE) ? 1 : 0;                                               // :627-628
            double ep_ret = p.env_f64[(size_t)EF_EP_RETURN * p.E + e] + reward;                // :637
            st_loc[FLEET_S_STEPS] = 1; st_loc[FLEET_S_REWARD] = reward; st_loc[FLEET_S_CASHFLOW] = cashflow;
            st_loc[FLEET_S_PENALTY] = reward - (cashflow * p.price_mult);                      // :659
            st_loc[FLEET_S_OVERLOAD_KW] = overload; st_loc[FLEET_S_SOC_VIOL] = soc_viol;
            st_loc[FLEET_S_N_VIOL] = n_viol;
            int t_new = es.t + 1, t_start = es.t_start, ep_count = es.ep_count;
            if (done) {
                st_loc[FLEET_S_EPISODES] = 1; st_loc[FLEET_S_EP_RETURN] = ep_ret;
                p.env_f64[(size_t)EF_LAST_EP_RETURN * p.E + e] = ep_ret;
                if (es.flags & EF_RESET) {
                    const int t0 = p.next_start ? p.next_start[e] : draw_start(p, p.env_id_offset + e, ep_count);
                    es.t0_new = t0;
                    t_new = t0; t_start = t0; ep_count += 1; ep_ret = 0;
                }
            }
            p.env_f64[(size_t)EF_EP_RETURN * p.E + e] = ep_ret;
            p.env4[e] = make_int4(t_new, t_start, ep_count, 0);
        }
        p.env_f64[(size_t)EF_REWARD64 * p.E + e] = reward;
        p.env_f64[(size_t)EF_CASHFLOW * p.E + e] = cashflow;
        p.env_f64[(size_t)EF_OVERLOAD * p.E + e] = overload;
        p.env_f64[(size_t)EF_SOC_VIOL * p.E + e] = soc_viol;
        if (p.reward) p.reward[e] = (float)reward;
        if (p.done) p.done[e] = (uint8_t)done;
    }

    // ---- P4: daily degradation for the envs of the tile whose new time is 14:45   fleet_environment.py:665-673
    const int any = s_any;  // written before the first barrier, stable since
    if (any & 1) {
        for (int j = tid; j < nslots; j += kThreads) {
            const int b = j / N, n = j - b * N;
            EnvS& es = envs[b];
            if (!(es.flags & EF_TRIGGER)) continue;
            const int e = e0 + b;
            const size_t i = (size_t)e * N + n;
            const int len = es.t - es.t_start + 2;  // history rows 0..k+1
            const double* hcol = p.hist + (size_t)e * p.R * N + n;
            double deg;
            if (p.deg_mode == FLEET_DEG_EMPIRICAL) {
                deg = empirical_eval(p, hcol, len);
                p.n_cycles[i] = 0;
            } else {
                RfStack<IdxT> st;
                st.s = stack_base + (j % kThreads);
                st.stride = kThreads;
                deg = sei_eval<IdxT>(p, i, hcol, len, st);
            }
            p.last_deg[i] = deg;
            p.soh[i] = p.soh[i] - deg;                                      // :671 (battery_cap is derived, :673)
            atomicAdd(&es.deg_sum, deg);
        }
        __syncthreads();
        if (tid < nb) st_loc[FLEET_S_DEGRADATION] = envs[tid].deg_sum;
    }

    // statistics: warp reduce over the env threads, one striped atomic per warp and quantity
    if (tid < ((nb + 31) & ~31)) {
#pragma unroll
        for (int q = 0; q < FLEET_S__COUNT; q++) {
            double v = st_loc[q];
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if ((tid & 31) == 0 && v != 0)
                atomicAdd(p.stats + (size_t)((blockIdx.x + (tid >> 5)) % kStatStripes) * FLEET_S__COUNT + q, v);
        }
    }

    // ---- P5: auto-reset of finished envs (SubprocVecEnv worker: reset right after a done step)
    if (any & 2) {
        __syncthreads();
        for (int j = tid; j < nslots; j += kThreads) {
            const int b = j / N, n = j - b * N;
            const EnvS& es = envs[b];
            if (!(es.flags & EF_RESET)) continue;
            const int e = e0 + b;
            reset_slot<kNorm, kAux>(p, e, n, es.t0_new, p.obs ? p.obs + (size_t)e * p.D : nullptr, !p.carry);
        }
    }
}

template <bool kNorm, bool kAux>
__global__ void __launch_bounds__(kThreads) fleet_reset_kernel(const StepParams p) {
    const int N = p.N, B = p.B;
    const int e0 = blockIdx.x * B;
    const int nb = min(B, p.E - e0);
    const int nslots = nb * N;
    for (int j = threadIdx.x; j < nslots; j += kThreads) {
        const int b = j / N, n = j - b * N;
        const int e = e0 + b;
        if (p.mask && !p.mask[e]) continue;
        const int4 ev = p.env4[e];
        const int t0 = p.start_idx ? p.start_idx[e] : draw_start(p, p.env_id_offset + e, ev.z);
        reset_slot<kNorm, kAux>(p, e, n, t0, p.obs ? p.obs + (size_t)e * p.D : nullptr, !p.carry);
    }
    __syncthreads();  // all slots have read env4 before it is rewritten
    if (threadIdx.x < nb) {
        const int e = e0 + threadIdx.x;
        if (!(p.mask && !p.mask[e])) {
            const int4 ev = p.env4[e];
            const int t0 = p.start_idx ? p.start_idx[e] : draw_start(p, p.env_id_offset + e, ev.z);
            p.env4[e] = make_int4(t0, t0, ev.z + 1, 0);
            p.env_f64[(size_t)EF_EP_RETURN * p.E + e] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------ small utility kernels
__global__ void init_state_kernel(StepParams p) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t EN = (size_t)p.E * p.N;
    if (i < EN) {
        p.soc[i] = 0; p.hl[i] = 0; p.soh[i] = p.init_soh;
        p.rf_len[i] = 1; p.fd_cyc[i] = 0; p.life[i] = 1 - p.init_soh; p.n_cycles[i] = 0; p.last_deg[i] = 0;
        p.tflip[i] = 0;
    }
    if (i < (size_t)p.E) p.env4[i] = make_int4(0, 0, 0, 0);
}

// field gathers that are not plain arrays
__global__ void gather_field_kernel(StepParams p, int field, void* dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t EN = (size_t)p.E * p.N;
    switch (field) {
        case FLEET_F_SOC_DEG:
            if (i < EN) {
                const int e = (int)(i / p.N), n = (int)(i - (size_t)e * p.N);
                const int4 ev = p.env4[e];
                const int k = ev.x - ev.y;
                ((double*)dst)[i] = p.hist[((size_t)e * p.R + (k % p.R)) * p.N + n];
            }
            break;
        case FLEET_F_TARGET_SOC:
            if (i < EN) ((double*)dst)[i] = (*p.n_flips != 0 && p.tflip[i]) ? 0.9 : p.target;
            break;
        case FLEET_F_TIME_IDX: if (i < (size_t)p.E) ((int*)dst)[i] = p.env4[i].x; break;
        case FLEET_F_FINISH_IDX: if (i < (size_t)p.E) ((int*)dst)[i] = p.env4[i].y + p.L; break;
        case FLEET_F_EP_COUNT: if (i < (size_t)p.E) ((int*)dst)[i] = p.env4[i].z; break;
        default: break;
    }
}

__global__ void scatter_target_kernel(StepParams p, const double* src) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)p.E * p.N) {
        const uint8_t f = (src[i] == 0.9 && p.target != 0.9) ? 1 : 0;
        p.tflip[i] = f;
        if (f) atomicAdd(p.n_flips, 1);
    }
}

__global__ void reduce_stats_kernel(const double* stats, double* dst) {
    const int q = threadIdx.x;
    if (q < FLEET_S__COUNT) {
        double s = 0;
        for (int k = 0; k < kStatStripes; k++) s += stats[(size_t)k * FLEET_S__COUNT + q];
        dst[q] = s;
    }
}

}  // namespace

// =============================================================================================== host side / C ABI

struct FleetHandle {
    FleetConsts c;
    int device = 0;
    int E = 0, N = 0, T = 0, D = 0;
    StepParams p;
    std::vector<void*> allocs;
    int64_t bytes = 0;
    int64_t launches = 0;
    std::string err;
    size_t smem_step = 0;
    int grid = 0;
    int max_smem_optin = 0;
    // host-call staging (fleet_step_host)
    float* h_actions_dev = nullptr; float* h_obs_dev = nullptr; float* h_reward_dev = nullptr; uint8_t* h_done_dev = nullptr;
};

namespace {

int fail(FleetHandle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return fail(h, FLEET_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

template <typename T>
int dev_alloc(FleetHandle* h, T** out, size_t count, bool zero = true) {
    void* ptr = nullptr;
    const size_t bytes = count * sizeof(T);
    cudaError_t e = cudaMalloc(&ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof buf, "cudaMalloc of %zu bytes failed: %s (handle already holds %lld bytes)", bytes,
                 cudaGetErrorString(e), (long long)h->bytes);
        return fail(h, FLEET_E_NOMEM, buf);
    }
    if (zero) cudaMemset(ptr, 0, bytes ? bytes : 16);
    h->allocs.push_back(ptr);
    h->bytes += (int64_t)bytes;
    *out = (T*)ptr;
    return FLEET_OK;
}

int obs_dim_of(const FleetConsts& c, int* Ha, int* Hb) {
    int a = 0, b = 0;
    if (c.include_price) a += 2 * (c.price_lookahead + 1);
    if (c.include_price && c.include_building) a += c.bl_pv_lookahead + 1;
    if (c.include_price && c.include_pv) a += c.bl_pv_lookahead + 1;
    if (c.aux) {
        b += 1 + 6;
        if (c.include_price && c.include_building) b += 3;
    }
    *Ha = a; *Hb = b;
    return 2 * c.num_evs + a + (c.aux ? 5 * c.num_evs : 0) + b;   // detect_dim_and_bounds, fleet_environment.py:854-949
}

// look-ahead element k at time index t (observer_bl_pv.py:50-80): 0 -> row t, k>=1 -> first row of the k-th next hour
inline int look_idx(const FleetConsts& c, const FleetTables& tb, int t, int k) {
    const int sph = c.steps_per_hour;
    const int pos = (int)tb.minute[t] * sph / 60;
    int i = (k == 0) ? t : t - pos + k * sph;
    if (i > c.table_len - 1) i = c.table_len - 1;
    return i;
}

using StepKernel = void (*)(const StepParams);

template <typename IdxT>
StepKernel pick_step(bool norm, bool aux) {
    if (norm) return aux ? fleet_step_kernel<true, true, IdxT> : fleet_step_kernel<true, false, IdxT>;
    return aux ? fleet_step_kernel<false, true, IdxT> : fleet_step_kernel<false, false, IdxT>;
}
StepKernel pick_step(const FleetHandle* h) {
    return h->p.stack_u16 ? pick_step<uint16_t>(h->c.normalize, h->c.aux) : pick_step<uint8_t>(h->c.normalize, h->c.aux);
}
StepKernel pick_reset(const FleetHandle* h) {
    if (h->c.normalize) return h->c.aux ? fleet_reset_kernel<true, true> : fleet_reset_kernel<true, false>;
    return h->c.aux ? fleet_reset_kernel<false, true> : fleet_reset_kernel<false, false>;
}

}  // namespace

extern "C" {

int fleet_abi_version(void) { return FLEETSTEP_ABI_VERSION; }

const char* fleet_last_error(const FleetHandle* h) { return h ? h->err.c_str() : "null handle"; }

int fleet_obs_dim(const FleetHandle* h) { return h ? h->D : FLEET_E_INVALID; }
int fleet_num_evs(const FleetHandle* h) { return h ? h->N : FLEET_E_INVALID; }
int fleet_num_envs(const FleetHandle* h) { return h ? h->E : FLEET_E_INVALID; }
int64_t fleet_launch_count(const FleetHandle* h) { return h ? h->launches : 0; }
int64_t fleet_device_bytes(const FleetHandle* h) { return h ? h->bytes : 0; }

int fleet_destroy(FleetHandle* h) {
    if (!h) return FLEET_E_INVALID;
    cudaSetDevice(h->device);
    for (void* ptr : h->allocs) cudaFree(ptr);
    delete h;
    return FLEET_OK;
}

int fleet_create(const FleetConsts* consts, const FleetTables* tb, int32_t num_envs, int32_t device, int64_t env_id_offset,
                 FleetHandle** out) {
    if (!out) return FLEET_E_INVALID;
    *out = nullptr;
    FleetHandle* h = new FleetHandle();
    *out = h;  // returned even on failure so that fleet_last_error works; caller destroys it
    if (!consts || !tb) return fail(h, FLEET_E_INVALID, "consts/tables is NULL");
    const FleetConsts& c = *consts;
    if (c.abi_version != FLEETSTEP_ABI_VERSION) return fail(h, FLEET_E_INVALID, "FleetConsts.abi_version mismatch");
    if (num_envs < 1 || c.num_evs < 1 || c.table_len < 3) return fail(h, FLEET_E_INVALID, "num_envs, num_evs must be >= 1 and table_len >= 3");
    if (!c.include_price)
        return fail(h, FLEET_E_INVALID, "include_price=False is unsupported: the reference raises KeyError('price_reward_curve') at ev_charger.py:155");
    if (c.normalize && c.include_pv && !c.include_building)
        return fail(h, FLEET_E_INVALID, "normalize_in_env with PV only is unsupported: the reference raises at oracle_normalization.py:120-121");
    if (!tb->there || !tb->time_left || !tb->soc_on_return || !tb->delu || !tb->tariff || !tb->price_reward_curve ||
        !tb->tariff_reward_curve || !tb->cal_sincos || !tb->hour || !tb->minute)
        return fail(h, FLEET_E_INVALID, "a mandatory table pointer is NULL");
    if (c.include_building && !tb->load) return fail(h, FLEET_E_INVALID, "include_building set but load table is NULL");
    if (c.include_pv && !tb->pv) return fail(h, FLEET_E_INVALID, "include_pv set but pv table is NULL");
    if (c.steps_per_hour < 1 || c.episode_steps < 1) return fail(h, FLEET_E_INVALID, "steps_per_hour and episode_steps must be >= 1");
    if ((double)(float)c.dt != c.dt) return fail(h, FLEET_E_INVALID, "dt is not exactly representable in float32 (hours_left is kept in float32)");
    if (c.episode_steps + 1 > 65535) return fail(h, FLEET_E_INVALID, "episode longer than 65535 steps is not supported yet");

    h->c = c; h->device = device; h->E = num_envs; h->N = c.num_evs; h->T = c.table_len;
    const int E = h->E, N = h->N, T = h->T;
    int Ha = 0, Hb = 0;
    h->D = obs_dim_of(c, &Ha, &Hb);

    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(h, FLEET_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce) + " (no CPU fallback exists)");
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, device));
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;

    // ---- build the HBM tables on the host (float64, reference operation order), then upload
    std::vector<EvRec> rec((size_t)T * N);
    for (int n = 0; n < N; n++) {
        for (int t = 0; t < T; t++) {
            const double tl = tb->time_left[(size_t)n * T + t];
            const float tlf = (float)tl;
            const double steps = tl / c.dt;
            if ((double)tlf != tl || steps != floor(steps) || steps > 4194304.0) {
                char buf[200];
                snprintf(buf, sizeof buf, "time_left[%d][%d]=%.17g is not an exact float32 multiple of dt=%.17g", n, t, tl, c.dt);
                return fail(h, FLEET_E_INVALID, buf);
            }
            EvRec& r = rec[(size_t)t * N + n];
            r.sr = tb->soc_on_return[(size_t)n * T + t];
            r.tl = tlf;
            r.there = tb->there[(size_t)n * T + t];
            r.there_prev = t > 0 ? tb->there[(size_t)n * T + t - 1] : 0;
            r.pad = 0;
        }
    }
    std::vector<StepRow> rows((size_t)T);
    const double spot_offset = c.fixed_markup / 1000;                                    // ev_charger.py:35
    for (int t = 0; t < T; t++) {
        StepRow& r = rows[t];
        double connected = 0;
        for (int n = 0; n < N; n++) connected += (double)tb->there[(size_t)n * T + t];   // ev_charger.py:138
        connected = connected > 1 ? connected : 1;                                       // :140
        const double pv_energy = tb->pv ? tb->pv[t] * c.dt : 0.0;                        // :133-136
        r.S = tb->delu[t] / 1000.0 + spot_offset;                                        // :145,149
        r.F_cr = -1 * c.price_multiplier * tb->price_reward_curve[t] / 1000;             // :154-156
        r.F_dr = -1 * c.price_multiplier * tb->tariff_reward_curve[t] / 1000;            // :204-206
        r.Rfac = c.discharging_eff * tb->tariff[t] / 1000 * (1 - c.feed_in_deduction);   // :196-199 (regrouped; cashflow tolerance)
        r.pv_share = pv_energy / connected;                                              // :142
        r.gml = c.grid_connection - ((c.include_building && tb->load) ? tb->load[t] : 0.0);  // load_calculation.py:93
        r.pvv = (c.include_pv && tb->pv) ? tb->pv[t] : 0.0;
        auto tf = [&](int tt) -> uint32_t {
            uint32_t f = 0;
            if (tb->hour[tt] == 14 && tb->minute[tt] == 45) f |= TF_TRIGGER;
            if (tb->hour[tt] > 11 && tb->hour[tt] < 15) f |= TF_LUNCH;
            return f;
        };
        r.flags = tf(t);
        r.flags_next = tf(t + 1 < T ? t + 1 : T - 1);
    }
    const int hdr_stride = ((Ha + Hb + 3) / 4) * 4 > 0 ? ((Ha + Hb + 3) / 4) * 4 : 4;
    std::vector<float> hdr((size_t)T * hdr_stride, 0.f);
    const bool norm = c.normalize != 0;
    for (int t = 0; t < T; t++) {
        float* o = hdr.data() + (size_t)t * hdr_stride;
        int q = 0;
        for (int k = 0; k <= c.price_lookahead; k++) {                                   // observer_bl_pv.py:63
            double v = (tb->delu[look_idx(c, *tb, t, k)] + c.fixed_markup) * c.variable_multiplier;
            if (norm) v = (v - c.min_price) / (c.max_price - c.min_price);               // oracle_normalization.py:70
            o[q++] = (float)v;
        }
        for (int k = 0; k <= c.price_lookahead; k++) {                                   // observer_bl_pv.py:64
            double v = tb->tariff[look_idx(c, *tb, t, k)] * (1 - c.feed_in_deduction);
            if (norm) v = (v - c.min_tariff) / (c.max_tariff - c.min_tariff);            // :72
            o[q++] = (float)v;
        }
        if (c.include_building)
            for (int k = 0; k <= c.bl_pv_lookahead; k++) {
                double v = tb->load[look_idx(c, *tb, t, k)];
                if (norm) v = v / c.max_building;
                o[q++] = (float)v;
            }
        if (c.include_pv)
            for (int k = 0; k <= c.bl_pv_lookahead; k++) {
                double v = tb->pv[look_idx(c, *tb, t, k)];
                if (norm) v = v / c.max_pv;
                o[q++] = (float)v;
            }
        if (c.aux) {
            double evse = c.evse_max_power;
            o[q++] = (float)(norm ? evse / c.evse_max_power : evse);                     // observer_bl_pv.py:93
            if (c.include_building) {
                double grid = c.grid_connection;                                         // :95
                double avail = grid - tb->load[t];                                       // :96
                if (c.include_pv) avail = avail + tb->pv[t];
                double poss = avail / ((double)N * evse);                                // :98
                if (poss > 1) poss = 1;
                if (norm) { grid = grid / c.grid_connection; avail = avail / c.grid_connection; poss = poss / 1; }
                o[q++] = (float)grid; o[q++] = (float)avail; o[q++] = (float)poss;
            }
            for (int k = 0; k < 6; k++) o[q++] = (float)tb->cal_sincos[(size_t)t * 6 + k];   // :100-107
        }
    }

    // ---- device allocations
    StepParams& p = h->p;
    memset(&p, 0, sizeof p);
    EvRec* d_rec; StepRow* d_rows; float* d_hdr;
    int rc;
    if ((rc = dev_alloc(h, &d_rec, rec.size(), false))) return rc;
    if ((rc = dev_alloc(h, &d_rows, rows.size(), false))) return rc;
    if ((rc = dev_alloc(h, &d_hdr, hdr.size(), false))) return rc;
    CUDA_TRY(h, cudaMemcpy(d_rec, rec.data(), rec.size() * sizeof(EvRec), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(d_rows, rows.data(), rows.size() * sizeof(StepRow), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(d_hdr, hdr.data(), hdr.size() * sizeof(float), cudaMemcpyHostToDevice));

    const int R = c.calc_degradation ? c.episode_steps + 1 : 2;
    const size_t EN = (size_t)E * N;
    if ((rc = dev_alloc(h, &p.env4, (size_t)E))) return rc;
    if ((rc = dev_alloc(h, &p.soc, EN))) return rc;
    if ((rc = dev_alloc(h, &p.hl, EN))) return rc;
    if ((rc = dev_alloc(h, &p.soh, EN))) return rc;
    if ((rc = dev_alloc(h, &p.hist, EN * (size_t)R))) return rc;
    if ((rc = dev_alloc(h, &p.tflip, EN))) return rc;
    if ((rc = dev_alloc(h, &p.n_flips, (size_t)4))) return rc;
    if ((rc = dev_alloc(h, &p.env_f64, (size_t)EF__COUNT * E))) return rc;
    if ((rc = dev_alloc(h, &p.rf_len, EN))) return rc;
    if ((rc = dev_alloc(h, &p.fd_cyc, EN))) return rc;
    if ((rc = dev_alloc(h, &p.life, EN))) return rc;
    if ((rc = dev_alloc(h, &p.n_cycles, EN))) return rc;
    if ((rc = dev_alloc(h, &p.last_deg, EN))) return rc;
    if ((rc = dev_alloc(h, &p.stats, (size_t)kStatStripes * FLEET_S__COUNT))) return rc;
    if ((rc = dev_alloc(h, &p.err_flags, (size_t)4))) return rc;

    p.E = E; p.N = N; p.T = T; p.R = R; p.L = c.episode_steps; p.D = h->D; p.Ha = Ha; p.Hb = Hb; p.hdr_stride = hdr_stride;
    p.B = N >= kThreads ? 1 : kThreads / N;
    p.is_ct = c.is_caretaker; p.calc_deg = c.calc_degradation; p.deg_mode = c.deg_mode; p.carry = c.carry_degradation_state;
    p.auto_reset = c.auto_reset; p.start_lo = c.start_lo; p.start_hi = c.start_hi; p.seed = c.seed;
    p.env_id_offset = env_id_offset;
    p.dt = c.dt; p.dt_f = (float)c.dt;
    p.P = c.obc_max_power < c.evse_max_power ? c.obc_max_power : c.evse_max_power;       // ev_charger.py:95
    p.eta_c = c.charging_eff; p.eta_d = c.discharging_eff; p.mult = c.variable_multiplier; p.cap0 = c.init_battery_cap;
    p.target = 1.0 * c.target_soc; p.target_lunch = c.target_soc_lunch; p.eps = c.soc_eps; p.def_soc = c.def_soc;
    p.min_lax = c.min_laxity; p.pen_inv = c.penalty_invalid_action; p.pen_oc = c.penalty_overcharging;
    p.clip_oc = c.clip_overcharging; p.pen_ovl = c.penalty_overloading; p.full_reward = c.fully_charged_reward;
    p.evse = c.evse_max_power; p.grid = c.grid_connection; p.init_soh = c.init_soh; p.lc_batt_cap = c.lc_batt_cap;
    p.hn_den = c.evse_max_power * c.charging_eff;                                        // observer_bl_pv.py:89
    p.price_mult = c.price_multiplier; p.temperature = c.temperature;
    p.max_tl = c.max_time_left; p.max_soc = c.target_soc;                                // oracle_normalization.py:34,49
    p.max_hn = (c.target_soc * c.init_battery_cap) / (c.evse_max_power * c.charging_eff);   // :50-51
    p.ev_rec = d_rec; p.step_row = d_rows; p.hdr = d_hdr;

    // shared memory of the step kernel: env scratch + contributions + sums + rainflow index stack
    const int slots = p.B * N;
    const int cap = c.episode_steps + 2;
    p.stack_u16 = cap > 255;
    size_t sm = smem_envs_bytes(p.B) + (size_t)kNQ * slots * 8 + (size_t)kNSum * p.B * 8;
    if (c.calc_degradation && c.deg_mode == FLEET_DEG_SEI) sm += (size_t)cap * kThreads * (p.stack_u16 ? 2 : 1);
    sm = (sm + 15) & ~(size_t)15;
    if ((int64_t)sm > (int64_t)h->max_smem_optin) {
        char buf[256];
        snprintf(buf, sizeof buf, "step kernel needs %zu bytes of shared memory (episode of %d steps) but the device allows %d; "
                 "long-episode global-memory rainflow stack is not implemented yet", sm, c.episode_steps, h->max_smem_optin);
        return fail(h, FLEET_E_INVALID, buf);
    }
    h->smem_step = sm;
    h->grid = (E + p.B - 1) / p.B;
    CUDA_TRY(h, cudaFuncSetAttribute(pick_step(h), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));

    init_state_kernel<<<(unsigned)((EN > (size_t)E ? EN : (size_t)E) + 255) / 256, 256>>>(p);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaDeviceSynchronize());
    return FLEET_OK;
}

int fleet_reset(FleetHandle* h, const int32_t* start_idx_dev, const uint8_t* mask_dev, float* obs_dev, void* stream) {
    if (!h) return FLEET_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    StepParams p = h->p;
    p.start_idx = start_idx_dev; p.mask = mask_dev; p.obs = obs_dev;
    pick_reset(h)<<<h->grid, kThreads, 0, (cudaStream_t)stream>>>(p);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return FLEET_OK;
}

int fleet_step(FleetHandle* h, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
               float* terminal_obs_dev, void* stream) {
    if (!h) return FLEET_E_INVALID;
    if (!actions_dev) return fail(h, FLEET_E_INVALID, "actions_dev is NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    StepParams p = h->p;
    p.actions = actions_dev; p.obs = obs_dev; p.reward = reward_dev; p.done = done_dev; p.terminal_obs = terminal_obs_dev;
    pick_step(h)<<<h->grid, kThreads, h->smem_step, (cudaStream_t)stream>>>(p);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, FLEET_E_CUDA, std::string("fleet_step launch: ") + cudaGetErrorString(e));
    return FLEET_OK;
}

int fleet_step_host(FleetHandle* h, const float* actions_host, float* obs_host, float* reward_host, uint8_t* done_host,
                    void* stream) {
    if (!h) return FLEET_E_INVALID;
    if (!actions_host) return fail(h, FLEET_E_INVALID, "actions_host is NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t EN = (size_t)h->E * h->N, ED = (size_t)h->E * h->D;
    int rc;
    if (!h->h_actions_dev) {
        if ((rc = dev_alloc(h, &h->h_actions_dev, EN, false))) return rc;
        if ((rc = dev_alloc(h, &h->h_obs_dev, ED, false))) return rc;
        if ((rc = dev_alloc(h, &h->h_reward_dev, (size_t)h->E, false))) return rc;
        if ((rc = dev_alloc(h, &h->h_done_dev, (size_t)h->E, false))) return rc;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->h_actions_dev, actions_host, EN * 4, cudaMemcpyHostToDevice, s));
    if ((rc = fleet_step(h, h->h_actions_dev, h->h_obs_dev, h->h_reward_dev, h->h_done_dev, nullptr, stream))) return rc;
    if (obs_host) CUDA_TRY(h, cudaMemcpyAsync(obs_host, h->h_obs_dev, ED * 4, cudaMemcpyDeviceToHost, s));
    if (reward_host) CUDA_TRY(h, cudaMemcpyAsync(reward_host, h->h_reward_dev, (size_t)h->E * 4, cudaMemcpyDeviceToHost, s));
    if (done_host) CUDA_TRY(h, cudaMemcpyAsync(done_host, h->h_done_dev, (size_t)h->E, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return FLEET_OK;
}

int fleet_set_next_start(FleetHandle* h, const int32_t* next_start_idx_dev) {
    if (!h) return FLEET_E_INVALID;
    h->p.next_start = next_start_idx_dev;
    return FLEET_OK;
}

int fleet_field_info(const FleetHandle* h, int32_t field, int32_t* elem_bytes, int64_t* count) {
    if (!h || field < 0 || field >= FLEET_F__COUNT) return FLEET_E_INVALID;
    const int64_t EN = (int64_t)h->E * h->N, E = h->E;
    int eb = 8; int64_t cnt = EN;
    switch (field) {
        case FLEET_F_SOC: case FLEET_F_SOC_DEG: case FLEET_F_SOH: case FLEET_F_TARGET_SOC: case FLEET_F_FD_CYC:
        case FLEET_F_LIFE: case FLEET_F_LAST_DEG: eb = 8; cnt = EN; break;
        case FLEET_F_HOURS_LEFT: eb = 4; cnt = EN; break;
        case FLEET_F_RF_LEN: case FLEET_F_N_CYCLES: eb = 4; cnt = EN; break;
        case FLEET_F_TIME_IDX: case FLEET_F_FINISH_IDX: case FLEET_F_EP_COUNT: eb = 4; cnt = E; break;
        default: eb = 8; cnt = E; break;
    }
    if (elem_bytes) *elem_bytes = eb;
    if (count) *count = cnt;
    return FLEET_OK;
}

static const void* plain_field_ptr(const FleetHandle* h, int32_t field) {
    const StepParams& p = h->p;
    switch (field) {
        case FLEET_F_SOC: return p.soc;
        case FLEET_F_HOURS_LEFT: return p.hl;
        case FLEET_F_SOH: return p.soh;
        case FLEET_F_REWARD64: return p.env_f64 + (size_t)EF_REWARD64 * h->E;
        case FLEET_F_CASHFLOW: return p.env_f64 + (size_t)EF_CASHFLOW * h->E;
        case FLEET_F_RF_LEN: return p.rf_len;
        case FLEET_F_FD_CYC: return p.fd_cyc;
        case FLEET_F_LIFE: return p.life;
        case FLEET_F_EP_RETURN: return p.env_f64 + (size_t)EF_EP_RETURN * h->E;
        case FLEET_F_LAST_EP_RETURN: return p.env_f64 + (size_t)EF_LAST_EP_RETURN * h->E;
        case FLEET_F_N_CYCLES: return p.n_cycles;
        case FLEET_F_LAST_DEG: return p.last_deg;
        case FLEET_F_OVERLOAD: return p.env_f64 + (size_t)EF_OVERLOAD * h->E;
        case FLEET_F_SOC_VIOL: return p.env_f64 + (size_t)EF_SOC_VIOL * h->E;
        default: return nullptr;
    }
}

int fleet_get_state(FleetHandle* h, int32_t field, void* dst_dev, void* stream) {
    if (!h || !dst_dev) return FLEET_E_INVALID;
    int32_t eb; int64_t cnt;
    if (fleet_field_info(h, field, &eb, &cnt)) return fail(h, FLEET_E_INVALID, "unknown field");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const void* src = plain_field_ptr(h, field);
    if (src) {
        CUDA_TRY(h, cudaMemcpyAsync(dst_dev, src, (size_t)cnt * eb, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    } else {
        gather_field_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->p, field, dst_dev);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
    }
    return FLEET_OK;
}

int fleet_set_state(FleetHandle* h, int32_t field, const void* src_dev, void* stream) {
    if (!h || !src_dev) return FLEET_E_INVALID;
    int32_t eb; int64_t cnt;
    if (fleet_field_info(h, field, &eb, &cnt)) return fail(h, FLEET_E_INVALID, "unknown field");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (field == FLEET_F_TARGET_SOC) {
        CUDA_TRY(h, cudaMemsetAsync(h->p.n_flips, 0, 4, (cudaStream_t)stream));
        scatter_target_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->p, (const double*)src_dev);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        return FLEET_OK;
    }
    void* dst = const_cast<void*>(plain_field_ptr(h, field));
    if (!dst) return fail(h, FLEET_E_INVALID, "field is not writable");
    CUDA_TRY(h, cudaMemcpyAsync(dst, src_dev, (size_t)cnt * eb, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FLEET_OK;
}

int fleet_get_stats(FleetHandle* h, double* dst_dev, void* stream) {
    if (!h || !dst_dev) return FLEET_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    reduce_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->p.stats, dst_dev);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return FLEET_OK;
}

int fleet_reset_stats(FleetHandle* h, void* stream) {
    if (!h) return FLEET_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemsetAsync(h->p.stats, 0, sizeof(double) * kStatStripes * FLEET_S__COUNT, (cudaStream_t)stream));
    return FLEET_OK;
}

int fleet_check_errors(FleetHandle* h, uint32_t* flags_host, void* stream) {
    if (!h) return FLEET_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    uint32_t f = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&f, h->p.err_flags, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(h, cudaStreamSynchronize((cudaStream_t)stream));
    CUDA_TRY(h, cudaMemsetAsync(h->p.err_flags, 0, 4, (cudaStream_t)stream));
    if (flags_host) *flags_host = f;
    if (f) {
        std::string m = "device error flags:";
        if (f & 1u) m += " NaN action (reference: TypeError ev_charger.py:209)";
        if (f & 2u) m += " negative battery life (rainflow_sei_degradation.py:179-180)";
        if (f & 4u) m += " DoD > 5 (rainflow_sei_degradation.py:164-167)";
        return fail(h, FLEET_E_STATE, m);
    }
    return FLEET_OK;
}

}  // extern "C"
