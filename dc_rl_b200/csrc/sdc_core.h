// sdc_core.h -- scalar (one-thread-per-env) part of the SustainDC step, shared by the CUDA kernels
// (sdc_kernels.cu) and by the host test harness (tests/hostsim), which compiles these same functions
// with g++ to unit-test the device logic without a GPU.
//
// Everything here is written from SURVEY.md Appendix A; each block cites the reference lines it
// implements (paths relative to the reference root).  Arithmetic is fp64 like the reference; the
// reward window is stored in fp32 (the 40 KB/env that dominates HBM traffic).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/sdc_b200.h"

#if defined(__CUDACC__)
#define SDC_HD __host__ __device__ __forceinline__
#define SDC_HDN __host__ __device__
#else
#define SDC_HD inline
#define SDC_HDN inline
#endif

#if !defined(__CUDACC__)
struct uint2 { unsigned int x, y; };      // host builds (tests/hostsim) only need the layout
#endif

namespace sdc {

constexpr int kQueueMax = 1000;       // sustaindc_env.py:148-149
constexpr int kListCap = 128;         // sorted quartile bracket (order statistics around a quartile rank)
constexpr int kWidenMargin = 4;       // safety net: extend a bracket side by one exact rank when fewer ranks than this remain
constexpr int kRecentreMargin = 10;   // ask the refresh pass to re-centre a bracket when fewer ranks than this remain
constexpr int kRecentreOk = 16;       // a re-centred bracket must have at least this many ranks on both sides
constexpr int kCollectAim = 110;      // half-width of the re-centring interval in average bracket gaps
constexpr int kCollectCap = 512;      // values a refresh can collect per bracket
constexpr int kWexpMax = 6;           // |log2| range of the adaptive interval width
constexpr int kTailCap = 128;         // values per tail band
constexpr int kBandTarget = 96;       // a refresh narrows the bands when one holds more values than this ...
constexpr int kBandSparse = 40;       // ... and widens them when both hold fewer than this (the bands are sorted: a step's cost does
                                      // not depend on their population, a wide band is re-centred less often)
constexpr int kAlphaOff = 12;         // narrowest tail band: alpha = 0.5 / 2^12 of the IQR
constexpr int kPassJobBytes = 272;     // size of one window-pass record (sdc_kernels.cu PassJob)
constexpr int kTailRetry = 200;       // steps of plain scans before another attempt at tail sets that did not fit
constexpr double kSpMin = 15.0, kSpMax = 21.6;   // utils/make_envs_pyenv.py:124-126

// ---- info table columns: keep in sync with dc_rl_b200/info_layout.py -------------------------
enum InfoCol {
    I_BAT_ACTION = 0, I_BAT_SOC, I_BAT_CO2, I_BAT_AVG_CI, I_BAT_E_WITHOUT, I_BAT_E_WITH, I_BAT_MAX_CAP,
    I_BAT_DCLOAD_MIN, I_BAT_DCLOAD_MAX,
    I_LS_ORIG_WORKLOAD, I_LS_SHIFTED_WORKLOAD, I_LS_ACTION, I_LS_NORM_LOAD_LEFT, I_LS_UNASSIGNED, I_LS_PENALTY_FLAG,
    I_LS_QUEUE_MAX_LEN, I_LS_TASKS_IN_QUEUE, I_LS_NORM_TASKS_IN_QUEUE, I_LS_TASKS_DROPPED, I_LS_CURRENT_HOUR,
    I_LS_TASKS_PROCESSED, I_LS_ENFORCED, I_LS_OLDEST_AGE, I_LS_AVG_AGE, I_LS_OVERDUE, I_LS_COMPUTED_TASKS,
    I_LS_HIST0, I_LS_HIST1, I_LS_HIST2, I_LS_HIST3, I_LS_HIST4,
    I_DC_ITE_KW, I_DC_CT_KW, I_DC_COMP_KW, I_DC_HVAC_KW, I_DC_TOTAL_KW, I_DC_SP_DELTA, I_DC_SP, I_DC_CPU_FRAC,
    I_DC_INT_TEMP, I_DC_AMBIENT, I_DC_POWER_LB, I_DC_POWER_UB, I_DC_CW_PUMP, I_DC_CT_PUMP, I_DC_WATER,
    I_OUTSIDE_TEMP, I_DAY, I_HOUR, I_NORM_CI,
    I_FORECAST0, I_FORECAST1, I_FORECAST2, I_FORECAST3, I_FORECAST4, I_FORECAST5, I_FORECAST6, I_FORECAST7,
    I_ISTERMINAL,
    I_COUNT
};
static_assert(I_COUNT == 59, "info layout");

enum Metric {
    M_ENERGY = 0, M_CO2, M_WATER, M_TASKS_IN_QUEUE, M_TASKS_DROPPED, M_ITE_KW, M_CT_KW, M_COMP_KW, M_HVAC_KW,
    M_STEPS, M_REWARD_SUM, M_EPISODES, M_REWARD_LS, M_REWARD_DC, M_OVERDUE, M_TOTAL_KW
};

// ---- device-resident tables and per-env state (structure of arrays) --------------------------
struct LocTables {
    const double* workload; const uint8_t* ns; const uint8_t* sh; const double* ci; const double* ci_min30;
    const double* ci_max30; const double* temp_base; const double* wetb_base;
};

struct State {
    int32_t n_envs, ep_len, hist_cap, ls_mask, win_len, n_loc, n_cfg, pad0;
    int32_t reward_kind[3];        // SDC_R_* per agent (ls, dc, bat)
    int32_t append_history;        // agent_ls uses default_ls_reward: the step's energy enters the reward window
    const LocTables* loc;          // [n_loc]
    const sdc_dc_params* dc;       // [n_cfg]
    const double* hour_cos;        // [96]
    const double* hour_sin;        // [96]
    const uint8_t* loc_id;         // [N]
    const uint8_t* cfg_id;         // [N]
    const int16_t* day_lo;         // [N]
    const int16_t* day_hi;         // [N]
    const uint64_t* seed;          // [N]
    uint32_t* episode;             // [N] episodes started (RNG counter)
    // exogenous
    int32_t* t;                    // trace index == quarter-hour stamp day*96+hour*4
    int32_t* t0;                   // episode start index
    int32_t* step_in_ep;
    double* ci_min; double* ci_max;          // CI_Manager 30-day normalisation    managers.py:435-437
    double* t_min; double* t_max;            // Weather_Manager normalisation      managers.py:606-608
    double* weather;               // [N][2 buffers][2][win_len]: realised dry bulb, wet bulb of the current and the staged episode
    uint8_t* cur_buf;              // [N] which buffer holds the current episode (a reset flips it instead of copying 11 KB)
    // load shifting queue as a ring of per-quarter-hour task counts
    int32_t* ls_head; int32_t* ls_len; int32_t* ls_sum;
    uint16_t* ls_bins;             // [N][4] tasks aged [0,6) [6,12) [12,18) [18,24) hours
    uint8_t* ls_ring;              // [N][ls_mask+1]
    // data centre
    double* setpoint; int32_t* dc_run; int32_t* dc_scale; int8_t* dc_last;
    // battery
    double* bat_load;
    // reward window + quartile brackets
    float* hist;                   // [N][hist_cap] window samples RELATIVE to hist_ref (fp32)
    double* hist_ref;              // [N] the env's first sample (fp64): what the fp32 window values are measured from
    int32_t* hist_len; int32_t* hist_head;
    float* qlist;                  // [N][2][kListCap] sorted order statistics around the quartile ranks
    int32_t* q_a;                  // [N][2] rank of qlist[.][0]
    int32_t* q_m;                  // [N][2] elements in the list
    // incremental reward normaliser: window moments about c0, tail multisets, adaptive refresh parameters
    double* mom_s1; double* mom_s2; double* mom_c0;   // [N]
    float* tails;                  // [N][2][kTailCap] the two tail bands, each SORTED ascending (tail_ptr)
    int32_t* tail_n;               // [N][2] band sizes, -1 = no valid bands
    int32_t* tail_nb;              // [N][2] band values beyond the fence: the first nb of the lower band, the last nb of the upper one
    double* tail_bs;               // [N][2][2] sum / sum of squares about c0 of those nb values
    float* tail_thr;               // [N][4] TL, TH (inner thresholds), TL2, TH2 (outer thresholds)
    int32_t* agg_n; double* agg_s; // [N][2] count, [N][2][2] sum / sum of squares about c0 of the values beyond TL2 / TH2
    uint32_t* fast_cfg;            // [N] byte 0 tail slack exponent, 1 retry countdown, 2/3 interval width exponents (int8)
    int32_t* err;                  // [N] SDC_F_* bits
    // staged next episodes: written by the look-ahead generation (with the reset observation) or by sdc_stage_episode
    uint8_t* pend_valid;           // [N] bit 0: the other weather buffer + pend_day / hour / tmin / tmax hold the next episode;
                                   //     bit 1: pend_obs holds its reset observation
    int32_t* pend_day; int32_t* pend_hour; double* pend_tmin; double* pend_tmax;
    float* pend_obs;               // [N][3][26]
};

// The env's current / staged weather window: [0, win_len) dry bulb, [win_len, 2 win_len) wet bulb.
SDC_HD double* weather_buf(const State& S, int env, int which) { return S.weather + ((size_t)env * 2 + which) * 2 * S.win_len; }
SDC_HD double* weather_cur(const State& S, int env) { return weather_buf(S, env, S.cur_buf[env]); }
SDC_HD double* weather_pend(const State& S, int env) { return weather_buf(S, env, S.cur_buf[env] ^ 1); }

// Where the shared tables are read from (k_step points these at a shared-memory copy).
struct Tables { const LocTables* loc; const sdc_dc_params* dc; };

// Error flags are OR-ed atomically on the device: a reset worker and the env's step warp may both flag an env.
SDC_HD void flag_error(const State& S, int env, int bits) {
#if defined(__CUDA_ARCH__)
    atomicOr(S.err + env, bits);
#else
    S.err[env] |= bits;
#endif
}

SDC_HD double round_dec(double x, double scale) { return rint(x * scale) / scale; }   // np.round(x, d)
SDC_HD double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }
SDC_HD double clampd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// ---- observation features (sustaindc_env.py:266-433) -----------------------------------------
// OLS slope of y[0..n) against 0..n-1 (what np.polyfit(range(n), y, 1)[0] solves).
template <int N>
SDC_HD double ols_slope(const double* y) {
    const double xm = (N - 1) * 0.5;
    double sxy = 0.0, sxx = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double dx = i - xm;
        sxy += dx * y[i];
        sxx += dx * dx;
    }
    return sxy / sxx;
}

// mean, std, (cur-mean)/(std+1e-8), first-peak/len, first-valley/len of `v[0..N)` given `cur`
// (sustaindc_env.py:273-286 and 366-384; np.gradient of [cur, v...]).
template <int N>
SDC_HD void trend_features(double cur, const double* v, double* out5) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += v[i];
    const double mean = s / N;
    double ss = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) { const double d = v[i] - mean; ss += d * d; }
    const double sd = sqrt(ss / N);
    // gradient over the N+1 points x[0]=cur, x[i]=v[i-1]
    double g_prev = v[0] - cur;                      // one-sided at the start
    int peak = N, valley = N;
#pragma unroll
    for (int i = 1; i <= N; ++i) {
        const double xm1 = (i == 1) ? cur : v[i - 2];
        const double g = (i < N) ? (v[i] - xm1) * 0.5 : (v[N - 1] - v[N - 2]);   // central / one-sided end
        if (peak == N && g_prev > 0.0 && g <= 0.0) peak = i - 1;
        if (valley == N && g_prev < 0.0 && g >= 0.0) valley = i - 1;
        g_prev = g;
    }
    out5[0] = mean; out5[1] = sd; out5[2] = (cur - mean) / (sd + 1e-8);
    out5[3] = (double)peak / N; out5[4] = (double)valley / N;
}

struct LsStats { double oldest, avg, norm_q, hist[5]; };

// Per-episode normalisation constants of an env (CI_Manager / Weather_Manager 30-day min-max) and its weather window,
// addressed by trace index: wrel[t] = dry bulb, wrel[win_len + t] = wet bulb at index t (wrel = window start - t0).
struct Norms { double cmin, crng, tmin, trng; const double* wrel; };

// The observation features come in two independent halves (carbon intensity of the location, temperature of the env's
// weather window); build_obs computes both in one thread, the look-ahead episode generation one half each on two threads.
struct CiFeat { double x0, f[7]; };          // norm CI at t + the 7 features of sustaindc_env.py:266-300
struct TempFeat { double nt0, nt1, f[6]; };  // norm temperature at t, t+1 + the 6 features of :366-384

// Min-max normalisation by a multiplication with the reciprocal range (two divisions instead of 42): the quotient may
// differ from the reference's in the last fp64 bit, which survives the fp32 cast of an observation with probability
// ~2e-9 per value (the golden replays stay bit-identical).
SDC_HDN void ci_features(const LocTables& L, int t, double cmin, double crng, CiFeat& out) {
    const int tp = t >= 16 ? t - 16 : 0;              // start of the 16-sample past window (empty before t = 16)
    double ci_raw[25];
#pragma unroll
    for (int i = 0; i < 16; ++i) ci_raw[i] = L.ci[tp + i];
#pragma unroll
    for (int i = 0; i < 9; ++i) ci_raw[16 + i] = L.ci[t + i];
    const double inv_crng = 1.0 / crng;
    double x[9];                                       // x[0..8] = cur, fut[8]
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = (ci_raw[16 + i] - cmin) * inv_crng;
    double sm[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sm[i] = (((x[i] + x[i + 1]) + x[i + 2]) + x[i + 3]) / 4;
    out.f[0] = ols_slope<6>(sm);
    double p[17];
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = (ci_raw[i] - cmin) * inv_crng;
    p[16] = x[0];
    double smp[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) smp[i] = (((p[i] + p[i + 1]) + p[i + 2]) + p[i + 3]) / 4;
    // empty past window at the start of the year -> zero slope
    out.f[1] = t >= 16 ? ols_slope<14>(smp) : 0.0;
    trend_features<8>(x[0], x + 1, out.f + 2);
    out.x0 = x[0];
}
// wtemp: the env's dry-bulb window at trace index t (17 samples are read)
SDC_HDN void temp_features(const double* wtemp, double tmin, double trng, TempFeat& out) {
    double wt_raw[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) wt_raw[i] = wtemp[i];
    const double inv_trng = 1.0 / trng;
    double nt[17];                                     // nt[0..16] = normT[t..t+16]
#pragma unroll
    for (int i = 0; i < 17; ++i) nt[i] = (wt_raw[i] - tmin) * inv_trng;
    out.f[0] = ols_slope<17>(nt);
    trend_features<16>(nt[0], nt + 1, out.f + 1);
    out.nt0 = nt[0]; out.nt1 = nt[1];
}
// The three rows.  Sink: void operator()(int agent, int idx, float v).  Layouts: SURVEY.md A.6 / sustaindc_env.py:302-433.
template <class Sink>
SDC_HDN void emit_obs_rows(double cos_h, double sin_h, double w, double w_next, const CiFeat& ci, const TempFeat& tf, const LsStats& ls,
                           double soc, Sink& sink) {
    // agent_ls [26]
    int k = 0;
    sink(0, k++, (float)cos_h); sink(0, k++, (float)sin_h); sink(0, k++, (float)ci.x0);
#pragma unroll
    for (int i = 0; i < 7; ++i) sink(0, k++, (float)ci.f[i]);
    sink(0, k++, (float)ls.oldest); sink(0, k++, (float)ls.avg); sink(0, k++, (float)ls.norm_q);
    sink(0, k++, (float)w); sink(0, k++, (float)tf.nt0);
#pragma unroll
    for (int i = 0; i < 6; ++i) sink(0, k++, (float)tf.f[i]);
#pragma unroll
    for (int i = 0; i < 5; ++i) sink(0, k++, (float)ls.hist[i]);
    // agent_dc [14] (+ zero padding to 26)
    k = 0;
    sink(1, k++, (float)cos_h); sink(1, k++, (float)sin_h); sink(1, k++, (float)ci.x0);
#pragma unroll
    for (int i = 0; i < 7; ++i) sink(1, k++, (float)ci.f[i]);
    sink(1, k++, (float)w); sink(1, k++, (float)w_next); sink(1, k++, (float)tf.nt0); sink(1, k++, (float)tf.nt1);
    for (; k < SDC_OBS_DIM; ++k) sink(1, k, 0.0f);
    // agent_bat [13]
    k = 0;
    sink(2, k++, (float)cos_h); sink(2, k++, (float)sin_h); sink(2, k++, (float)ci.x0);
#pragma unroll
    for (int i = 0; i < 7; ++i) sink(2, k++, (float)ci.f[i]);
    sink(2, k++, (float)w); sink(2, k++, (float)tf.nt0); sink(2, k++, (float)soc);
    for (; k < SDC_OBS_DIM; ++k) sink(2, k, 0.0f);
}

// Builds the three observations at trace index t (one thread).
// All trace reads are issued first, in one straight-line batch, so that their memory latencies overlap: a dependent
// global read costs ~1-2 us under load (profiles/r01_summary.md).
template <class Sink>
SDC_HDN void build_obs(const State& S, const Tables& T, int env, int t, const LsStats& ls, double soc, const Norms& nm, Sink& sink, int loc = -1) {
    const LocTables& L = T.loc[loc >= 0 ? loc : S.loc_id[env]];     // callers that already know the location pass it (one dependent load less)
    const double w = L.workload[t], w_next = L.workload[t + 1];
    const int hq = t % 96;
    const double cos_h = S.hour_cos[hq], sin_h = S.hour_sin[hq];
    CiFeat ci;
    TempFeat tf;
    ci_features(L, t, nm.cmin, nm.crng, ci);
    temp_features(nm.wrel + t, nm.tmin, nm.trng, tf);
    emit_obs_rows(cos_h, sin_h, w, w_next, ci, tf, ls, soc, sink);
}

// ---- EnergyPlus-style electric chiller (envs/datacenter.py:356-429) ---------------------------
SDC_HD double chiller_power(double max_cap, double load, double ambient) {
    const double d_t = (ambient - 35.0) / 2.778 - (6.67 - 35.0);
    const double rat = 0.94483600 + (-0.05700880) * d_t + 0.00185486 * (d_t * d_t);
    const double avail = (rat != 0.0) ? max_cap * rat : 0.0;
    const double fpr = 2.333 + (-1.975) * rat + 0.6121 * (rat * rat);
    const double la = (avail > 0.0) ? load / avail : 0.0;
    const double plr = (avail > 0.0) ? fmax(0.05, fmin(la, 1.0)) : 0.0;
    const double ffl = 0.03303 + 0.6852 * plr + 0.2818 * (plr * plr);
    double opl = 0.0;
    if (avail > 0.0) opl = (la < 0.05) ? la : plr;
    const double frac = (opl < 0.05) ? fmin(1.0, opl / 0.05) : 1.0;
    const double power = ffl * fpr * avail / 3.0 * frac;
    return (opl > 0.0) ? power : 0.0;
}

// Result of the scalar phase of one env-step that the reward phase needs.
struct StepResult {
    double energy;        // bat_total_energy_with_battery_KWh -> appended to the reward window
    double nci_next;      // norm_CI of the next step (ci_i_future[0], sustaindc_env.py:681)
    double ls_penalty;    // (-0.3*sqrt(overdue)+0.3) + (-0.1*oldest_age)   reward_creator.py:67-82
    int terminal;
    int step_after;       // steps taken in the episode including this one
    // logger metrics of this step
    double co2, water, ite_kw, ct_kw, comp_kw, hvac_kw, total_kw;
    int tasks_in_queue, tasks_dropped, overdue;
    // reward window cursor and the value the new sample evicts (read early, with the other state)
    int hist_len, hist_head;
    float evicted;
};

// What the observation builder needs from the scalar phase (the observations are built after the normaliser).
struct ObsDeferred { LsStats ls; double soc; Norms nm; int tn; int loc; };

// The part of StepResult the reward needs (kept small: it lives in registers across the window passes).
struct RewardInputs { double energy, nci_next, ls_penalty; };
// What the alternate reward methods read from the step (utils/reward_creator.py:133-318).
struct AltInputs { double ite_kw, total_kw, water; int hour_q; };      // hour_q: quarter-hours since midnight at the new time

SDC_HD bool any_alt_reward(const State& S) { return (S.reward_kind[0] | S.reward_kind[1] | S.reward_kind[2]) > SDC_R_DEFAULT_DC; }
// Alternate (stateless) reward methods.  tou_reward indexes its price table with the float hour and raises KeyError
// off the full hour (three steps out of four): flagged here, priced at the hour's tariff.
SDC_HD double alt_reward(int kind, double energy, const AltInputs& ai, int& err) {
    switch (kind) {
        case SDC_R_TOU: {
            const int h = ai.hour_q >> 2;
            if (ai.hour_q & 3) err |= SDC_F_REWARD_DOMAIN;
            const double price = (h < 6 || h >= 22) ? 0.25 : (h < 11 ? 0.41 : (h < 16 ? 0.30 : 0.27));   // :169-192
            return -1.0 * energy * price;
        }
        case SDC_R_ENERGY_EFFICIENCY: return ai.ite_kw / ai.total_kw;
        case SDC_R_PUE: return ai.ite_kw != 0.0 ? -fabs(ai.total_kw / ai.ite_kw - 1.0) : -INFINITY;
        case SDC_R_WATER: return -0.01 * ai.water;
        default: return 0.0;                                             // SDC_R_CUSTOM
    }
}
// The three agents' alternate rewards of one step (entries of default-method agents are unused).
SDC_HD void alt_rewards(const State& S, int env, double energy, const AltInputs& ai, float* alt3) {
    int err = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) alt3[a] = S.reward_kind[a] > SDC_R_DEFAULT_DC ? (float)alt_reward(S.reward_kind[a], energy, ai, err) : 0.f;
    if (err) flag_error(S, env, err);
}

// One env-step of the three sub-envs + managers + observations + info (everything except the
// reward normaliser).  InfoSink: void operator()(int col, float v).
template <class InfoSink>
SDC_HDN void physics_step(const State& S, const Tables& T, int env, int a_ls, int a_dc, int a_bat, InfoSink& info,
                          StepResult& out, ObsDeferred& od, bool load_evicted = true) {
    // ---- level 1: every per-env scalar, issued back to back (no stores in between) ----
    const int t = S.t[env], t0 = S.t0[env], step0 = S.step_in_ep[env];
    int head = S.ls_head[env], len = S.ls_len[env], sum = S.ls_sum[env];
    int b0 = S.ls_bins[env * 4 + 0], b1 = S.ls_bins[env * 4 + 1], b2 = S.ls_bins[env * 4 + 2], b3 = S.ls_bins[env * 4 + 3];
    const double sp0 = S.setpoint[env];
    int run = S.dc_run[env], scale = S.dc_scale[env];
    const int last = S.dc_last[env];                              // 2 == None (after reset, dc_gym.py:115)
    double b = S.bat_load[env];
    Norms nm;
    nm.cmin = S.ci_min[env]; nm.crng = S.ci_max[env] - nm.cmin;
    nm.tmin = S.t_min[env]; nm.trng = S.t_max[env] - nm.tmin; nm.wrel = weather_cur(S, env) - t0;
    const int h_len = S.hist_len[env], h_head = S.hist_head[env];
    const int loc = S.loc_id[env];
    const LocTables& L = T.loc[loc];
    const sdc_dc_params& P = T.dc[S.cfg_id[env]];
    // ---- level 2: reads whose address depends on level 1 ----
    const int q = t;                                     // quarter-hour stamp == trace index
    const int mask = S.ls_mask;
    uint8_t* ring = S.ls_ring + (size_t)env * (S.ls_mask + 1);
    const int m1 = ring[(q - 24) & mask], m2 = ring[(q - 48) & mask], m3 = ring[(q - 72) & mask], m4r = ring[(q - 96) & mask];
    const double w = L.workload[t];
    const int ns = L.ns[t], sh = L.sh[t];
    const double ambient = nm.wrel[t], wet_bulb = nm.wrel[S.win_len + t], outside_next = nm.wrel[t + 1];
    const double ci_now = L.ci[t];
    double ci_fut[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ci_fut[i] = L.ci[t + 2 + i];
    out.hist_len = h_len; out.hist_head = h_head;
    // The sample the window append will evict: a DRAM miss whose value this function would have to STORE into `out` right
    // away (an output struct passed by reference lives in memory) -- the CUDA kernel reads it after the physics instead, from
    // the L2 line its prefetch brought in.
    out.evicted = load_evicted ? S.hist[(size_t)env * S.hist_cap + h_head] : 0.f;
    int err = 0;
    if (t + 18 > SDC_YEAR_STEPS) err |= SDC_F_TRACE_DOMAIN;    // the reference crashes here
    // Out-of-range action ids saturate to {0, 2}.  All branches below test the RAW ids and the reported
    // ids are clamped in fp32: nvcc 12.9 ptxas for sm_100a miscompiles `x = clamp(x,0,2); if (x == 0) .. else
    // if (x == 2) ..` into VIMNMX.RELU with a predicate output that does not mean `x == 2` (reproduced on
    // B200 with scratch-size kernels, see DESIGN.md "toolchain notes"); build() rejects SASS of that form.
    const bool ls_defer = a_ls <= 0, ls_process = a_ls >= 2;
    const bool dc_down = a_dc <= 0, dc_up = a_dc >= 2;
    const bool bat_charge = a_bat <= 0, bat_discharge = a_bat == 1;
    const float a_ls_f = fminf(fmaxf((float)a_ls, 0.f), 2.f), a_bat_f = fminf(fmaxf((float)a_bat, 0.f), 2.f);

    // ------------------------------------------------------------------------------------------
    // Load shifting (envs/carbon_ls.py:172-324).  The FIFO is a ring of task counts per stamp.
    // ------------------------------------------------------------------------------------------
    if (w < 0.0 || w > 1.0) err |= SDC_F_WORKLOAD_RANGE;
    int m4 = 0;
    if (len > 0) {
        // the clock moved one quarter-hour: tasks reaching an age of exactly 6/12/18/24 h change bin
        m4 = m4r;
        b0 -= m1; b1 += m1 - m2; b2 += m2 - m3; b3 += m3 - m4;
    }
    int b4 = len - (b0 + b1 + b2 + b3);
    const int over = b4 - m4;                            // age > 24 h, strict          carbon_ls.py:208-209
    // pops the `m` oldest tasks
    auto pop_oldest = [&](int m) {
        while (m > 0 && len > 0) {
            const int c = ring[head & mask];
            const int take = c < m ? c : m;
            ring[head & mask] = (uint8_t)(c - take);
            m -= take; len -= take; sum -= take * head;
            const int age_bin = (q - head) / 24;
            if (age_bin == 0) b0 -= take; else if (age_bin == 1) b1 -= take; else if (age_bin == 2) b2 -= take;
            else if (age_bin == 3) b3 -= take;
            if (c == take && len > 0) {
                if (head >= q) { err |= SDC_F_BRACKET; len = 0; break; }      // corrupted ring (bug guard)
                do { ++head; } while (ring[head & mask] == 0 && head < q);
            }
        }
    };
    int cap = 90 - (ns + sh);
    int otp = 0;
    if (cap > 0 && over > 0) { otp = over < cap ? over : cap; pop_oldest(otp); }     // carbon_ls.py:212-226
    cap = 90 - (ns + sh + otp);
    int dropped = 0, proc = 0;
    double util;
    if (ls_defer) {                                                                   // defer   :231-242
        int add = kQueueMax - len; add = sh < add ? sh : add;
        dropped = sh - add;
        if (add > 0) {
            if (len == 0) head = q;
            ring[q & mask] = (uint8_t)(ring[q & mask] + add);
            len += add; sum += add * q; b0 += add;
        }
        util = (double)(otp + (sh - add)) / 100;
    } else if (ls_process) {                                                          // process :244-264
        if (cap >= 1) {
            proc = sh < cap ? sh : cap; proc = proc < len ? proc : len;
            pop_oldest(proc);
            util = (double)(sh + proc + otp) / 100;
        } else {
            util = (double)(sh + otp) / 100;
        }
    } else {
        util = (double)(sh + otp) / 100;
    }
    util += (double)ns / 100;                                                         // :275-276
    LsStats ls;
    const double inv_len = 1.0 / (double)(len > 1 ? len : 1);
    if (len > 0) {
        ls.oldest = (double)(q - head) / 96.0;              // (x / 4) / 24 of an integer x: one correctly rounded division
        ls.avg = (((double)((long long)len * q - sum) * 0.25) * inv_len) * (1.0 / 24.0);
    } else {
        ls.oldest = 0.0; ls.avg = 0.0;
    }
    b4 = len - (b0 + b1 + b2 + b3);
    // (fp64 divisions cost ~25 dependent instructions each; quotients that only reach fp32 outputs are formed with one
    //  reciprocal instead: <= 1 ulp of fp64 off, which survives the fp32 cast with probability ~2e-9 per value)
    ls.hist[0] = (double)b0 * inv_len; ls.hist[1] = (double)b1 * inv_len; ls.hist[2] = (double)b2 * inv_len; ls.hist[3] = (double)b3 * inv_len;
    ls.hist[4] = b4 > 0 ? 1.0 : 0.0;                                                  // :63-73
    ls.norm_q = (double)len / kQueueMax;
    S.ls_head[env] = head; S.ls_len[env] = len; S.ls_sum[env] = sum;
    S.ls_bins[env * 4 + 0] = (uint16_t)b0; S.ls_bins[env * 4 + 1] = (uint16_t)b1;
    S.ls_bins[env * 4 + 2] = (uint16_t)b2; S.ls_bins[env * 4 + 3] = (uint16_t)b3;
    const double hour_now = (double)(t % 96) * 0.25;
    info(I_LS_ORIG_WORKLOAD, (float)w); info(I_LS_SHIFTED_WORKLOAD, (float)util); info(I_LS_ACTION, a_ls_f);
    info(I_LS_NORM_LOAD_LEFT, 0.f); info(I_LS_UNASSIGNED, 0.f); info(I_LS_PENALTY_FLAG, 0.f);
    info(I_LS_QUEUE_MAX_LEN, (float)kQueueMax); info(I_LS_TASKS_IN_QUEUE, (float)len);
    info(I_LS_NORM_TASKS_IN_QUEUE, (float)ls.norm_q); info(I_LS_TASKS_DROPPED, (float)dropped);
    info(I_LS_CURRENT_HOUR, (float)hour_now); info(I_LS_TASKS_PROCESSED, (float)proc); info(I_LS_ENFORCED, 0.f);
    info(I_LS_OLDEST_AGE, (float)ls.oldest); info(I_LS_AVG_AGE, (float)ls.avg); info(I_LS_OVERDUE, (float)over);
    info(I_LS_COMPUTED_TASKS, (float)(int)(util * 100));
#pragma unroll
    for (int i = 0; i < 5; ++i) info(I_LS_HIST0 + i, (float)ls.hist[i]);

    // ------------------------------------------------------------------------------------------
    // Data centre (envs/dc_gym.py:142-237; envs/datacenter.py)
    // ------------------------------------------------------------------------------------------
    if (!(util >= 0.0 && util <= 1.0)) { err |= SDC_F_CPU_LOAD_RANGE; util = clampd(util, 0.0, 1.0); }
    const int delta = dc_down ? -1 : (dc_up ? 1 : 0);             // {0:-1, 1:0, 2:+1}  make_envs_pyenv.py:127-131
    if (delta == last && !dc_down) { run += 1; } else { run = 1; scale = 1; }    // dc_gym.py:163-167
    if (run > 3) scale += 1;                                                          // :170-171
    double sp = sp0 + (double)(delta * scale);
    sp = fmax(fmin(sp, kSpMax), kSpMin);                                              // :173-174
    S.setpoint[env] = sp; S.dc_run[env] = run; S.dc_scale[env] = scale; S.dc_last[env] = (int8_t)delta;
    const double load_pct = util * 100;
    double p_it = 0.0, sum_out = 0.0;
    const double load_cpu = P.shift_cpu * (load_pct / 100), load_fan = P.shift_fan * (load_pct / 20);
    const double k_out = 1.918 / (P.c_air * P.rho_air * 0.526);
    // datacenter.py:157-181,250-317 per rack class, two classes per trip (the second one a zero-weight copy of the first when
    // the count is odd): each class is a ~150-instruction dependent chain through log / exp, and two independent chains
    // in flight nearly halve the time a lane spends here
    for (int c = 0; c < P.n_classes; c += 2) {
        const int c1 = c + 1 < P.n_classes ? c + 1 : c;
        const double w1 = c + 1 < P.n_classes ? P.cls_mult[c1] : 0.0;
        const double t_in0 = P.cls_supply[c] + sp, t_in1 = P.cls_supply[c1] + sp;
        const double ratio0 = ((P.m_cpu + 0.05) * t_in0 + P.c_cpu) + load_cpu, ratio1 = ((P.m_cpu + 0.05) * t_in1 + P.c_cpu) + load_cpu;
        const double n0 = P.cls_ncpu[c], n1 = P.cls_ncpu[c1];
        const double pc0 = fmax(P.cls_idle[c], P.cls_full[c] * ratio0) * n0, pc1 = fmax(P.cls_idle[c1], P.cls_full[c1] * ratio1) * n1;
        const double v0 = (P.m_fan * 10 * t_in0 + P.c_fan * 5) + load_fan, v1 = (P.m_fan * 10 * t_in1 + P.c_fan * 5) + load_fan;
        const double pf0 = (P.itfan_ref_p * (v0 / P.itfan_ref_v_ratio)) * n0, pf1 = (P.itfan_ref_p * (v1 / P.itfan_ref_v_ratio)) * n1;
        const double vf0 = (P.itfan_full_load_v * v0) * n0, vf1 = (P.itfan_full_load_v * v1) * n1;
        // 1.918 P^1.096 / (c_air rho V^0.824 0.526) as ONE exponential of a difference of logarithms: a few ulp instead of
        // pow's < 1 ulp (invisible after the fp32 cast of every output that depends on it, ~1e-15 relative on the energy)
        // for a quarter of the instructions of two pow calls and a division
        const double lp0 = log(pc0 + pf0), lp1 = log(pc1 + pf1);
        const double lv0 = log(vf0), lv1 = log(vf1);
        const double r0 = exp(1.096 * lp0 - 0.824 * lv0), r1 = exp(1.096 * lp1 - 0.824 * lv1);
        const double t_out0 = t_in0 + k_out * r0 + (-14.01), t_out1 = t_in1 + k_out * r1 + (-14.01);
        if (t_out0 - t_in0 < 2.0 || t_out1 - t_in1 < 2.0) err |= SDC_F_OUTLET_DELTA;  // :295-300
        p_it += P.cls_mult[c] * (pc0 + pf0) + w1 * (pc1 + pf1);
        sum_out += P.cls_mult[c] * t_out0 + w1 * t_out1;
    }
    const double mean_out = sum_out / P.n_racks;
    const double t_ret = P.ret_mean + mean_out;                                       // :531-541
    // HVAC (datacenter.py:432-474)
    const double m_sys = P.rho_air * P.crac_supply_flow_pu * p_it;
    const double q_crac = m_sys * P.c_air * fmax(0.0, t_ret - sp);
    const double comp = chiller_power(P.ct_fan_ref_p, q_crac, ambient);
    double ct = 0.0;
    if (!(ambient < 5.0)) {
        const double dlt = fmax(50 - (ambient - sp), 1.0);
        const double v_air = q_crac / (P.c_air * dlt) / P.rho_air;
        const double r = fmin(v_air / P.ctafr, 1.0);
        ct = P.ct_fan_ref_p * (r * r * r);
    }
    // cooling-tower water (datacenter.py:325-353)
    double wtr = 0.044 * wet_bulb + (0.3528 * (t_ret - sp) + 0.101);
    wtr = fmax(wtr, 0.0);
    wtr += wtr * 0.01;
    const double water = round_dec((wtr * 1000) / 4, 1e4);
    const double total_kw = (p_it + ct + comp) / 1e3;
    info(I_DC_ITE_KW, (float)(p_it * 1e-3)); info(I_DC_CT_KW, (float)(ct * 1e-3)); info(I_DC_COMP_KW, (float)(comp * 1e-3));
    info(I_DC_HVAC_KW, (float)((ct + comp) * 1e-3)); info(I_DC_TOTAL_KW, (float)total_kw);
    info(I_DC_SP_DELTA, (float)delta); info(I_DC_SP, (float)sp); info(I_DC_CPU_FRAC, (float)util);
    info(I_DC_INT_TEMP, (float)mean_out); info(I_DC_AMBIENT, (float)ambient);
    info(I_DC_POWER_LB, (float)P.power_lb_kw); info(I_DC_POWER_UB, (float)P.power_ub_kw);
    info(I_DC_CW_PUMP, (float)P.cw_pump_w); info(I_DC_CT_PUMP, (float)P.ct_pump_w); info(I_DC_WATER, (float)water);

    // ------------------------------------------------------------------------------------------
    // Battery (envs/bat_env_fwd_view.py:84-126,194-284; envs/battery_model.py:94-139)
    // ------------------------------------------------------------------------------------------
    const double dcl = total_kw / 1e3;                           // MW                 sustaindc_env.py:652
    const double capb = P.bat_capacity_mwh, inv_capb = 1.0 / capb;
    const double soc0 = b * inv_capb;
    double energy, co2;
    if (bat_charge) {                                            // charge
        const double t_u = round_dec(0.5 * (1 - sigmoid(10 * (soc0 - 0.5))), 1e4) * 15 / 60;
        const double max_c = fmin(capb * 0.1, (capb - b) / (t_u - (-0.04)));
        const double chg = fmin(max_c, capb) * t_u;
        b = round_dec(b + chg, 1e8);
        energy = dcl * 1e3 * 0.25 + chg * 1e3;
        co2 = energy * ci_now;
    } else if (bat_discharge) {                                  // discharge
        const double t_u = fmax(0.5, 4 * sigmoid(10 * (soc0 - 0.25))) * 15 / 60;
        const double max_d = fmin(fmin(capb, b / (0.01 + t_u)), dcl / 4);
        b = round_dec(b - fmin(max_d, capb) * t_u, 1e8);
        const double dis = (max_d < capb) ? max_d * t_u : capb * t_u;
        if (!(dcl * 1e3 * 0.25 >= dis * 1e3)) err |= SDC_F_BATTERY;
        energy = dcl * 1e3 * 0.25 - dis * 1e3;
        co2 = fmax(energy, 0.0) * ci_now;
    } else {                                                     // idle
        energy = dcl * 1e3 * 0.25;
        co2 = energy * ci_now;
    }
    S.bat_load[env] = b;
    const double soc = b * inv_capb;
    info(I_BAT_ACTION, a_bat_f); info(I_BAT_SOC, (float)soc); info(I_BAT_CO2, (float)co2);
    info(I_BAT_AVG_CI, (float)ci_now); info(I_BAT_E_WITHOUT, (float)(dcl * 1e3 * 0.25)); info(I_BAT_E_WITH, (float)energy);
    info(I_BAT_MAX_CAP, (float)capb); info(I_BAT_DCLOAD_MIN, (float)(P.power_lb_kw / 4));
    info(I_BAT_DCLOAD_MAX, (float)(P.power_ub_kw / 4));

    // ------------------------------------------------------------------------------------------
    // Managers advance (utils/managers.py:127-147,285-302,452-474,633-654), observations, info
    // ------------------------------------------------------------------------------------------
    const int tn = t + 1;
    const int step_in_ep = step0 + 1;
    S.t[env] = tn; S.step_in_ep[env] = step_in_ep;
    const int terminal = step_in_ep >= S.ep_len;
    od.ls = ls; od.soc = soc; od.nm = nm; od.tn = tn; od.loc = loc;
    const double inv_crng = 1.0 / nm.crng;
    const double nci_next = (ci_fut[0] - nm.cmin) * inv_crng;
    info(I_OUTSIDE_TEMP, (float)outside_next); info(I_DAY, (float)(tn / 96)); info(I_HOUR, (float)((tn % 96) * 0.25));
    info(I_NORM_CI, (float)nci_next);
#pragma unroll
    for (int i = 0; i < 8; ++i) info(I_FORECAST0 + i, (float)((ci_fut[i] - nm.cmin) * inv_crng));
    info(I_ISTERMINAL, terminal ? 1.f : 0.f);

    out.energy = energy; out.nci_next = nci_next;
    out.ls_penalty = (-0.3 * sqrt((double)over) + 0.3) + (-0.1 * ls.oldest);
    out.terminal = terminal; out.step_after = step_in_ep;
    out.co2 = co2; out.water = water; out.ite_kw = p_it * 1e-3; out.ct_kw = ct * 1e-3; out.comp_kw = comp * 1e-3;
    out.hvac_kw = (ct + comp) * 1e-3; out.total_kw = total_kw;
    out.tasks_in_queue = len; out.tasks_dropped = dropped; out.overdue = over;
    if (err) flag_error(S, env, err);
}

// Observations of the step (sustaindc_env.py:578-582), from what physics_step left in `od`.
template <class ObsSink>
SDC_HDN void emit_obs(const State& S, const Tables& T, int env, const ObsDeferred& od, ObsSink& obs) {
    build_obs(S, T, env, od.tn, od.ls, od.soc, od.nm, obs, od.loc);
}

// ---- reward normaliser (utils/reward_creator.py:16-45) without a window pass per step ----------
// The reference recomputes, every step, q1/q3 of the 10 000-sample energy window, clips the window to the IQR fences
// and takes mean / std of the clipped values.  One sample enters and at most one leaves per step, so all of that
// is maintained incrementally and EXACTLY (same multiset, same order statistics); the 40 KB window is streamed only
// when a maintained structure runs out of slack ("refresh", a few times per thousand steps per env):
//
//  * quartile brackets  lst[j][0..m) = order statistics at ranks a .. a+m-1 around rank k_j = floor((n-1) p_j)
//    (rank-contiguous, ties allowed).  Insert / evict update (a, m, lst) by comparing against the list ends only.  The
//    position of k_j inside the list performs a random walk; when fewer than kRecentreMargin ranks remain on a side
//    the next refresh re-centres the list: it collects ALL window values inside a value interval around the
//    quartile plus the count of values below the interval, sorts them and cuts ranks k-31 .. k+32 out of them.
//    Should the interval turn out too narrow / too wide (it adapts), the list is still extended by one exact rank
//    per step from the same scan (count and extreme of the values beyond the list end), so ranks k, k+1 never
//    leave the list whatever the distribution.
//  * moments  S1 = sum(x - c0), S2 = sum((x - c0)^2) over the whole window in fp64 (add the new sample, subtract
//    the evicted one; resynchronised exactly by every refresh).
//  * tails    clipping only changes the values beyond the fences, so sum(clip(x)) = S1 + sum_{x<lo}(lo - x) -
//    sum_{x>hi}(x - hi) (likewise for squares).  Around each fence a BAND of values (fence -/+ alpha IQR) is kept
//    individually as a SORTED list together with the split position of the fence inside it and the (sum, sum of
//    squares) of the band values beyond the fence; everything beyond the band as (count, sum, sum of squares).  A
//    step moves the split by the 0-2 values the fence crossed -- O(1) instead of a walk over the band -- and the rare
//    sample that enters / leaves a band (~2 % of the env-steps) is a sorted-list edit done by the whole warp.  When a
//    fence leaves its band, or a band overflows, the refresh rebuilds both (alpha adapts to the density at the fences;
//    distributions with heavy ties exactly at a fence fall back to a plain clipped-moment pass every step).
struct QView {              // the two lists of an env (rows of S.qlist)
    float* lst[2];
    int a[2], m[2];
};
enum ScanDir { SCAN_NONE = 0, SCAN_BELOW = 1, SCAN_ABOVE = 2 };
enum ScanKind { SCAN_SKIP = 0, SCAN_PLAIN = 1, SCAN_REFRESH = 2 };
struct ScanRequest {
    int n;                  // window length including the new value
    float lo, hi, shift;    // IQR fences and the centring shift of the scanned moment sums
    int dir[2];             // per quartile list: SCAN_BELOW -> count x < thr and track their max (rank a-1),
    float thr[2];           //                    SCAN_ABOVE -> count x > thr and track their min (rank a+m)
    int degenerate;         // q1 == q3: sigma is exactly 0 (utils/reward_creator.py:43-45 divides by 1)
    double q1;
    double lo64, hi64;      // the fences before the fp32 cast (the incremental path clips in fp64 like the reference)
    // refresh plan
    int kind;               // ScanKind
    int rc[2], k[2];        // re-centre list j around rank k[j] from all values in [ca[j], cb[j]]
    float ca[2], cb[2];
    int tails;              // rebuild the tail bands: tl2 <= x < tl and th < x <= th2 individually, beyond that aggregated
    float tl, th, tl2, th2;
    float e, o; int evict;  // the appended / evicted sample
};
struct ScanResult {
    float s1, s2;           // sum(c - shift), sum((c - shift)^2) over the clipped window
    int cnt[2];
    float ext[2];
    int recentred;          // bit j: the refresh rewrote list j (a, m in new_a / new_m)
    int new_a[2], new_m[2];
};
struct Moments { double c1, c2, c0; int ok; };   // clipped sums about c0 from the incremental path

#if defined(__CUDA_ARCH__)
#define SDC_INF_F __int_as_float(0x7f800000)
#else
#define SDC_INF_F INFINITY
#endif

// tail band `side` (0: below TL, 1: above TH) of an env: kTailCap floats, sorted ascending.
SDC_HD float* tail_ptr(const State& S, int env, int side) { return S.tails + ((size_t)env * 2 + side) * kTailCap; }

// Shifts inside a bracket run in batches of 8 (all loads of a batch before its stores): one lane moves up to a
// hundred floats through global memory, and element-by-element that is a chain of dependent round trips.
SDC_HD void shift_down(float* lst, int from, int to) {        // lst[i] = lst[i + 1] for i in [from, to)
    int i = from;
    for (; i + 8 <= to; i += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = lst[i + 1 + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) lst[i + u] = v[u];
    }
    for (; i < to; ++i) lst[i] = lst[i + 1];
}
SDC_HD void shift_up(float* lst, int from, int to) {          // lst[i] = lst[i - 1] for i in (from, to], descending
    int i = to;
    for (; i - 8 >= from; i -= 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = lst[i - 1 - u];
#pragma unroll
        for (int u = 0; u < 8; ++u) lst[i - u] = v[u];
    }
    for (; i > from; --i) lst[i] = lst[i - 1];
}
// first index in [0, m) with lst[i] == o (batched compares), or m
SDC_HD int find_equal(const float* lst, int m, float o) {
    for (int i0 = 0; i0 < m; i0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i0 + u < m) ? lst[i0 + u] : o;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (v[u] == o && i0 + u < m) return i0 + u;
    }
    return m;
}

// A structural edit of a bracket is PLANNED by the lane that owns the env and applied afterwards (on the device by the
// whole warp, coalesced; a lane moving a hundred floats by itself stalls its 31 neighbours):
//   B (memory) --remove index rm--> R --drop first / last--> R' --insert val at ins--> F (the edited list)
struct ListEdit { int rm, drop, ins; float val; int m0; };   // rm / ins -1 = none; drop 0 none, 1 first, 2 last; m0 = length of B
SDC_HD bool edit_trivial(const ListEdit& ed) { return ed.rm < 0 && ed.drop == 0 && ed.ins < 0; }
SDC_HD int edit_len(const ListEdit& ed) { return ed.m0 - (ed.rm >= 0) - (ed.drop != 0) + (ed.ins >= 0); }
SDC_HD float edit_at(const float* B, const ListEdit& ed, int j) {       // element j of F
    if (ed.ins >= 0) { if (j == ed.ins) return ed.val; if (j > ed.ins) j -= 1; }
    if (ed.drop == 1) j += 1;
    if (ed.rm >= 0 && j >= ed.rm) j += 1;
    return B[j];
}
SDC_HD float removed_at(const float* B, int rm, int j) { return B[(rm >= 0 && j >= rm) ? j + 1 : j]; }   // element j of R
// serial application (host build; the CUDA kernel applies edits warp-cooperatively)
SDC_HD void edit_apply(float* B, const ListEdit& ed) {
    if (edit_trivial(ed)) return;
    float tmp[kListCap];
    const int len = edit_len(ed);
    for (int j = 0; j < len; ++j) tmp[j] = edit_at(B, ed, j);
    for (int j = 0; j < len; ++j) B[j] = tmp[j];
}

// first / last: B[0], B[m-1] (cached by the caller)
// hint_rm: find_equal's answer when the caller already has it (the CUDA kernel searches warp-cooperatively), kNoHint otherwise
constexpr int kNoHint = -2;
SDC_HD void list_remove(const float* B, int& a, int& m, float o, float first, float last, ListEdit& ed, int& err, int hint_rm = kNoHint) {
    if (m == 0) { err |= SDC_F_BRACKET; return; }
    if (o < first) { a -= 1; return; }
    if (o > last) return;
    const int i = hint_rm != kNoHint ? hint_rm : find_equal(B, m, o);
    if (i == m) { err |= SDC_F_BRACKET; return; }
    ed.rm = i;
    m -= 1;
}

// m: length after the removal; k_after: rank that must stay inside the list (chooses the side to drop from when the list is full)
// hint_pos: the number of values <= e in R when the caller already has it, kNoHint otherwise
SDC_HD void list_insert(const float* B, int& a, int& m, float e, int n_after, int k_after, float first, float last, ListEdit& ed,
                        int hint_pos = kNoHint) {
    if (ed.rm >= 0 && m > 0) { first = removed_at(B, ed.rm, 0); last = removed_at(B, ed.rm, m - 1); }
    if (m > 0 && e < first && a > 0) { a += 1; return; }
    if (m > 0 && e > last && a + m < n_after - 1) return;           // ranks above the list, list not at the top
    // e belongs inside the list (or extends a list that reaches the end of the window): position after ties in R
    int lo = 0, hi = m;
    if (hint_pos != kNoHint) lo = hi = hint_pos;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (removed_at(B, ed.rm, mid) <= e) lo = mid + 1; else hi = mid; }
    const int pos = lo;
    if (m == kListCap) {
        // full: drop from the side with more spare ranks around the target k_after
        const int r = k_after - a;
        const int margin_lo = r, margin_hi = m - 1 - r;
        if (margin_lo > margin_hi) {            // drop the first value
            if (pos == 0) { a += 1; return; }   // e itself would be the dropped element
            ed.drop = 1; ed.ins = pos - 1; ed.val = e; a += 1; return;
        } else {                                // drop the last value
            if (pos == m) return;
            ed.drop = 2; ed.ins = pos; ed.val = e; return;
        }
    }
    ed.ins = pos; ed.val = e;
    m += 1;
}

// Appends `energy` to the env's reward window (fp32 ring; `o` = the value it evicts, if any), updates both
// quartile brackets and derives the fences.  utils/reward_creator.py:16-45.
// The window holds fp32 values measured from the env's first sample (hist_ref, fp64): the normaliser is translation
// invariant, and a fresh window -- whose spread is a tiny fraction of the energy itself, so that z = (E - mean) / std
// amplifies the storage rounding of absolute values -- is represented (nearly) exactly.  On return `energy` is the
// relative value, which is what reward_finish prices.
// Two halves so that the CUDA kernel can run the searches inside the brackets (the evicted value's index, the new value's
// position: a chain of dependent round trips for a single lane) with the whole warp in between:
//   reward_prepare_a  window append, the cached ends of both brackets, which lists need a search (PrepState::search)
//   reward_prepare_c  bracket updates (with the searches' answers as hints, or searching itself), quartiles, fences
struct PrepState { float first[2], last[2]; uint32_t fc; int err; int search[2]; };
struct ListHints { int rm[2], pos[2]; };

// bracket ends, for callers that want them in flight early (the CUDA kernel loads them before the physics)
SDC_HD void load_list_ends(const QView& Q, PrepState& ps) {
    for (int j = 0; j < 2; ++j) {
        const int m = Q.m[j];
        ps.first[j] = m > 0 ? Q.lst[j][0] : 0.f;
        ps.last[j] = m > 0 ? Q.lst[j][m - 1] : 0.f;
    }
}
SDC_HDN void reward_prepare_a(const State& S, int env, double& energy, int len, int head, float evicted, const QView& Q, ScanRequest& rq,
                              PrepState& ps, bool ends_loaded = false) {
    int err = 0;
    double ref = S.hist_ref[env];
    if (len == 0) { ref = (fabs(energy) <= 3.0e38) ? energy : 0.0; S.hist_ref[env] = ref; }
    energy -= ref;
    float e = (float)energy;
    if (!(fabs(energy) <= 3.0e38)) { err |= SDC_F_NONFINITE; e = 0.f; energy = 0.0; }
    const int cap = S.hist_cap;
    float* h = S.hist + (size_t)env * cap;
    float o = 0.f; bool evict = false;
    if (len == cap) { o = evicted; evict = true; } else { len += 1; }
    h[head] = e;
    head += 1; if (head == cap) head = 0;
    S.hist_len[env] = len; S.hist_head[env] = head;
    const int n = len;
    rq.n = n; rq.degenerate = 0; rq.e = e; rq.o = o; rq.evict = evict;
    ps.fc = S.fast_cfg[env];
    // The ends of both brackets are read together, before any store to the lists: almost every step only compares
    // against them (each a dependent L2 / DRAM round trip otherwise).
    if (!ends_loaded) load_list_ends(Q, ps);
    for (int j = 0; j < 2; ++j) {
        const int m = Q.m[j];
        // bit 0: the evicted value lies inside the list (its index is needed); bit 1: so does the new one (its position)
        ps.search[j] = m > 0 ? ((evict && o >= ps.first[j] && o <= ps.last[j]) ? 1 : 0) | ((e >= ps.first[j] && e <= ps.last[j]) ? 2 : 0) : 0;
    }
    ps.err = err;
}

SDC_HDN void reward_prepare_c(const State& S, int env, QView& Q, ScanRequest& rq, ListEdit* edits, const PrepState& ps, const ListHints& hints) {
    int err = ps.err;
    const int n = rq.n;
    const float e = rq.e, o = rq.o;
    const bool evict = rq.evict != 0;
    const uint32_t fc = ps.fc;
    const float* first = ps.first; const float* last = ps.last;
    double qv[2] = {0.0, 0.0};
    for (int j = 0; j < 2; ++j) {
        float* lst = Q.lst[j];
        int a = Q.a[j], m = Q.m[j];
        rq.dir[j] = SCAN_NONE; rq.thr[j] = 0.f; rq.rc[j] = 0; rq.ca[j] = rq.cb[j] = 0.f;
        const int num = (j == 0 ? 1 : 3) * (n - 1);
        const int k = num / 4;                                   // np.percentile 'linear': idx = (n-1)*p
        const double frac = (double)(num % 4) * 0.25;
        rq.k[j] = k;
        ListEdit& ed = edits[j];
        ed.rm = -1; ed.drop = 0; ed.ins = -1; ed.val = 0.f; ed.m0 = m;
        if (evict) list_remove(lst, a, m, o, first[j], last[j], ed, err, hints.rm[j]);
        if (m == 0 && ed.rm < 0) { ed.ins = 0; ed.val = e; a = 0; m = 1; }   // first value ever
        else list_insert(lst, a, m, e, n, k, first[j], last[j], ed, hints.pos[j]);
        const bool plain = edit_trivial(ed);                     // contents unchanged: the cached ends are valid
        if (n >= 2) {
            const int r = k - a;
            if (r < 0 || r + 1 >= m) {
                err |= SDC_F_BRACKET;
            } else {
                const double lo_v = edit_at(lst, ed, r), hi_v = edit_at(lst, ed, r + 1);
                const double d = hi_v - lo_v;                    // numpy _lerp
                qv[j] = (frac >= 0.5) ? hi_v - d * (1.0 - frac) : lo_v + d * frac;
                const int margin_lo = (a > 0) ? r : kListCap, margin_hi = (a + m < n) ? m - 2 - r : kListCap;
                const bool short_side = margin_lo < kRecentreMargin || margin_hi < kRecentreMargin;
                float f0 = first[j], f1 = last[j];
                if (short_side && !plain) { f0 = edit_at(lst, ed, 0); f1 = edit_at(lst, ed, m - 1); }
                if (margin_lo < kWidenMargin && margin_lo <= margin_hi) { rq.dir[j] = SCAN_BELOW; rq.thr[j] = f0; }
                else if (margin_hi < kWidenMargin) { rq.dir[j] = SCAN_ABOVE; rq.thr[j] = f1; }
                if (short_side) {
                    // all values within ~kCollectAim average list gaps of the quartile pair (width adapts per env and list)
                    const int wexp = (int)(int8_t)(fc >> (16 + 8 * j));
                    const float gap = (f1 - f0) / (float)(m - 1);
                    const float w = ldexpf(gap * (float)kCollectAim, wexp);
                    rq.rc[j] = 1; rq.ca[j] = (float)lo_v - w; rq.cb[j] = (float)hi_v + w;
                }
            }
        }
        Q.a[j] = a; Q.m[j] = m;
    }
    const double iqr = qv[1] - qv[0];
    rq.q1 = qv[0];
    rq.lo64 = qv[0] - 1.5 * iqr; rq.hi64 = qv[1] + 1.5 * iqr;
    rq.lo = (float)rq.lo64; rq.hi = (float)rq.hi64;
    rq.shift = (float)(0.5 * (qv[0] + qv[1]));
    rq.degenerate = (qv[0] == qv[1]);
    if (err) flag_error(S, env, err);
}

// serial statement (host build)
SDC_HDN void reward_prepare(const State& S, int env, double& energy, int len, int head, float evicted, QView& Q, ScanRequest& rq,
                            ListEdit* edits) {
    PrepState ps;
    reward_prepare_a(S, env, energy, len, head, evicted, Q, rq, ps);
    ListHints h; h.rm[0] = h.rm[1] = h.pos[0] = h.pos[1] = kNoHint;
    reward_prepare_c(S, env, Q, rq, edits, ps, h);
}

// Incremental side of the normaliser: updates the window moments, the tail bands and the far-tail aggregates with
// the step's sample, and returns the clipped sums when the bands still contain both fences.  Also decides what kind
// of window pass (if any) the step needs and, for a refresh, the new band thresholds.
//
// Per side the values beyond the inner threshold are split once more: those in the BAND between the inner and the
// outer threshold (TL2 <= x < TL, TH < x <= TH2; the fence lies inside the band) are kept individually, those
// beyond the outer threshold only as (count, sum, sum of squares) -- they are clipped whatever the fence does inside
// the band.  So a large outlier population (e.g. after a regime change) costs nothing per step.
//
// A step's work on the bands has three phases:
//   A (reward_plan_a, the env's lane)  window moments, far-tail aggregates, and whether the step's appended / evicted
//                                      sample is a band value (BandPlan)
//   B (band_edit; on the device the whole warp, k_step)  sorted removal / insertion of those samples
//   C (reward_plan_c, the env's lane)  split bookkeeping of the edits, the 0-2 values the fence crossed, clipped sums,
//                                      and what kind of window pass (if any) the step needs
struct BandPlan { int rm[2], ins[2]; };    // per side: the evicted sample leaves / the new sample enters the band
struct BandDone { int rm[2], pos[2]; };    // per side: index the evicted sample was removed at, index the new one went to (-1: none)

// Phase B, serial statement (host build; the CUDA kernel does the same with ballots over the 32 lanes).
// Removes the first element equal to `o` (if do_rm) and inserts `e` after its ties (if do_ins and the band has room).
SDC_HDN void band_edit(float* B, int n, bool do_rm, bool do_ins, float o, float e, int& rm, int& pos) {
    rm = -1; pos = -1;
    if (do_rm) for (int i = 0; i < n; ++i) if (B[i] == o) { rm = i; break; }
    const int n1 = n - (rm >= 0);
    if (do_ins && n1 < kTailCap) {
        pos = 0;
        for (int i = 0; i < n; ++i) pos += (i != rm && B[i] <= e);
    }
    ListEdit ed; ed.rm = rm; ed.drop = 0; ed.ins = pos; ed.val = e; ed.m0 = n;
    edit_apply(B, ed);
}

SDC_HDN void reward_plan_a(const State& S, int env, const ScanRequest& rq, Moments& M, BandPlan& bp) {
    const float e = rq.e, o = rq.o;
    const bool evict = rq.evict != 0;
    // all loads of the env's incremental state first (independent, in flight together), stores afterwards
    const double c0 = S.mom_c0[env];
    double s1 = S.mom_s1[env], s2 = S.mom_s2[env];
    const bool valid = S.tail_n[2 * env] >= 0;
    const float tl = S.tail_thr[4 * env], th = S.tail_thr[4 * env + 1], tl2 = S.tail_thr[4 * env + 2], th2 = S.tail_thr[4 * env + 3];
    int an[2] = {S.agg_n[2 * env], S.agg_n[2 * env + 1]};
    double a1[2] = {S.agg_s[4 * env], S.agg_s[4 * env + 2]}, a2[2] = {S.agg_s[4 * env + 1], S.agg_s[4 * env + 3]};
    const double ye = (double)e - c0, yo = (double)o - c0;
    s1 += ye; s2 = fma(ye, ye, s2);
    if (evict) { s1 -= yo; s2 = fma(-yo, yo, s2); }
    S.mom_s1[env] = s1; S.mom_s2[env] = s2;
    M.ok = valid; M.c0 = c0; M.c1 = s1; M.c2 = s2;       // whole-window sums; the clip corrections are added below and in phase C
    bp.rm[0] = bp.rm[1] = bp.ins[0] = bp.ins[1] = 0;
    if (!valid) return;
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
        const bool below = sd == 0;
        const float t_in = below ? tl : th, t_out = below ? tl2 : th2;
        // far tail: clipped to the fence as a whole
        const bool e_far = below ? e < t_out : e > t_out, o_far = evict && (below ? o < t_out : o > t_out);
        if (e_far) { an[sd] += 1; a1[sd] += ye; a2[sd] = fma(ye, ye, a2[sd]); }
        if (o_far) { an[sd] -= 1; a1[sd] -= yo; a2[sd] = fma(-yo, yo, a2[sd]); }
        const double yf = (below ? rq.lo64 : rq.hi64) - c0;
        M.c1 += (double)an[sd] * yf - a1[sd];
        M.c2 += (double)an[sd] * yf * yf - a2[sd];
        bp.rm[sd] = evict && !o_far && (below ? o < t_in : o > t_in);
        bp.ins[sd] = !e_far && (below ? e < t_in : e > t_in);
    }
    S.agg_n[2 * env] = an[0]; S.agg_n[2 * env + 1] = an[1];
    S.agg_s[4 * env] = a1[0]; S.agg_s[4 * env + 1] = a2[0]; S.agg_s[4 * env + 2] = a1[1]; S.agg_s[4 * env + 3] = a2[1];
}

SDC_HDN void reward_plan_c(const State& S, int env, ScanRequest& rq, Moments& M, const BandPlan& bp, const BandDone& bd) {
    const float e = rq.e, o = rq.o;
    uint32_t fc = S.fast_cfg[env];
    const float tl = S.tail_thr[4 * env], th = S.tail_thr[4 * env + 1], tl2 = S.tail_thr[4 * env + 2], th2 = S.tail_thr[4 * env + 3];
    const bool had_bands = M.ok != 0;
    if (M.ok) {
        const double c0 = M.c0;
        bool valid = true;
        int err = 0;
        int n[2] = {S.tail_n[2 * env], S.tail_n[2 * env + 1]}, nb[2] = {S.tail_nb[2 * env], S.tail_nb[2 * env + 1]};
        double b1[2] = {S.tail_bs[4 * env], S.tail_bs[4 * env + 2]}, b2[2] = {S.tail_bs[4 * env + 1], S.tail_bs[4 * env + 3]};
        float near_lo[2][2];           // per side: the band values just below / at the split, loaded together (see below)
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            const bool below = sd == 0;
            // bookkeeping of the edits phase B applied: the lower band's beyond-set is its first nb values, the upper band's its last nb
            if (bp.rm[sd]) {
                if (bd.rm[sd] < 0) { err |= SDC_F_BRACKET; valid = false; }
                else {
                    if (below ? bd.rm[sd] < nb[sd] : bd.rm[sd] >= n[sd] - nb[sd]) { const double y = (double)o - c0; nb[sd] -= 1; b1[sd] -= y; b2[sd] = fma(-y, y, b2[sd]); }
                    n[sd] -= 1;
                }
            }
            if (bp.ins[sd]) {
                if (bd.pos[sd] < 0) valid = false;                               // band full: rebuilt (narrower) by the next refresh
                else {
                    if (below ? bd.pos[sd] < nb[sd] : bd.pos[sd] > n[sd] - nb[sd]) { const double y = (double)e - c0; nb[sd] += 1; b1[sd] += y; b2[sd] = fma(y, y, b2[sd]); }
                    n[sd] += 1;
                }
            }
        }
        // The values the fence crossed since the last step (usually none or one).  The two band values on either side of each
        // split decide the common case; all four are loaded before any of them is used (four dependent round trips otherwise).
        if (valid) {
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                const float* p = tail_ptr(S, env, sd);
                const int s = sd == 0 ? nb[sd] : n[sd] - nb[sd];                 // first index past (lower) / of (upper) the beyond-set
                near_lo[sd][0] = s > 0 ? p[s - 1] : 0.f;
                near_lo[sd][1] = s < n[sd] ? p[s] : 0.f;
            }
        }
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            const bool below = sd == 0;
            const float* p = tail_ptr(S, env, sd);
            const double fence = below ? rq.lo64 : rq.hi64;
            if (valid) {
                if (below) {
                    int k = nb[sd];
                    if (k > 0 && !((double)near_lo[sd][0] < fence)) {
                        k -= 1; { const double y = (double)near_lo[sd][0] - c0; b1[sd] -= y; b2[sd] = fma(-y, y, b2[sd]); }
                        while (k > 0 && !((double)p[k - 1] < fence)) { k -= 1; const double y = (double)p[k] - c0; b1[sd] -= y; b2[sd] = fma(-y, y, b2[sd]); }
                    } else if (k < n[sd] && (double)near_lo[sd][1] < fence) {
                        { const double y = (double)near_lo[sd][1] - c0; b1[sd] += y; b2[sd] = fma(y, y, b2[sd]); } k += 1;
                        while (k < n[sd] && (double)p[k] < fence) { const double y = (double)p[k] - c0; b1[sd] += y; b2[sd] = fma(y, y, b2[sd]); k += 1; }
                    }
                    nb[sd] = k;
                } else {
                    int sx = n[sd] - nb[sd];                                     // first index of the beyond-set
                    if (sx < n[sd] && !((double)near_lo[sd][1] > fence)) {
                        { const double y = (double)near_lo[sd][1] - c0; b1[sd] -= y; b2[sd] = fma(-y, y, b2[sd]); } sx += 1;
                        while (sx < n[sd] && !((double)p[sx] > fence)) { const double y = (double)p[sx] - c0; b1[sd] -= y; b2[sd] = fma(-y, y, b2[sd]); sx += 1; }
                    } else if (sx > 0 && (double)near_lo[sd][0] > fence) {
                        sx -= 1; { const double y = (double)near_lo[sd][0] - c0; b1[sd] += y; b2[sd] = fma(y, y, b2[sd]); }
                        while (sx > 0 && (double)p[sx - 1] > fence) { sx -= 1; const double y = (double)p[sx] - c0; b1[sd] += y; b2[sd] = fma(y, y, b2[sd]); }
                    }
                    nb[sd] = n[sd] - sx;
                }
                if (nb[sd] == 0) { b1[sd] = 0.0; b2[sd] = 0.0; }                 // no rounding residue survives an empty set
                const double yf = fence - c0;
                M.c1 += (double)nb[sd] * yf - b1[sd];
                M.c2 += (double)nb[sd] * yf * yf - b2[sd];
            }
        }
        if (err) flag_error(S, env, err);
        if (!valid) { n[0] = -1; n[1] = -1; }
        S.tail_n[2 * env] = n[0]; S.tail_n[2 * env + 1] = n[1];
        S.tail_nb[2 * env] = nb[0]; S.tail_nb[2 * env + 1] = nb[1];
        S.tail_bs[4 * env] = b1[0]; S.tail_bs[4 * env + 1] = b2[0]; S.tail_bs[4 * env + 2] = b1[1]; S.tail_bs[4 * env + 3] = b2[1];
        const double lo = rq.lo64, hi = rq.hi64;
        M.ok = valid && (double)tl2 <= lo && lo <= (double)tl && (double)th <= hi && hi <= (double)th2;
    }
    // ---- what does this step need from the window? ----
    const bool moments_needed = rq.n >= 2 && !rq.degenerate;
    const bool lists = rq.rc[0] || rq.rc[1] || rq.dir[0] || rq.dir[1];
    int retry = (int)((fc >> 8) & 0xffu);
    rq.kind = SCAN_SKIP; rq.tails = 0; rq.tl = rq.tl2 = -SDC_INF_F; rq.th = rq.th2 = SDC_INF_F;
    bool want_tails = false;
    if (moments_needed && !M.ok) {
        if (retry > 0) { retry -= 1; fc = (fc & ~0xff00u) | ((uint32_t)retry << 8); S.fast_cfg[env] = fc; }
        else want_tails = true;
    }
    // A fence that has drifted into the outer twelfth of its band gets the band re-centred NOW, by a maintenance pass,
    // while the step can still be priced incrementally; once the fence is outside, the pass is on the step's critical path.
    bool near_edge = false;
    if (moments_needed && M.ok) {
        const double lo = rq.lo64, hi = rq.hi64;
        const double ql = 0.08 * ((double)tl - (double)tl2), qh = 0.08 * ((double)th2 - (double)th);
        near_edge = (lo - (double)tl2) < ql || ((double)tl - lo) < ql || (hi - (double)th) < qh || ((double)th2 - hi) < qh;
    }
    if (lists || want_tails || near_edge) {
        rq.kind = SCAN_REFRESH;
        if (moments_needed && retry == 0) {
            const int aexp = (int)(fc & 0xffu);
            const double iqr4 = rq.hi64 - rq.lo64;              // = 4 (q3 - q1)
            const double slack = ldexp(0.125 * iqr4, -aexp);    // half-width of the bands: alpha = 0.5 / 2^aexp of the IQR
            float tl = (float)(rq.lo64 + slack), th = (float)(rq.hi64 - slack);
            float tl2 = (float)(rq.lo64 - slack), th2 = (float)(rq.hi64 + slack);
            if ((double)tl < rq.lo64) tl = nextafterf(tl, SDC_INF_F);
            if ((double)th > rq.hi64) th = nextafterf(th, -SDC_INF_F);
            if ((double)tl2 > rq.lo64) tl2 = nextafterf(tl2, -SDC_INF_F);
            if ((double)th2 < rq.hi64) th2 = nextafterf(th2, SDC_INF_F);
            rq.tails = (want_tails && had_bands) ? 2 : 1;       // 2: a fence left its (still valid) band
            rq.tl = tl; rq.th = th; rq.tl2 = tl2; rq.th2 = th2;
        }
    } else if (moments_needed && !M.ok) {
        rq.kind = SCAN_PLAIN;
    }
}

// Raw output of a refresh pass over the window (the CUDA warp scan / the serial host scan fill this).
struct RefreshRaw {
    double s1, s2;          // exact unclipped sums about the new centre (ScanRequest::shift)
    int n_tail[2];          // values inside the lower / upper band (may exceed kTailCap: then only the count is valid)
    int band_nb[2];         // of those, the values beyond the fence of the requesting step, with their
    double band_b1[2], band_b2[2];   // sum / sum of squares about the new centre
    int agg_n[2]; double agg_s1[2], agg_s2[2];   // values beyond the outer thresholds: count and sums about the new centre
    int c[2], below[2];     // per list: values inside [ca, cb] (may exceed kCollectCap), values below ca
};

// Turns the collected values into the env's new incremental state.  `sorted[j]`: the c[j] collected values of list j in
// ascending order (valid when c[j] <= kCollectCap); `bands[sd]`: the n_tail[sd] band values in ascending order (valid when
// both fit).  Lane-strided so that a warp can share the copying; all lanes get the same return values.  On the host
// lane = 0, n_lanes = 1.
SDC_HDN void refresh_commit(const State& S, int env, const ScanRequest& rq, const RefreshRaw& raw, const float* const* sorted,
                            const float* const* bands, QView& Q, ScanResult& rs, int lane, int n_lanes) {
    uint32_t fc = S.fast_cfg[env];
    int aexp = (int)(fc & 0xffu), retry = (int)((fc >> 8) & 0xffu);
    int wexp[2] = {(int)(int8_t)(fc >> 16), (int)(int8_t)(fc >> 24)};
    rs.recentred = 0;
    for (int j = 0; j < 2; ++j) {
        rs.new_a[j] = Q.a[j]; rs.new_m[j] = Q.m[j];
        if (!rq.rc[j]) continue;
        const int n = rq.n, k = rq.k[j], c = raw.c[j], below = raw.below[j];
        bool ok = false;
        if (c > kCollectCap) {
            if (wexp[j] > -kWexpMax) wexp[j] -= 1;                        // too wide for the scratch: halve next time
        } else {
            // The side that ran short tells which way the rank is drifting (a window in transition moves it steadily one
            // way): give that side two thirds of the new list.
            const int r_old = k - Q.a[j];
            const bool short_hi = (Q.m[j] - 2 - r_old) < r_old;
            int a_new = k - (short_hi ? kListCap / 3 : (2 * kListCap) / 3);
            if (a_new < below) a_new = below;
            int end = a_new + kListCap;
            if (end > below + c) end = below + c;
            if (end - a_new < kListCap) { a_new = end - kListCap; if (a_new < below) a_new = below; }
            const int m_new = end - a_new, r = k - a_new;
            if (r >= 0 && r + 1 < m_new) {
                const int mlo = (a_new > 0) ? r : kListCap, mhi = (a_new + m_new < n) ? m_new - 2 - r : kListCap;
                ok = mlo >= kRecentreOk && mhi >= kRecentreOk;
            }
            if (ok) {
                for (int i = lane; i < m_new; i += n_lanes) Q.lst[j][i] = sorted[j][a_new - below + i];
                rs.new_a[j] = a_new; rs.new_m[j] = m_new; rs.recentred |= 1 << j;
            } else if (wexp[j] < kWexpMax) {
                wexp[j] += 1;                                             // too narrow: double next time
            }
        }
    }
    if (rq.tails) {
        const bool fits = raw.n_tail[0] <= kTailCap && raw.n_tail[1] <= kTailCap;
        if (fits) {
            const int big = raw.n_tail[0] > raw.n_tail[1] ? raw.n_tail[0] : raw.n_tail[1];
            for (int sd = 0; sd < 2; ++sd) {
                float* p = tail_ptr(S, env, sd);
                for (int i = lane; i < raw.n_tail[sd]; i += n_lanes) p[i] = bands[sd][i];
            }
            if (lane == 0) {
                S.tail_n[2 * env] = raw.n_tail[0]; S.tail_n[2 * env + 1] = raw.n_tail[1];
                S.tail_nb[2 * env] = raw.band_nb[0]; S.tail_nb[2 * env + 1] = raw.band_nb[1];
                S.tail_bs[4 * env] = raw.band_b1[0]; S.tail_bs[4 * env + 1] = raw.band_b2[0];
                S.tail_bs[4 * env + 2] = raw.band_b1[1]; S.tail_bs[4 * env + 3] = raw.band_b2[1];
                S.tail_thr[4 * env] = rq.tl; S.tail_thr[4 * env + 1] = rq.th; S.tail_thr[4 * env + 2] = rq.tl2; S.tail_thr[4 * env + 3] = rq.th2;
                S.agg_n[2 * env] = raw.agg_n[0]; S.agg_n[2 * env + 1] = raw.agg_n[1];
                S.agg_s[4 * env] = raw.agg_s1[0]; S.agg_s[4 * env + 1] = raw.agg_s2[0];
                S.agg_s[4 * env + 2] = raw.agg_s1[1]; S.agg_s[4 * env + 3] = raw.agg_s2[1];
                S.mom_s1[env] = raw.s1; S.mom_s2[env] = raw.s2; S.mom_c0[env] = (double)rq.shift;
            }
            // band width policy: widen after a refresh that was forced by a fence leaving its band (rq.tails == 2) or when both
            // bands came out sparse, narrow when one is close to its capacity
            if (rq.tails == 2 || big < kBandSparse) { if (aexp > 0) aexp -= 1; }
            else for (int b = big; b > kBandTarget && aexp < kAlphaOff; b >>= 1) aexp += 1;
        } else {
            if (lane == 0) { S.tail_n[2 * env] = -1; S.tail_n[2 * env + 1] = -1; }
            if (aexp < kAlphaOff) aexp += 1;                             // dense around a fence: narrower bands
            else retry = kTailRetry;                                     // hopeless (heavy ties at a fence): plain scans for a while
        }
    }
    if (lane == 0)
        S.fast_cfg[env] = (uint32_t)(aexp & 0xff) | ((uint32_t)(retry & 0xff) << 8) | ((uint32_t)(uint8_t)(int8_t)wexp[0] << 16) |
                          ((uint32_t)(uint8_t)(int8_t)wexp[1] << 24);
}

// Applies the scan's results to the brackets (re-centred lists are taken over, the others extended by the one exact
// rank the scan found) and turns the clipped moments into the three rewards.
SDC_HDN void reward_finish(const State& S, int env, const ScanRequest& rq, const ScanResult& rs, const Moments& M,
                           const RewardInputs& st, const float* alt3, QView& Q, float* rew3) {
    int err = 0;
    const int n = rq.n;
    if (rq.kind == SCAN_REFRESH) {
        for (int j = 0; j < 2; ++j) {
            if (rs.recentred & (1 << j)) { Q.a[j] = rs.new_a[j]; Q.m[j] = rs.new_m[j]; continue; }
            float* lst = Q.lst[j];
            int a = Q.a[j], m = Q.m[j];
            const int c = rs.cnt[j];
            if (rq.dir[j] == SCAN_BELOW) {                           // rank a-1
                if (c > a) err |= SDC_F_BRACKET;
                const float v = (a > c) ? lst[0] : rs.ext[j];        // a tie copy of lst[0] sits below the list
                if (m == kListCap) m -= 1;                           // drop the top (far side)
                shift_up(lst, 0, m);
                lst[0] = v; a -= 1; m += 1;
            } else if (rq.dir[j] == SCAN_ABOVE) {                    // rank a+m
                const int above = n - a - m;
                if (c > above) err |= SDC_F_BRACKET;
                const float v = (above > c) ? lst[m - 1] : rs.ext[j];
                if (m == kListCap) { shift_down(lst, 0, m - 1); m -= 1; a += 1; }
                lst[m] = v; m += 1;
            }
            Q.a[j] = a; Q.m[j] = m;
        }
    }
    double z = 0.0;
    if (n >= 2) {
        if (rq.degenerate) {
            z = st.energy - rq.q1;                                   // every clipped value equals q1: std == 0 -> / 1
        } else if (M.ok) {
            const double md = M.c1 / n;
            const double var = M.c2 / n - md * md;
            const double sd = var > 0.0 ? sqrt(var) : 0.0;
            z = (st.energy - (M.c0 + md)) / (sd > 0.0 ? sd : 1.0);
        } else {
            const double md = (double)rs.s1 / n;
            const double var = (double)rs.s2 / n - md * md;
            const double sd = var > 0.0 ? sqrt(var) : 0.0;
            z = (st.energy - ((double)rq.shift + md)) / (sd > 0.0 ? sd : 1.0);
        }
    }
    const double foot = -1.0 * (st.nci_next * z / 0.50);         // reward_creator.py:67-72
    double r_ls = foot + st.ls_penalty;
    r_ls = fmin(fmax(r_ls, -10.0), 10.0);                        // :82
    rew3[0] = (float)r_ls; rew3[1] = (float)foot; rew3[2] = (float)foot;
    if (any_alt_reward(S)) {                                     // uniform across the batch
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int kind = S.reward_kind[a];
            if (kind > SDC_R_DEFAULT_DC) rew3[a] = alt3[a];
            else if (kind == SDC_R_DEFAULT_DC) rew3[a] = (float)foot;
        }
    }
    if (err) flag_error(S, env, err);
}

// ---- counter-based RNG for on-device episode starts and weather noise ---------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Replaces the reference's global `random` / legacy
// `np.random` draws at reset (sustaindc_env.py:454-455, utils/managers.py:35-48,596-603) in
// generated mode; replay mode injects the reference's realised values instead (sdc_stage_episode).
struct U4 { uint32_t x, y, z, w; };
SDC_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
SDC_HD U4 philox4x32(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        U4 n;
        n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}
enum RngStream { RS_START = 0, RS_NOISE = 1 };
SDC_HD U4 env_random(uint64_t seed, uint32_t episode, uint32_t stream, uint32_t idx) {
    U4 c; c.x = idx; c.y = episode; c.z = stream; c.w = 0x5DCB200u;
    return philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}
// Weather noise: the year-long random walk is cut into kNoiseSegs segments of kNoiseSeg samples; segment i draws its
// normals from its own PCG32 stream (O'Neill 2014, XSH-RR 64/32), seeded by one Philox block keyed by (env seed, episode,
// segment).  ~10 integer instructions per 32-bit draw instead of ~23 for Philox: the walk is 35 040 normals per episode,
// a fifth of all instructions of a step when it ran on Philox alone.
struct Pcg32 { uint64_t state, inc; };
SDC_HD uint32_t pcg32_next(Pcg32& g) {
    const uint64_t old = g.state;
    g.state = old * 6364136223846793005ULL + g.inc;
    const uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27), rot = (uint32_t)(old >> 59);
    return (xs >> rot) | (xs << ((32u - rot) & 31u));
}
SDC_HD Pcg32 noise_stream(uint64_t seed, uint32_t episode, uint32_t segment) {
    const U4 r = env_random(seed, episode, RS_NOISE, segment);
    Pcg32 g;
    g.state = (uint64_t)r.x | ((uint64_t)r.y << 32);
    g.inc = ((uint64_t)r.z | ((uint64_t)r.w << 32)) | 1ull;
    return g;
}
// Two N(0,1) samples from ONE 32-bit draw: radius from its top 16 bits (u1 = (i + 0.5) / 65536, |z| <= 4.9), angle in
// [-pi, pi) from its low 16 bits read as a signed integer (Box-Muller, fp32).  On the device the logarithm, square root and
// sine / cosine are the hardware approximations (MUFU; absolute error ~2^-21 in the range used): ~12 instructions per pair
// instead of ~150, and a difference from the host statement below the 1e-5 level in z (tests compare the realised weather at
// 1e-4 C).  One draw per pair halves the generator work per walk sample; the year-long walk (35 040 increments of 0.02 z,
// rescaled to a standard deviation of 0.75 C) does not resolve 16-bit uniforms from 24- / 32-bit ones.
SDC_HD void noise_pair(Pcg32& g, float* z2) {
    const uint32_t a = pcg32_next(g);
    const float u1 = ((float)(a >> 16) + 0.5f) * (1.0f / 65536.0f);
    const float th = (float)(int)(int16_t)(uint16_t)(a & 0xffffu) * 9.587379924285257e-5f;     // pi / 2^15
#if defined(__CUDA_ARCH__)
    float r, s, c;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
    __sincosf(th, &s, &c);
#else
    const float r = sqrtf(-2.0f * logf(u1)), s = sinf(th), c = cosf(th);
#endif
    z2[0] = r * c; z2[1] = r * s;
}
// Episode start (day, hour) and weather day-roll: random.randint(lo, hi), random.randint(0, 23),
// np.random.randint(0, 14).
SDC_HD void draw_episode_start(uint64_t seed, uint32_t episode, int day_lo, int day_hi, int* day, int* hour, int* roll) {
    const U4 r = env_random(seed, episode, RS_START, 0);
    *day = day_lo + (int)(r.x % (uint32_t)(day_hi - day_lo + 1));
    *hour = (int)(r.y % 24u);
    *roll = (int)(r.z % 14u);
}
constexpr int kNoiseThreads = 256;                        // threads of an episode generation, one walk segment each
constexpr int kNoiseSegs = kNoiseThreads;                 // segments of the year-long random walk (one PCG32 stream each)
constexpr int kNoiseSeg = 140;                            // even segment length, 256 * 140 >= 35040

// Buffers and knobs of one step launch (all device pointers; see sdc_step in include/sdc_b200.h).
struct StepArgs {
    const int32_t* actions; float* obs; float* share; float* rew; uint8_t* done; float* info; float* term_obs;
    // compact outputs (sdc_step_compact): the env's 29 distinct observation values, [N][SDC_OBS_COMPACT]; obs / share /
    // term_obs may then be null
    float* obs_c; float* term_c;
    // this step's counters (ctr) and the next step's (ctr_next, zeroed by this launch):
    //   [0] unit tickets  [1] finished envs appended to reset_list  [2] units past the scalar phase
    //   [3] reset_list slots claimed by workers  [4..7] statistics: plain passes, refresh passes, by brackets, by tails
    //   [8] envs appended to pre_list  [9] pre_list_prev entries claimed by workers
    //   [10] maintenance passes published  [11] claimed by workers
    // three counter blocks rotate: the previous step's block (ctr_prev) still holds its pre_list count
    int32_t* ctr;
    int32_t* ctr_next;
    const int32_t* ctr_prev;
    int32_t* reset_list;   // [N + slack], -1 = empty slot
    void* pass_jobs;       // [N] records of maintenance passes (window refreshes whose result this step does not need)
    int32_t* pass_ready;   // [N] == seq once record i is complete
    int32_t seq;           // step tag (never 0)
    int32_t* pre_list;     // [N] envs that finish two steps from now (filled by this launch)
    const int32_t* pre_list_prev;   // the list the previous launch filled: episodes to pre-generate now
    double* metrics;
    unsigned long long* hvac_hist;      // [SDC_HVAC_BINS] counts of positive HVAC power samples
    float hvac_bins_per_kw;             // SDC_HVAC_BINS / range
    unsigned long long* phase_clocks;   // optional [16]: summed per-warp clock64 deltas of the k_step phases (diagnostics)
    uint32_t* unit_log;                 // optional [units][8]: per-unit phase clocks, SM id, start time (diagnostics, with phase_clocks)
    unsigned long long* pass_total;     // [4] running totals of ctr[4..7] (window passes: plain, refresh, by brackets, by tails)
    int32_t unit_envs, blocks_per_sm;
};

// compact column c of [agent_ls 26 | workload(t+1) | norm T(t+1) | SoC] -> column of the zero-padded [3][26] row
// (agent_dc[11], agent_dc[13], agent_bat[12]; every other dc / bat entry repeats an agent_ls entry, sustaindc_env.py:302-433)
SDC_HD int compact_to_padded(int c) { return c < 26 ? c : (c == 26 ? SDC_OBS_DIM + 11 : (c == 27 ? SDC_OBS_DIM + 13 : 2 * SDC_OBS_DIM + 12)); }

// HARL shared observation (harl/envs/sustaindc/harlsustaindc_env.py:78-85): ls[0:26] | dc[11] | dc[13] |
// last element of the zero-padded battery row (always 0.0).
SDC_HD void share_from_obs(const float* obs78, float* share29) {
    for (int i = 0; i < SDC_OBS_DIM; ++i) share29[i] = obs78[i];
    share29[26] = obs78[SDC_OBS_DIM + 11];
    share29[27] = obs78[SDC_OBS_DIM + 13];
    share29[28] = obs78[2 * SDC_OBS_DIM + SDC_OBS_DIM - 1];
}

// ---- reset of the scalar sub-env state (sustaindc_env.py:436-531) ----------------------------
// The weather window / norms must already be in place. Set-point and reward window survive.
// `Lk`: the env's location tables when the caller has them at hand (the CUDA kernel: its shared-memory copy), else looked up
SDC_HD void reset_scalars(const State& S, int env, int t0, const LocTables* Lk = nullptr) {
    const LocTables& L = Lk ? *Lk : S.loc[S.loc_id[env]];
    const double cmin = L.ci_min30[t0], cmax = L.ci_max30[t0];         // both loads before the first store
    S.t[env] = t0; S.t0[env] = t0; S.step_in_ep[env] = 0;
    S.ci_min[env] = cmin; S.ci_max[env] = cmax;
    S.ls_head[env] = t0; S.ls_len[env] = 0; S.ls_sum[env] = 0;
    S.ls_bins[env * 4 + 0] = 0; S.ls_bins[env * 4 + 1] = 0; S.ls_bins[env * 4 + 2] = 0; S.ls_bins[env * 4 + 3] = 0;
    S.dc_run[env] = 0; S.dc_scale[env] = 1; S.dc_last[env] = 2;              // dc_gym.py:114-116
    S.bat_load[env] = 0.0;                                                    // battery_model.py:90-91
    if (t0 + S.ep_len + 18 > SDC_YEAR_STEPS) flag_error(S, env, SDC_F_TRACE_DOMAIN);
}
// The observation SustainDC.reset returns (:488-494, 531): empty queue, SoC 0, trace index t0, the given window / range.
template <class ObsSink>
SDC_HDN void reset_observation(const State& S, int env, int t0, const double* window, double tmin, double tmax, ObsSink& obs) {
    const LocTables& L = S.loc[S.loc_id[env]];
    LsStats ls;
    ls.oldest = 0.0; ls.avg = 0.0; ls.norm_q = 0.0;
    for (int i = 0; i < 5; ++i) ls.hist[i] = 0.0;
    const Tables T{S.loc, S.dc};
    Norms nm;
    nm.cmin = L.ci_min30[t0]; nm.crng = L.ci_max30[t0] - nm.cmin;
    nm.tmin = tmin; nm.trng = tmax - tmin; nm.wrel = window - t0;
    build_obs(S, T, env, t0, ls, 0.0, nm, obs);
}

}  // namespace sdc
