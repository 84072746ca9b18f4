// TEST INFRASTRUCTURE ONLY -- serial host build of the device logic.
//
// Compiles dc_rl_b200/csrc/sdc_core.h (the scalar per-env functions the CUDA kernels call) and
// dc_rl_b200/csrc/sdc_api.inc (the C-ABI host layer) with g++ against a trivial in-memory backend, so
// that the step logic, the rolling-quartile brackets and the reset/auto-reset sequencing can be
// unit-tested against the oracle on machines without a GPU.  It is built by tests/ into
// tests/hostsim/_build/libsdc_hostsim.so and is never loaded by the dc_rl_b200 package: the product
// path is libsdc_b200.so (CUDA) only and fails loudly without it.
//
// The "kernels" below are plain loops; the full-window scan is a scalar loop (the CUDA kernel does
// the same sums warp-parallel, so low-order bits of the fp32 moments differ).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../dc_rl_b200/csrc/sdc_core.h"

namespace backend {
struct Context { int unused = 0; };
using StepArgs = sdc::StepArgs;
static const char* init(Context&, int) { return nullptr; }
static void shutdown(Context&) {}
static const char* dev_alloc(Context&, void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? nullptr : "malloc"; }
static void dev_free(Context&, void* p) { free(p); }
static const char* dev_zero(Context&, void* p, size_t bytes) { memset(p, 0, bytes); return nullptr; }
static const char* h2d(Context&, void* d, const void* s, size_t n) { memcpy(d, s, n); return nullptr; }
static const char* d2h(Context&, void* d, const void* s, size_t n) { memcpy(d, s, n); return nullptr; }
static const char* h2d_async(Context&, void* d, const void* s, size_t n, void*) { memcpy(d, s, n); return nullptr; }
static const char* d2h_async(Context&, void* d, const void* s, size_t n, void*) { memcpy(d, s, n); return nullptr; }
static const char* pinned_alloc(Context&, void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? nullptr : "malloc"; }
static void pinned_free(Context&, void* p) { free(p); }
static bool host_memory_is_device_visible(Context&) { return true; }
static const char* stream_create(Context&, void** s) { *s = nullptr; return nullptr; }
static const char* stream_sync(Context&, void*) { return nullptr; }
static const char* sync(Context&) { return nullptr; }
static const char* dev_fill_bytes(Context&, void* p, int v, size_t bytes) { memset(p, v, bytes); return nullptr; }
static const char* event_create(Context&, void** ev) { *ev = nullptr; return nullptr; }
static const char* event_record(Context&, void*, void*) { return nullptr; }
static const char* event_elapsed_ms(Context&, void*, void*, double* ms) { *ms = 0.0; return nullptr; }
static void event_destroy(Context&, void*) {}
static void range_push(const char*) {}
static void range_pop() {}
static int max_window_len() { return SDC_YEAR_STEPS; }

struct ObsRow {
    float* row;   // [3][26]
    void operator()(int agent, int idx, float v) { row[agent * SDC_OBS_DIM + idx] = v; }
};
struct InfoCol {
    float* info; int n, env;
    void operator()(int col, float v) { if (info) info[(size_t)col * n + env] = v; }
};

// plain pass: clipped moments
static void scan_plain(const sdc::State& S, int env, const sdc::ScanRequest& rq, sdc::ScanResult& rs) {
    const float* h = S.hist + (size_t)env * S.hist_cap;
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < rq.n; ++i) {
        const float d = fminf(fmaxf(h[i], rq.lo), rq.hi) - rq.shift;
        s1 += d; s2 += d * d;
    }
    rs.s1 = s1; rs.s2 = s2;
}

// refresh pass: serial statement of scan_refresh in sdc_kernels.cu
static void scan_refresh(const sdc::State& S, int env, const sdc::ScanRequest& rq, sdc::RefreshRaw& raw, sdc::ScanResult& rs,
                         std::vector<float> coll[2], std::vector<float> band[2]) {
    const float* h = S.hist + (size_t)env * S.hist_cap;
    float s1 = 0.f, s2 = 0.f;
    double S1 = 0.0, S2 = 0.0;
    const double c0 = (double)rq.shift;
    band[0].clear(); band[1].clear();
    raw.n_tail[0] = raw.n_tail[1] = 0;
    for (int j = 0; j < 2; ++j) { raw.agg_n[j] = 0; raw.agg_s1[j] = raw.agg_s2[j] = 0.0; }
    for (int j = 0; j < 2; ++j) {
        rs.cnt[j] = 0; rs.ext[j] = rq.dir[j] == sdc::SCAN_ABOVE ? INFINITY : -INFINITY;
        raw.c[j] = 0; raw.below[j] = 0; coll[j].clear();
    }
    for (int i = 0; i < rq.n; ++i) {
        const float x = h[i];
        const float d = fminf(fmaxf(x, rq.lo), rq.hi) - rq.shift;
        s1 += d; s2 += d * d;
        const double y = (double)x - c0;
        S1 += y; S2 += y * y;
        for (int j = 0; j < 2; ++j) {
            if (rq.dir[j] == sdc::SCAN_BELOW && x < rq.thr[j]) { rs.cnt[j]++; rs.ext[j] = fmaxf(rs.ext[j], x); }
            if (rq.dir[j] == sdc::SCAN_ABOVE && x > rq.thr[j]) { rs.cnt[j]++; rs.ext[j] = fminf(rs.ext[j], x); }
            if (rq.rc[j]) {
                if (x < rq.ca[j]) raw.below[j]++;
                else if (x <= rq.cb[j]) { raw.c[j]++; if ((int)coll[j].size() < sdc::kCollectCap) coll[j].push_back(x); }
            }
        }
        if (x < rq.tl2) { raw.agg_n[0]++; raw.agg_s1[0] += y; raw.agg_s2[0] += y * y; }
        else if (x < rq.tl) { if (raw.n_tail[0] < sdc::kTailCap) band[0].push_back(x); raw.n_tail[0]++; }
        if (x > rq.th2) { raw.agg_n[1]++; raw.agg_s1[1] += y; raw.agg_s2[1] += y * y; }
        else if (x > rq.th) { if (raw.n_tail[1] < sdc::kTailCap) band[1].push_back(x); raw.n_tail[1]++; }
    }
    for (int j = 0; j < 2; ++j) {
        if (rq.dir[j] == sdc::SCAN_NONE) rs.ext[j] = 0.f;
        std::sort(coll[j].begin(), coll[j].end());
    }
    rs.s1 = s1; rs.s2 = s2; raw.s1 = S1; raw.s2 = S2;
    for (int sd = 0; sd < 2; ++sd) {                 // sorted bands + the part of each beyond the requesting step's fence
        std::sort(band[sd].begin(), band[sd].end());
        raw.band_nb[sd] = 0; raw.band_b1[sd] = raw.band_b2[sd] = 0.0;
        for (float x : band[sd]) {
            if (sd == 0 ? (double)x < rq.lo64 : (double)x > rq.hi64) { const double y = (double)x - c0; raw.band_nb[sd]++; raw.band_b1[sd] += y; raw.band_b2[sd] += y * y; }
        }
    }
}

static void reset_envs(const sdc::State& S, const int32_t* list, int count, float* obs, float* share, float* obs_c);
static void store_rows(const float* row, int env, float* obs, float* share, float* obs_c) {
    if (obs) memcpy(obs + (size_t)env * 3 * SDC_OBS_DIM, row, 3 * SDC_OBS_DIM * sizeof(float));
    if (share) sdc::share_from_obs(row, share + (size_t)env * SDC_SHARE_DIM);
    if (obs_c) for (int k = 0; k < SDC_OBS_COMPACT; ++k) obs_c[(size_t)env * SDC_OBS_COMPACT + k] = row[sdc::compact_to_padded(k)];
}

static const char* launch_step(Context& cx, const sdc::State& S, const StepArgs& a, void*) {
    for (int k = 0; k < 16; ++k) a.ctr_next[k] = 0;
    const int N = S.n_envs;
    for (int env = 0; env < N; ++env) {
        float row78[3 * SDC_OBS_DIM];
        ObsRow obs{row78};
        InfoCol info{a.info, N, env};
        sdc::StepResult st;
        const sdc::Tables T{S.loc, S.dc};
        sdc::ObsDeferred od;
        sdc::physics_step(S, T, env, a.actions[env * 3 + 0], a.actions[env * 3 + 1], a.actions[env * 3 + 2], info, st, od);
        sdc::emit_obs(S, T, env, od, obs);
        store_rows(row78, env, a.obs, a.share, a.obs_c);
        sdc::ScanRequest rq; sdc::ScanResult rs; sdc::Moments mo;
        rs.s1 = rs.s2 = 0.f; rs.cnt[0] = rs.cnt[1] = 0; rs.ext[0] = rs.ext[1] = 0.f; rs.recentred = 0;
        sdc::QView Q;
        for (int j = 0; j < 2; ++j) {
            Q.lst[j] = S.qlist + ((size_t)env * 2 + j) * sdc::kListCap; Q.a[j] = S.q_a[env * 2 + j]; Q.m[j] = S.q_m[env * 2 + j];
        }
        sdc::ListEdit edits[2];
        rq.kind = sdc::SCAN_SKIP; rq.n = 0; rq.degenerate = 0; rq.dir[0] = rq.dir[1] = 0; mo.ok = 0;
        double e_rel = st.energy;                                                  // becomes relative to the env's hist_ref
        if (S.append_history) {
            sdc::reward_prepare(S, env, e_rel, st.hist_len, st.hist_head, st.evicted, Q, rq, edits);
            for (int j = 0; j < 2; ++j) sdc::edit_apply(Q.lst[j], edits[j]);      // the CUDA kernel does this warp-cooperatively
            sdc::BandPlan bp;
            sdc::BandDone bd; bd.rm[0] = bd.rm[1] = bd.pos[0] = bd.pos[1] = -1;
            sdc::reward_plan_a(S, env, rq, mo, bp);
            for (int sd = 0; sd < 2; ++sd)                                         // the CUDA kernel does this warp-cooperatively too
                if (bp.rm[sd] || bp.ins[sd])
                    sdc::band_edit(sdc::tail_ptr(S, env, sd), S.tail_n[2 * env + sd], bp.rm[sd], bp.ins[sd], rq.o, rq.e, bd.rm[sd], bd.pos[sd]);
            sdc::reward_plan_c(S, env, rq, mo, bp, bd);
        }
        if (rq.kind == sdc::SCAN_PLAIN) {
            scan_plain(S, env, rq, rs);
            a.ctr[4] += 1; a.pass_total[0] += 1;
        } else if (rq.kind == sdc::SCAN_REFRESH) {
            sdc::RefreshRaw raw;
            std::vector<float> coll[2], band[2];
            scan_refresh(S, env, rq, raw, rs, coll, band);
            const float* sorted[2] = {coll[0].data(), coll[1].data()};
            const float* bands[2] = {band[0].data(), band[1].data()};
            sdc::refresh_commit(S, env, rq, raw, sorted, bands, Q, rs, 0, 1);
            a.ctr[5] += 1; a.pass_total[1] += 1;
        }
        sdc::RewardInputs en{e_rel, st.nci_next, st.ls_penalty};
        float alt3[3] = {0.f, 0.f, 0.f};
        if (sdc::any_alt_reward(S)) {
            const sdc::AltInputs ai{st.ite_kw, st.total_kw, st.water, od.tn % 96};
            sdc::alt_rewards(S, env, st.energy, ai, alt3);
        }
        sdc::reward_finish(S, env, rq, rs, mo, en, alt3, Q, a.rew + (size_t)env * 3);
        for (int j = 0; j < 2; ++j) { S.q_a[env * 2 + j] = Q.a[j]; S.q_m[env * 2 + j] = Q.m[j]; }
        a.done[env] = (uint8_t)st.terminal;
        if (st.terminal) {
            store_rows(row78, env, a.term_obs, nullptr, a.term_c);
            a.reset_list[a.ctr[1]++] = env;
        }
        double* M = a.metrics;
        M[sdc::M_ENERGY] += st.energy; M[sdc::M_CO2] += st.co2; M[sdc::M_WATER] += st.water;
        M[sdc::M_TASKS_IN_QUEUE] += st.tasks_in_queue; M[sdc::M_TASKS_DROPPED] += st.tasks_dropped;
        M[sdc::M_ITE_KW] += st.ite_kw; M[sdc::M_CT_KW] += st.ct_kw; M[sdc::M_COMP_KW] += st.comp_kw; M[sdc::M_HVAC_KW] += st.hvac_kw;
        M[sdc::M_STEPS] += 1; M[sdc::M_EPISODES] += st.terminal;
        if (st.hvac_kw > 0.0) {
            int bin = (int)(st.hvac_kw * (double)a.hvac_bins_per_kw);
            bin = bin < 0 ? 0 : (bin >= SDC_HVAC_BINS ? SDC_HVAC_BINS - 1 : bin);
            a.hvac_hist[bin] += 1;
        }
        const float* r = a.rew + (size_t)env * 3;
        M[sdc::M_REWARD_SUM] += (double)r[0] + r[1] + r[2]; M[sdc::M_REWARD_LS] += r[0]; M[sdc::M_REWARD_DC] += r[1];
        M[sdc::M_OVERDUE] += st.overdue; M[sdc::M_TOTAL_KW] += st.total_kw;
    }
    (void)cx;
    reset_envs(S, a.reset_list, a.ctr[1], a.obs, a.share, a.obs_c);             // the CUDA kernel does this inside the launch
    for (int i = 0; i < a.ctr[1]; ++i) a.reset_list[i] = -1;
    return nullptr;
}

static const char* launch_build_reset_list(Context&, const sdc::State& S, const uint8_t* mask, int32_t* list, int32_t* count, void*) {
    int c = 0;
    for (int env = 0; env < S.n_envs; ++env) if (!mask || mask[env]) list[c++] = env;
    *count = c;
    return nullptr;
}

// Weather of one episode from the device RNG -- serial statement of what k_reset does block-parallel
// (utils/managers.py:35-48,594-613).
static void generate_weather(const sdc::State& S, int env, int t0, int roll, uint32_t episode) {
    const sdc::LocTables& L = S.loc[S.loc_id[env]];
    const int n = SDC_YEAR_STEPS;
    const uint64_t seed = S.seed[env];
    std::vector<float> inc(sdc::kNoiseSegs * sdc::kNoiseSeg, 0.f);
    for (int seg = 0; seg < sdc::kNoiseSegs; ++seg) {    // one PCG32 stream per segment, two normals per draw
        sdc::Pcg32 g = sdc::noise_stream(seed, episode, (uint32_t)seg);
        for (int q = 0; q < sdc::kNoiseSeg; q += 2) {
            float z[2];
            sdc::noise_pair(g, z);
            inc[seg * sdc::kNoiseSeg + q] = 0.02f * z[0]; inc[seg * sdc::kNoiseSeg + q + 1] = 0.02f * z[1];
        }
    }
    std::vector<double> walk(n);
    double acc = 0.0;
    for (int i = 0; i < sdc::kNoiseSegs; ++i) {             // same segment structure as the kernel
        double seg = 0.0;
        for (int j = i * sdc::kNoiseSeg; j < (i + 1) * sdc::kNoiseSeg && j < n; ++j) { seg += (double)inc[j]; walk[j] = acc + seg; }
        acc += seg;
    }
    double sum = 0.0;
    for (int j = 0; j < n; ++j) sum += walk[j];
    const double mean = sum / n;
    double ss = 0.0;
    for (int j = 0; j < n; ++j) ss += (walk[j] - mean) * (walk[j] - mean);
    const double scale = 0.75 / std::sqrt(ss / n);
    double* wt = sdc::weather_pend(S, env);                  // generated into the staging buffer, flipped in by the reset
    double* ww = wt + S.win_len;
    for (int i = 0; i < S.win_len; ++i) { wt[i] = 0.0; ww[i] = 0.0; }
    double tmin = INFINITY, tmax = -INFINITY;
    for (int j = 0; j < n; ++j) {
        const int t = (j + 96 * roll) % n;
        if (t < t0) continue;
        const double noise = walk[j] * scale;
        const double vt = std::fmin(std::fmax(L.temp_base[j] + noise, 0.0), 45.0);
        if (t < t0 + 2880) { tmin = std::fmin(tmin, vt); tmax = std::fmax(tmax, vt); }
        if (t < t0 + S.win_len) {
            wt[t - t0] = vt;
            ww[t - t0] = std::fmin(std::fmax(L.wetb_base[j] + noise, 0.0), 45.0);
        }
    }
    S.pend_tmin[env] = tmin; S.pend_tmax[env] = tmax;
}

static const char* launch_reset(Context&, const sdc::State& S, const int32_t* list, const int32_t* count, float* obs, float* share, void*) {
    reset_envs(S, list, *count, obs, share, nullptr);
    return nullptr;
}

static void reset_envs(const sdc::State& S, const int32_t* list, int count, float* obs, float* share, float* obs_c) {
    for (int i = 0; i < count; ++i) {
        const int env = list[i];
        int day, hour, roll = 0;
        if (S.pend_valid[env] & 1) {
            day = S.pend_day[env]; hour = S.pend_hour[env];
        } else {
            const uint32_t ep = S.episode[env];
            sdc::draw_episode_start(S.seed[env], ep, S.day_lo[env], S.day_hi[env], &day, &hour, &roll);
            generate_weather(S, env, day * 96 + hour * 4, roll, ep);
        }
        const double* window = sdc::weather_pend(S, env);
        S.t_min[env] = S.pend_tmin[env]; S.t_max[env] = S.pend_tmax[env];
        S.cur_buf[env] ^= 1;
        S.pend_valid[env] = 0;
        S.episode[env] += 1;
        memset(S.ls_ring + (size_t)env * (S.ls_mask + 1), 0, S.ls_mask + 1);
        float row78[3 * SDC_OBS_DIM];
        ObsRow o{row78};
        const int t0 = day * 96 + hour * 4;
        sdc::reset_scalars(S, env, t0);
        sdc::reset_observation(S, env, t0, window, S.t_min[env], S.t_max[env], o);
        store_rows(row78, env, obs, share, obs_c);
    }
}

static const char* launch_rebuild(Context&, const sdc::State& S, void*) {
    for (int env = 0; env < S.n_envs; ++env) {
        const int n = S.hist_len[env];
        std::vector<float> v(S.hist + (size_t)env * S.hist_cap, S.hist + (size_t)env * S.hist_cap + n);
        std::sort(v.begin(), v.end());
        for (int j = 0; j < 2; ++j) {
            float* lst = S.qlist + ((size_t)env * 2 + j) * sdc::kListCap;
            int a = 0, m = 0;
            if (n > 0) {
                const int k = ((j == 0 ? 1 : 3) * (n - 1)) / 4;
                a = k - (sdc::kListCap / 2 - 1); if (a + sdc::kListCap > n) a = n - sdc::kListCap; if (a < 0) a = 0;
                m = n - a < sdc::kListCap ? n - a : sdc::kListCap;
                for (int i = 0; i < m; ++i) lst[i] = v[a + i];
            }
            S.q_a[env * 2 + j] = a; S.q_m[env * 2 + j] = m;
            S.tail_n[env * 2 + j] = -1;
        }
    }
    return nullptr;
}
}  // namespace backend

#include "../../dc_rl_b200/csrc/sdc_api.inc"
