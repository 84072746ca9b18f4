// sdc_kernels.cu -- CUDA (sm_100a) backend of libsdc_b200.so.
//
// Kernels
//   k_step    one launch per env-step for all N envs.  A warp takes a unit of U consecutive envs:
//             phase A  one lane per env: load-shifting queue, IT/HVAC model, battery, trace gathers,
//                      observations, info row, append of the step energy to the reward window and O(1)
//                      update of the rolling quartile brackets (sdc_core.h, fp64 like the reference);
//             phase B  the whole warp streams each env's fp32 reward window (40 KB at steady state, the
//                      dominant HBM traffic) once with 128-bit loads and warp-shuffle reductions: clipped
//                      moments, plus the next rank of a bracket side that runs short;
//             phase C  one lane per env: z-score -> three rewards, bracket extension, metrics.
//             Observation rows are staged in shared memory and written as one contiguous tile per unit.
//   k_reset   one CTA per finished env: start day/hour, year-long weather random walk (Philox), day roll,
//             clip, 30-day normalisation, queue clear, reset observation (or copies a staged episode).
//   k_rebuild one CTA per env: full bitonic sort of the window in shared memory -> fresh brackets.
//   k_build_reset_list  mask -> env list.
//
// No tensor cores: there is no dense contraction on this path (HBM-bound streaming + scalar physics).
#include <cuda_runtime.h>

#include "sdc_core.h"

namespace backend {

struct Context {
    int device = 0;
    int sm_count = 148;
    int step_blocks_per_sm = 2;
};
using StepArgs = sdc::StepArgs;

#define CU(expr)                                              \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return cudaGetErrorString(_e); \
    } while (0)

static const char* set_kernel_attributes();
static const char* init(Context& c, int device) {
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return "device ordinal out of range";
    c.device = device;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return "libsdc_b200 requires an sm_100a (B200) device";
    c.sm_count = prop.multiProcessorCount;
    return set_kernel_attributes();
}
static void shutdown(Context&) {}
static const char* dev_alloc(Context& c, void** p, size_t bytes) { CU(cudaSetDevice(c.device)); CU(cudaMalloc(p, bytes ? bytes : 16)); return nullptr; }
static void dev_free(Context&, void* p) { cudaFree(p); }
static const char* dev_zero(Context&, void* p, size_t bytes) { CU(cudaMemset(p, 0, bytes)); return nullptr; }
static const char* h2d(Context&, void* d, const void* s, size_t n) { CU(cudaMemcpy(d, s, n, cudaMemcpyHostToDevice)); return nullptr; }
static const char* d2h(Context&, void* d, const void* s, size_t n) { CU(cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost)); return nullptr; }
static const char* h2d_async(Context&, void* d, const void* s, size_t n, void* st) {
    CU(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, (cudaStream_t)st)); return nullptr;
}
static const char* d2h_async(Context&, void* d, const void* s, size_t n, void* st) {
    CU(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, (cudaStream_t)st)); return nullptr;
}
static const char* pinned_alloc(Context&, void** p, size_t bytes) { CU(cudaMallocHost(p, bytes ? bytes : 16)); return nullptr; }
static void pinned_free(Context&, void* p) { cudaFreeHost(p); }
static const char* stream_create(Context&, void** s) { cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); *s = st; return nullptr; }
static const char* stream_sync(Context&, void* s) { CU(cudaStreamSynchronize((cudaStream_t)s)); return nullptr; }
static const char* event_create(Context&, void** ev) { cudaEvent_t e; CU(cudaEventCreate(&e)); *ev = e; return nullptr; }
static const char* event_record(Context&, void* ev, void* st) { CU(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)st)); return nullptr; }
static const char* event_elapsed_ms(Context&, void* a, void* b, double* ms) {
    float f = 0.f; CU(cudaEventElapsedTime(&f, (cudaEvent_t)a, (cudaEvent_t)b)); *ms = f; return nullptr;
}
static void event_destroy(Context&, void* ev) { cudaEventDestroy((cudaEvent_t)ev); }
static size_t reset_scratch_floats(Context&) { return 4; }      // the reset workers need no global scratch any more
static const char* dev_fill_bytes(Context&, void* p, int v, size_t bytes) { CU(cudaMemset(p, v, bytes)); return nullptr; }
static const char* sync(Context& c) { CU(cudaSetDevice(c.device)); CU(cudaDeviceSynchronize()); return nullptr; }

// =================================================================================================
// device helpers
// =================================================================================================
constexpr int kStepThreads = 256;
constexpr int kWarpsPerBlock = kStepThreads / 32;
constexpr int kObsRow = 3 * SDC_OBS_DIM;          // 78 floats per env
constexpr int kListRow = 2 * sdc::kListCap + 4;   // both quartile lists of an env; 16-byte aligned rows for cp.async
constexpr int kTableBytes = 8192;                 // shared-memory copy of the location / dc parameter tables

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// TMA-engine prefetch of a contiguous global range into L2 (one instruction, no registers, no smem).
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void prefetch_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Issues, up front and all at once, the second-level (address-dependent) reads of one env-step so that their
// DRAM latencies overlap instead of being paid one after another inside the scalar phase.
__device__ __forceinline__ void prefetch_env(const sdc::State& S, const sdc::Tables& T, int env) {
    const int t = S.t[env], t0 = S.t0[env], head = S.ls_head[env], hh = S.hist_head[env];
    const sdc::LocTables& L = T.loc[S.loc_id[env]];
    const double* wt = S.weather + (size_t)env * 2 * S.win_len + (t - t0);
    prefetch_line(wt); prefetch_line(wt + 16); prefetch_line(wt + S.win_len);
    const uint8_t* ring = S.ls_ring + (size_t)env * (S.ls_mask + 1);
    prefetch_line(ring + ((t - 24) & S.ls_mask)); prefetch_line(ring + ((t - 48) & S.ls_mask));
    prefetch_line(ring + ((t - 72) & S.ls_mask)); prefetch_line(ring + ((t - 96) & S.ls_mask));
    prefetch_line(ring + (t & S.ls_mask)); prefetch_line(ring + (head & S.ls_mask));
    prefetch_line(S.hist + (size_t)env * S.hist_cap + hh);
    prefetch_line(L.ci + t - 16); prefetch_line(L.ci + t); prefetch_line(L.ci + t + 9);
    prefetch_line(L.workload + t); prefetch_line(L.ns + t); prefetch_line(L.sh + t);
}

// Observation rows written straight to the output tensors: obs[env][3][26], the HARL shared observation
// (ls[0:26] | dc[11] | dc[13] | padded battery row [25], harlsustaindc_env.py:78-85) and, for finished envs, term_obs.
struct GlobalObsSink {
    float* obs; float* share; float* term;
    __device__ __forceinline__ void operator()(int agent, int idx, float v) {
#ifdef SDC_EXPERIMENT_NO_OBS_STORES
        if (v == 123.456f) obs[0] = v;
        return;
#endif
        obs[agent * SDC_OBS_DIM + idx] = v;
        if (term) term[agent * SDC_OBS_DIM + idx] = v;
        if (agent == 0) share[idx] = v;
        else if (agent == 1 && idx == 11) share[26] = v;
        else if (agent == 1 && idx == 13) share[27] = v;
        else if (agent == 2 && idx == 25) share[28] = v;
    }
};
struct GlobalInfoSink {
    float* info; int n, env;
    __device__ __forceinline__ void operator()(int col, float v) { if (info) info[(size_t)col * n + env] = v; }
};

// Streaming 128-bit load of the reward window.  Deliberately a WEAK load (no `volatile`, evict-first hint): the
// CUDA intrinsics (__ldcg/__ldcs) are `asm volatile` and ld.global.cg compiles to LDG...STRONG.GPU, which ptxas
// keeps in order and interleaves with the arithmetic -- only ~3 loads in flight when the warp first stalls.  Weak
// loads let it issue the whole batch up front.  Coherence: a window line is read once per launch, after its new
// sample was stored (ordering via the tagged queue words), and L1 is invalidated at kernel boundaries.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ---- phase B: one warp streams one env's window -------------------------------------------------
template <int D>
__device__ __forceinline__ void track(int& cnt, float& ext, float x, float thr) {
    if (D == sdc::SCAN_BELOW) { const bool b = x < thr; cnt += b; ext = fmaxf(ext, b ? x : -SDC_INF_F); }
    if (D == sdc::SCAN_ABOVE) { const bool b = x > thr; cnt += b; ext = fminf(ext, b ? x : SDC_INF_F); }
}

template <int D0, int D1>
struct Acc {
    float s1[4], s2[4];
    int cnt[2];
    float ext[2];
    float lo, hi, shift, thr0, thr1;
    __device__ __forceinline__ void one(int k, float x) {
        const float c = fminf(fmaxf(x, lo), hi);
        const float d = c - shift;
        s1[k] += d;
        s2[k] = fmaf(d, d, s2[k]);
        track<D0>(cnt[0], ext[0], x, thr0);
        track<D1>(cnt[1], ext[1], x, thr1);
    }
    __device__ __forceinline__ void four(const float4& v) { one(0, v.x); one(1, v.y); one(2, v.z); one(3, v.w); }
};

template <int D0, int D1, int UNROLL>
__device__ __forceinline__ void scan_window(const float* h, int n, float lo, float hi, float shift, float thr0, float thr1,
                                            int lane, sdc::ScanResult& rs) {
    Acc<D0, D1> A;
#pragma unroll
    for (int k = 0; k < 4; ++k) { A.s1[k] = 0.f; A.s2[k] = 0.f; }
    A.cnt[0] = A.cnt[1] = 0;
    A.ext[0] = D0 == sdc::SCAN_ABOVE ? SDC_INF_F : -SDC_INF_F;
    A.ext[1] = D1 == sdc::SCAN_ABOVE ? SDC_INF_F : -SDC_INF_F;
    A.lo = lo; A.hi = hi; A.shift = shift; A.thr0 = thr0; A.thr1 = thr1;
    const float4* p = reinterpret_cast<const float4*>(h);
    const int n4 = n >> 2;
    // Full batches of UNROLL rows (UNROLL x 512 B per warp in flight).  Deeper per-warp queues (double buffering,
    // UNROLL 16, bulk L2 prefetch) were measured SLOWER on B200: with ~2 400 independent 40 KB streams more
    // outstanding requests only add DRAM row conflicts (profiles/r01_summary.md).
    int c = lane;
    const int nb = n4 / (UNROLL * 32);               // warp-uniform trip count (the fence below is a warp barrier)
    for (int b = 0; b < nb; ++b, c += UNROLL * 32) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ld_stream(p + c + u * 32);
        __syncwarp();            // scheduling fence: ptxas must issue the whole batch before the first use (see scan_job)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) A.four(v[u]);
    }
    for (; c + 3 * 32 < n4; c += 4 * 32) {           // tail: groups of four rows, then single rows
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_stream(p + c + u * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) A.four(v[u]);
    }
    for (; c < n4; c += 32) A.four(ld_stream(p + c));
    const int rem = n & 3;
    if (lane < rem) A.one(0, ld_stream(h + (n4 << 2) + lane));
    rs.s1 = warp_sum((A.s1[0] + A.s1[1]) + (A.s1[2] + A.s1[3]));
    rs.s2 = warp_sum((A.s2[0] + A.s2[1]) + (A.s2[2] + A.s2[3]));
    rs.cnt[0] = D0 ? warp_sum(A.cnt[0]) : 0;
    rs.cnt[1] = D1 ? warp_sum(A.cnt[1]) : 0;
    rs.ext[0] = D0 == sdc::SCAN_BELOW ? warp_max(A.ext[0]) : (D0 == sdc::SCAN_ABOVE ? warp_min(A.ext[0]) : 0.f);
    rs.ext[1] = D1 == sdc::SCAN_BELOW ? warp_max(A.ext[1]) : (D1 == sdc::SCAN_ABOVE ? warp_min(A.ext[1]) : 0.f);
}

template <int UNROLL>
// Inlined on purpose: results and parameters stay in registers.  Anything that goes through local memory in the
// per-job path (a by-reference result struct, a State copy made for a non-inlined callee) misses the small L1 most of
// the time and then costs a ~1.5 us round trip while HBM is saturated -- measured: +50 % per window scan.
__device__ __forceinline__ void scan_dispatch(const float* h, int n, float lo, float hi, float shift, int d0, int d1, float t0, float t1,
                                           int lane, sdc::ScanResult& rs) {
    switch (d0 * 3 + d1) {
        case 0: scan_window<0, 0, UNROLL>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
        case 1: scan_window<0, 1, 8>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
        case 2: scan_window<0, 2, 8>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
        case 3: scan_window<1, 0, 8>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
        case 4: scan_window<1, 1, 8>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
        case 5: scan_window<1, 2, 8>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
        case 6: scan_window<2, 0, 8>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
        case 7: scan_window<2, 1, 8>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
        default: scan_window<2, 2, 8>(h, n, lo, hi, shift, t0, t1, lane, rs); break;
    }
}

// =================================================================================================
// episode reset of one env by one CTA (used by k_reset and by the reset workers inside k_step)
// =================================================================================================
constexpr int kResetThreads = sdc::kNoiseThreads;   // 256 == kStepThreads

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kResetThreads / 32; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ double block_minmax(double v, bool is_min, double* red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_min ? fmin(v, other) : fmax(v, other);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int w = 1; w < kResetThreads / 32; ++w) t = is_min ? fmin(t, red[w]) : fmax(t, red[w]);
    return t;
}

struct RowSink {
    float* row;
    __device__ __forceinline__ void operator()(int agent, int idx, float v) { row[agent * SDC_OBS_DIM + idx] = v; }
};

struct ResetShared {
    double red[kResetThreads / 32];
    double seg_off[kResetThreads];
    int start[3];
    float row[kObsRow];
};

constexpr int kNormWindow = 2880;                 // 30 days of quarter-hours (utils/managers.py:435,606)

// Episode reset of one env by one CTA.  `runbuf` = kNormWindow doubles of shared memory.
// Weather noise (utils/managers.py:35-48,596-613): the year-long random walk is generated ONCE (Philox, 140 samples per
// thread); its mean / variance come from per-thread partial sums combined with the segment offsets, and only the walk
// values that land in the 30-day window after the start are kept (in shared memory).  The emit pass then runs over
// window positions, so its trace reads are coalesced and independent.  No global scratch: a dependent global access
// costs ~2 us while the other CTAs saturate HBM with window scans.
__device__ __forceinline__ void reset_one_env(const sdc::State& S, int env, float* obs, float* share, double* runbuf, ResetShared& sh) {
    const int tid = threadIdx.x;
    const int n = SDC_YEAR_STEPS;
    double* wt = S.weather + (size_t)env * 2 * S.win_len;
    double* ww = wt + S.win_len;
    const bool staged = S.pend_valid && S.pend_valid[env];
    __syncthreads();
    if (staged) {
        if (tid == 0) { sh.start[0] = S.pend_day[env]; sh.start[1] = S.pend_hour[env]; sh.start[2] = 0; }
        const double* src = S.pend_weather + (size_t)env * 2 * S.win_len;
        for (int k = tid; k < 2 * S.win_len; k += kResetThreads) wt[k] = src[k];
        if (tid == 0) { S.t_min[env] = S.pend_tmin[env]; S.t_max[env] = S.pend_tmax[env]; }
    } else {
        const uint32_t ep = S.episode[env];
        const uint64_t seed = S.seed[env];
        if (tid == 0) sdc::draw_episode_start(seed, ep, S.day_lo[env], S.day_hi[env], &sh.start[0], &sh.start[1], &sh.start[2]);
        __syncthreads();
        const int t0 = sh.start[0] * 96 + sh.start[1] * 4, roll = sh.start[2];
        const int k_max = min(kNormWindow, n - t0);            // the reference's slice is truncated at the year end
        // pass 1: this thread's segment of the walk (utils/managers.py:45-46)
        double run = 0.0, sum_run = 0.0, sum_run2 = 0.0;
        int cnt = 0;
        for (int q = 0; q < sdc::kNoiseSeg / 4; ++q) {
            float z[4];
            const int j0 = tid * sdc::kNoiseSeg + q * 4;
            sdc::noise_normals4(seed, ep, (uint32_t)(j0 >> 2), z);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j < n) {
                    run += (double)(0.02f * z[u]);
                    sum_run += run; sum_run2 += run * run; cnt += 1;
                    int t = j + 96 * roll; if (t >= n) t -= n;
                    const int k = t - t0;
                    if (k >= 0 && k < k_max) runbuf[k] = run;
                }
            }
        }
        sh.seg_off[tid] = run;
        __syncthreads();
        if (tid == 0) {                           // serial exclusive prefix, same order as the host statement
            double acc = 0.0;
            for (int k = 0; k < kResetThreads; ++k) { const double s = sh.seg_off[k]; sh.seg_off[k] = acc; acc += s; }
        }
        __syncthreads();
        const double off = sh.seg_off[tid];
        // walk_j = off + run_j  ->  sums of w and w^2 from the partial sums
        const double sw = block_sum(cnt * off + sum_run, sh.red);
        const double sw2 = block_sum(cnt * off * off + 2.0 * off * sum_run + sum_run2, sh.red);
        const double mean = sw / n;
        const double scale = 0.75 / sqrt(sw2 / n - mean * mean);              // managers.py:46-48
        // emit: roll, clip, window, 30-day min/max (managers.py:598-608), one window position per thread and trip
        const sdc::LocTables& L = S.loc[S.loc_id[env]];
        double tmin = INFINITY, tmax = -INFINITY;
        for (int k = tid; k < max(k_max, S.win_len); k += kResetThreads) {
            if (k < k_max) {
                int j = t0 + k - 96 * roll; if (j < 0) j += n;
                const double noise = (sh.seg_off[j / sdc::kNoiseSeg] + runbuf[k]) * scale;
                const double vt = fmin(fmax(L.temp_base[j] + noise, 0.0), 45.0);
                tmin = fmin(tmin, vt); tmax = fmax(tmax, vt);
                if (k < S.win_len) { wt[k] = vt; ww[k] = fmin(fmax(L.wetb_base[j] + noise, 0.0), 45.0); }
            } else if (k < S.win_len) {
                wt[k] = 0.0; ww[k] = 0.0;                                      // beyond the year end (flagged domain)
            }
        }
        tmin = block_minmax(tmin, true, sh.red);
        tmax = block_minmax(tmax, false, sh.red);
        if (tid == 0) { S.t_min[env] = tmin; S.t_max[env] = tmax; }
    }
    uint8_t* ring = S.ls_ring + (size_t)env * (S.ls_mask + 1);
    for (int k = tid; k <= S.ls_mask; k += kResetThreads) ring[k] = 0;
    __syncthreads();                               // weather window + norms visible to thread 0
    if (tid == 0) {
        if (staged) S.pend_valid[env] = 0;
        S.episode[env] += 1;
        RowSink sink{sh.row};
        sdc::reset_scalar_state(S, env, sh.start[0] * 96 + sh.start[1] * 4, sink);
    }
    __syncthreads();
    for (int k = tid; k < kObsRow; k += kResetThreads) obs[(size_t)env * kObsRow + k] = sh.row[k];
    if (tid < SDC_SHARE_DIM) {
        const int k = tid;
        const int src = k < 26 ? k : (k == 26 ? SDC_OBS_DIM + 11 : (k == 27 ? SDC_OBS_DIM + 13 : 2 * SDC_OBS_DIM + 25));
        share[(size_t)env * SDC_SHARE_DIM + k] = sh.row[src];
    }
}

// =================================================================================================
// k_reset: explicit resets (sdc_reset), one CTA per listed env
// =================================================================================================
__global__ void __launch_bounds__(kResetThreads) k_reset(const sdc::State S, const int32_t* __restrict__ list,
                                                         const int32_t* __restrict__ count, float* obs, float* share) {
    extern __shared__ double runbuf[];              // [kNormWindow] walk values inside the 30-day window
    __shared__ ResetShared sh;
    const int total = *count;
    for (int i = blockIdx.x; i < total; i += gridDim.x) reset_one_env(S, list[i], obs, share, runbuf, sh);
}

// =================================================================================================
// k_step
// =================================================================================================
// The window scan as a separate (non-inlined) function: with its own register allocation ptxas keeps all UNROLL
// 128-bit loads of a batch in flight before the first use; inlined into the large kernel it sinks the loads next to
// their uses to save registers and the scan runs ~45 % slower.  Results go through SHARED memory (8 words per warp):
// a by-reference struct would live in local memory, which misses the small L1 and then costs a DRAM-latency round
// trip per access while HBM is saturated.
template <int UNROLL>
__device__ __noinline__ void scan_job(const float* h, int n, float lo, float hi, float shift, int dirs, float t0, float t1,
                                      float* out) {
    const int lane = threadIdx.x & 31;
    sdc::ScanResult rs;
    scan_dispatch<UNROLL>(h, n, lo, hi, shift, dirs & 3, dirs >> 2, t0, t1, lane, rs);
    if (lane == 0) {
        out[0] = rs.s1; out[1] = rs.s2; out[2] = rs.ext[0]; out[3] = rs.ext[1];
        reinterpret_cast<int*>(out)[4] = rs.cnt[0] | (rs.cnt[1] << 16);
    }
    __syncwarp();
}

// Flag-tagged 8-byte words (value, step tag) for the lock-free hand-offs between warps: an aligned 8-byte store is
// atomic, so a reader that sees the current step's tag also sees the value -- no fences (a gpu-scope fence costs
// microseconds while HBM is saturated).
__device__ __forceinline__ void st_pair(uint2* p, uint32_t val, uint32_t tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(val), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_pair(const uint2* p) {
    uint2 r;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
enum { JP_ENV = 0, JP_N, JP_LO, JP_HI, JP_SHIFT, JP_DIRS, JP_THR0, JP_THR1, JP_WORDS = 8 };  // words of one queue record
enum { JR_S1 = 0, JR_S2, JR_EXT0, JR_EXT1, JR_CNTS, JR_WORDS = 8 };                     // job result words per env

template <int UNROLL>
__global__ void __launch_bounds__(kStepThreads, 2) k_step(const sdc::State S, const StepArgs a, const int n_unit_ctas) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int U = a.unit_envs;
    // location / dc-parameter tables -> shared memory (removes one level of pointer chasing per env)
    sdc::Tables T{S.loc, S.dc};
    {
        const int loc_bytes = S.n_loc * (int)sizeof(sdc::LocTables), dc_bytes = S.n_cfg * (int)sizeof(sdc_dc_params);
        if (loc_bytes + dc_bytes <= kTableBytes) {
            int* dst = reinterpret_cast<int*>(smem_raw);
            const int* src_loc = reinterpret_cast<const int*>(S.loc);
            const int* src_dc = reinterpret_cast<const int*>(S.dc);
            for (int i = threadIdx.x; i < loc_bytes / 4; i += kStepThreads) dst[i] = src_loc[i];
            for (int i = threadIdx.x; i < dc_bytes / 4; i += kStepThreads) dst[loc_bytes / 4 + i] = src_dc[i];
            T.loc = reinterpret_cast<const sdc::LocTables*>(smem_raw);
            T.dc = reinterpret_cast<const sdc_dc_params*>(smem_raw + loc_bytes);
        }
        __syncthreads();
    }
    __shared__ float scan_out[kWarpsPerBlock * 8];
    float* tile = reinterpret_cast<float*>(smem_raw + kTableBytes) + (size_t)warp * U * kListRow;     // this warp's bracket lists
    const int N = S.n_envs;
    const int n_units = (N + U - 1) / U;
    const uint32_t seq = (uint32_t)a.seq;
    const int total_jobs = (N / U) * max(U - a.local_jobs, 0) + max(N % U - a.local_jobs, 0);       // records the queue will hold
    if (blockIdx.x == 0 && threadIdx.x < 8) a.ctr_next[threadIdx.x] = 0;

    // ------------------------------------------------------------------------------------------------------------
    // Every warp of a unit CTA runs this small state machine until no unit and no scan job is left:
    //   produce  take a unit of U envs: scalar phase (one lane per env), then publish one window-scan job per env
    //   consume  take any published job (global queue, any env of any warp) and stream that env's window
    //   finish   once all jobs of the own unit have results: rewards, bracket write-back
    // Jobs are balanced over all warps of the chip; a warp never blocks on a single condition, so there is no
    // hold-and-wait cycle whatever the number of units per warp.
    // ------------------------------------------------------------------------------------------------------------
    if (a.phase_clocks && threadIdx.x == 0) atomicMin(a.phase_clocks + 8, gtime_ns());
    bool have_unit = false, units_left = blockIdx.x < n_unit_ctas, jobs_left = blockIdx.x < n_unit_ctas;
    int pending = -1;                                              // claimed job ticket not yet consumed
    int env0 = 0, n_here = 0;
    sdc::RewardInputs en;
    en.energy = 0.0; en.nci_next = 0.0; en.ls_penalty = 0.0;
    sdc::QView Q;
    Q.lst[0] = tile + lane * kListRow; Q.lst[1] = Q.lst[0] + sdc::kListCap;
    Q.a[0] = Q.a[1] = Q.m[0] = Q.m[1] = 0;
    sdc::ScanRequest rq;                                           // the own env's request (phase C needs n, dirs, shift, q1)
    rq.n = 0; rq.lo = rq.hi = rq.shift = 0.f; rq.dir[0] = rq.dir[1] = 0; rq.thr[0] = rq.thr[1] = 0.f; rq.degenerate = 0; rq.q1 = 0.0;
    sdc::ScanResult mine;                                          // results of the locally scanned envs (lane l <-> env l)
    mine.s1 = mine.s2 = 0.f; mine.cnt[0] = mine.cnt[1] = 0; mine.ext[0] = mine.ext[1] = 0.f;
    int unit_n_local = 0;
    long long clk_scan = 0, clk_idle = 0;
    int n_jobs_done = 0;

    while (have_unit || units_left || jobs_left || pending >= 0) {
        // ---------------- produce ----------------
        if (!have_unit && units_left) {
            int unit = 0;
            if (lane == 0) unit = atomicAdd(a.ctr + 0, 1);
            unit = __shfl_sync(0xffffffffu, unit, 0);
            if (unit >= n_units) { units_left = false; continue; }
            env0 = unit * U;
            const int env = env0 + lane;
            const bool active = lane < U && env < N;
            n_here = min(U, N - env0);
        const long long tk0 = clock64();
        // ---------------- scalar phase, one lane per env ----------------
        // (1) the unit's quartile brackets (one contiguous 8 KB block) start moving into shared memory asynchronously
        {
            const float4* src = reinterpret_cast<const float4*>(S.qlist + (size_t)env0 * 2 * sdc::kListCap);
            const int total4 = n_here * (2 * sdc::kListCap / 4);
#if defined(SDC_LIST_STAGING_SERIAL)
            for (int i = lane; i < total4; i += 32) {
                const float4 v = src[i];
                float* d = tile + (i >> 4) * kListRow + (i & 15) * 4;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
#elif defined(SDC_LIST_STAGING_CA)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = lane + 32 * j;
                if (i < total4) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + (i >> 4) * kListRow + (i & 15) * 4);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + i) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
#else
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = lane + 32 * j;
                if (i < total4) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + (i >> 4) * kListRow + (i & 15) * 4);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + i) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
#endif
        }
        int2 qa = make_int2(0, 0), qm = make_int2(0, 0);
        if (active) { qa = reinterpret_cast<const int2*>(S.q_a)[env]; qm = reinterpret_cast<const int2*>(S.q_m)[env]; }
        // (2) load shifting, data centre, battery -> the step's energy
        en.energy = 0.0; en.nci_next = 0.0; en.ls_penalty = 0.0;
        sdc::StepResult st;
        sdc::ObsDeferred od;
        st.terminal = 0;
        if (active) {
            prefetch_env(S, T, env);
            if (a.prefetch & 1) {
                const int len = S.hist_len[env];
                if (len >= 4 && lane < 2) l2_prefetch_bulk(S.hist + (size_t)env * S.hist_cap, (unsigned)((len * 4) & ~15));
            }
            const int a_ls = a.actions[env * 3 + 0], a_dc = a.actions[env * 3 + 1], a_bat = a.actions[env * 3 + 2];
            GlobalInfoSink info{a.info, N, env};
            sdc::physics_step(S, T, env, a_ls, a_dc, a_bat, info, st, od);
            en.energy = st.energy; en.nci_next = st.nci_next; en.ls_penalty = st.ls_penalty;
        }
        const long long tk1 = clock64();
        // (3) append the energy to the reward window, update the brackets, publish the window-scan jobs
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        Q.a[0] = Q.a[1] = Q.m[0] = Q.m[1] = 0;
        rq.n = 0; rq.dir[0] = rq.dir[1] = 0; rq.degenerate = 0; rq.q1 = 0.0; rq.shift = 0.f;
        if (active) {
            Q.a[0] = qa.x; Q.a[1] = qa.y; Q.m[0] = qm.x; Q.m[1] = qm.y;
            sdc::reward_prepare(S, env, en.energy, st.hist_len, st.hist_head, st.evicted, Q, rq);
        }
        // The first `local_jobs` envs of the unit are scanned by this warp right away (parameters stay in registers);
        // the others become records of the global job queue, which any warp of the chip consumes (load balance).
        const int n_local = min(a.local_jobs, n_here);
        unit_n_local = n_local;
        mine.s1 = mine.s2 = 0.f; mine.cnt[0] = mine.cnt[1] = 0; mine.ext[0] = mine.ext[1] = 0.f;
        {
            const bool queued = active && lane >= n_local;
            const unsigned act = __ballot_sync(0xffffffffu, queued);
            int base = 0;
            if (lane == 0 && act) base = atomicAdd(a.ctr + 4, __popc(act));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (queued) {
                uint2* jp = a.job_queue + (size_t)(base + __popc(act & ((1u << lane) - 1u))) * JP_WORDS;
                st_pair(jp + JP_ENV, (uint32_t)env, seq); st_pair(jp + JP_N, (uint32_t)rq.n, seq);
                st_pair(jp + JP_LO, __float_as_uint(rq.lo), seq); st_pair(jp + JP_HI, __float_as_uint(rq.hi), seq);
                st_pair(jp + JP_SHIFT, __float_as_uint(rq.shift), seq); st_pair(jp + JP_DIRS, (uint32_t)(rq.dir[0] | (rq.dir[1] << 2)), seq);
                st_pair(jp + JP_THR0, __float_as_uint(rq.thr[0]), seq); st_pair(jp + JP_THR1, __float_as_uint(rq.thr[1]), seq);
            }
        }
        __syncwarp();
        const long long tl0 = clock64();
        for (int l = 0; l < n_local; ++l) {
            const int n = __shfl_sync(0xffffffffu, rq.n, l);
            if (n < 2) continue;                                   // z = 0 (utils/reward_creator.py:26-27)
            const int dirs = __shfl_sync(0xffffffffu, rq.dir[0] | (rq.dir[1] << 2), l);
            float* rs_slot = scan_out + warp * 8;
            scan_job<UNROLL>(S.hist + (size_t)(env0 + l) * S.hist_cap, n, __shfl_sync(0xffffffffu, rq.lo, l),
                             __shfl_sync(0xffffffffu, rq.hi, l), __shfl_sync(0xffffffffu, rq.shift, l), dirs,
                             __shfl_sync(0xffffffffu, rq.thr[0], l), __shfl_sync(0xffffffffu, rq.thr[1], l), rs_slot);
            if (lane == l) {
                mine.s1 = rs_slot[0]; mine.s2 = rs_slot[1]; mine.ext[0] = rs_slot[2]; mine.ext[1] = rs_slot[3];
                const int c = reinterpret_cast<const int*>(rs_slot)[4];
                mine.cnt[0] = c & 0xffff; mine.cnt[1] = c >> 16;
            }
            __syncwarp();
            n_jobs_done += 1;
        }
        clk_scan += clock64() - tl0;
        __syncwarp();
        have_unit = true;
        const long long tk2 = clock64();
        // (4) off the queue's critical path: observations (written straight to obs / share / term_obs; the L2 merges the
        //     per-lane 4-byte stores into full sectors), logger sums, hand-over of finished envs to the reset workers
        {
            double m[13];
#pragma unroll
            for (int k = 0; k < 13; ++k) m[k] = 0.0;
            const int finished = st.terminal;
            if (active) {
                GlobalObsSink obs{a.obs + (size_t)env * kObsRow, a.share + (size_t)env * SDC_SHARE_DIM,
                                  (finished && a.term_obs) ? a.term_obs + (size_t)env * kObsRow : nullptr};
                sdc::emit_obs(S, T, env, od, obs);
                a.done[env] = (uint8_t)finished;
                m[0] = st.energy; m[1] = st.co2; m[2] = st.water; m[3] = st.tasks_in_queue; m[4] = st.tasks_dropped;
                m[5] = st.ite_kw; m[6] = st.ct_kw; m[7] = st.comp_kw; m[8] = st.hvac_kw; m[9] = 1.0; m[10] = st.terminal;
                m[11] = st.overdue; m[12] = st.total_kw;
            }
            // logger sums (harl/envs/sustaindc/sustaindc_logger.py:86-101): warp reduce, one atomic per metric
            constexpr int slot[13] = {sdc::M_ENERGY, sdc::M_CO2, sdc::M_WATER, sdc::M_TASKS_IN_QUEUE, sdc::M_TASKS_DROPPED,
                                      sdc::M_ITE_KW, sdc::M_CT_KW, sdc::M_COMP_KW, sdc::M_HVAC_KW, sdc::M_STEPS, sdc::M_EPISODES,
                                      sdc::M_OVERDUE, sdc::M_TOTAL_KW};
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                const double v = warp_sum(m[k]);
                if (lane == 0) atomicAdd(a.metrics + slot[k], v);
            }
            // Finished envs go to the reset workers (other CTAs of this launch).  Order matters: the terminal observation
            // is in global memory before the env is published, because the worker overwrites obs/share with the reset
            // ones.  Only the ~5 % of units that contain a finished env pay the gpu-scope fences.
            if (__any_sync(0xffffffffu, finished)) {
                __threadfence();
                if (finished) a.reset_list[atomicAdd(a.ctr + 1, 1)] = env;
                __threadfence();
                __syncwarp();
            }
            if (lane == 0) atomicAdd(a.ctr + 2, 1);
        }
        if (a.phase_clocks && lane == 0) {
            const long long tk2b = clock64();
            atomicAdd(a.phase_clocks + 0, (unsigned long long)(tk1 - tk0));   // load shifting + data centre + battery
            atomicAdd(a.phase_clocks + 1, (unsigned long long)(tk2 - tk1));   // window append, bracket update, publish (+ local scans)
            atomicAdd(a.phase_clocks + 13, (unsigned long long)(tk2b - tk2)); // deferred observations + metrics
            atomicAdd(a.phase_clocks + 4, 1ull);                              // units
            atomicMax(a.phase_clocks + 9, gtime_ns());
        }
        continue;
        }

        // ---------------- consume ----------------
        if (pending < 0 && jobs_left) {
            int j = 0;
            if (lane == 0) j = atomicAdd(a.ctr + 5, 1);
            j = __shfl_sync(0xffffffffu, j, 0);
            if (j >= total_jobs) jobs_left = false; else pending = j;
        }
        if (pending >= 0) {
            uint2 w = make_uint2(0u, seq);
            if (lane < JP_WORDS) w = ld_pair(a.job_queue + (size_t)pending * JP_WORDS + lane);       // one round trip: whole record
            if (__all_sync(0xffffffffu, w.y == seq)) {
                const long long tc0 = clock64();
                const int jenv = (int)__shfl_sync(0xffffffffu, w.x, JP_ENV);
                const int n = (int)__shfl_sync(0xffffffffu, w.x, JP_N);
                float* rs_slot = scan_out + warp * 8;
                if (n >= 2) {                                      // n < 2: z = 0 (utils/reward_creator.py:26-27)
                    scan_job<UNROLL>(S.hist + (size_t)jenv * S.hist_cap, n, __uint_as_float(__shfl_sync(0xffffffffu, w.x, JP_LO)),
                                     __uint_as_float(__shfl_sync(0xffffffffu, w.x, JP_HI)),
                                     __uint_as_float(__shfl_sync(0xffffffffu, w.x, JP_SHIFT)), (int)__shfl_sync(0xffffffffu, w.x, JP_DIRS),
                                     __uint_as_float(__shfl_sync(0xffffffffu, w.x, JP_THR0)),
                                     __uint_as_float(__shfl_sync(0xffffffffu, w.x, JP_THR1)), rs_slot);
                } else {
                    if (lane < 5) rs_slot[lane] = 0.f;
                    __syncwarp();
                }
                if (lane < 5) st_pair(a.job_results + (size_t)jenv * JR_WORDS + lane, __float_as_uint(rs_slot[lane]), seq);
                __syncwarp();
                pending = -1;
                clk_scan += clock64() - tc0; n_jobs_done += 1;
                if (a.phase_clocks && lane == 0) atomicMax(a.phase_clocks + 10, gtime_ns());
                continue;
            }
        }

        // ---------------- finish ----------------
        if (have_unit) {
            const int env = env0 + lane;
            const bool active = lane < U && env < N;
            uint2 r[5];
            bool ok = true;
            const bool remote = active && lane >= unit_n_local;
            if (remote) {
#pragma unroll
                for (int k = 0; k < 5; ++k) { r[k] = ld_pair(a.job_results + (size_t)env * JR_WORDS + k); ok = ok && r[k].y == seq; }
            }
            if (__all_sync(0xffffffffu, ok)) {
                const long long tk3 = clock64();
                double m_sum = 0.0, m_ls = 0.0, m_dc = 0.0;
                if (active) {
                    if (remote) {
                        mine.s1 = __uint_as_float(r[JR_S1].x); mine.s2 = __uint_as_float(r[JR_S2].x);
                        mine.ext[0] = __uint_as_float(r[JR_EXT0].x); mine.ext[1] = __uint_as_float(r[JR_EXT1].x);
                        mine.cnt[0] = (int)(r[JR_CNTS].x & 0xffffu); mine.cnt[1] = (int)(r[JR_CNTS].x >> 16);
                    }
                    float r3[3];
                    sdc::reward_finish(S, env, rq, mine, en, Q, r3);
                    reinterpret_cast<int2*>(S.q_a)[env] = make_int2(Q.a[0], Q.a[1]);
                    reinterpret_cast<int2*>(S.q_m)[env] = make_int2(Q.m[0], Q.m[1]);
                    a.rew[env * 3 + 0] = r3[0]; a.rew[env * 3 + 1] = r3[1]; a.rew[env * 3 + 2] = r3[2];
                    m_sum = (double)r3[0] + r3[1] + r3[2]; m_ls = r3[0]; m_dc = r3[1];
                }
                m_sum = warp_sum(m_sum); m_ls = warp_sum(m_ls); m_dc = warp_sum(m_dc);
                if (lane == 0) {
                    atomicAdd(a.metrics + sdc::M_REWARD_SUM, m_sum); atomicAdd(a.metrics + sdc::M_REWARD_LS, m_ls);
                    atomicAdd(a.metrics + sdc::M_REWARD_DC, m_dc);
                }
                __syncwarp();
                {
                    float4* dst = reinterpret_cast<float4*>(S.qlist + (size_t)env0 * 2 * sdc::kListCap);
                    const int total4 = n_here * (2 * sdc::kListCap / 4);
                    for (int i = lane; i < total4; i += 32) {
                        const int e = i >> 4, k = (i & 15) * 4;
                        const float* d = tile + e * kListRow + k;
                        dst[i] = make_float4(d[0], d[1], d[2], d[3]);
                    }
                }
                __syncwarp();
                have_unit = false;
                if (a.phase_clocks && lane == 0) { atomicAdd(a.phase_clocks + 3, (unsigned long long)(clock64() - tk3)); atomicMax(a.phase_clocks + 11, gtime_ns()); }
                continue;
            }
        }
        // nothing to do right now: the claimed job is not published yet and the own unit is still being scanned elsewhere
        { const long long ti = clock64(); __nanosleep(100); clk_idle += clock64() - ti; }
    }
    if (a.phase_clocks && lane == 0) {
        atomicAdd(a.phase_clocks + 2, (unsigned long long)clk_scan);      // window scans (all jobs this warp consumed)
        atomicAdd(a.phase_clocks + 5, (unsigned long long)clk_idle);      // waiting
        atomicAdd(a.phase_clocks + 6, (unsigned long long)n_jobs_done);
        atomicAdd(a.phase_clocks + 7, 1ull);                              // warps
    }

    // ---------------- episode resets: every CTA turns into a reset worker once it has no unit left ----------------
    // CTAs beyond n_unit_ctas start here immediately, so resets of envs that finished in this step overlap with the
    // window scans of the other CTAs.  A worker claims the next slot of reset_list and waits until it is filled or
    // until every unit is past phase A (then no further env can be appended).
    __shared__ ResetShared rsh;
    __shared__ int s_env;
    __syncthreads();
    double* runbuf = reinterpret_cast<double*>(smem_raw + kTableBytes);   // the warps' scratch tiles are free by now
    for (;;) {
        if (threadIdx.x == 0) {
            const int my = atomicAdd(a.ctr + 3, 1);
            volatile int32_t* list = a.reset_list;
            volatile int32_t* units_done = a.ctr + 2;
            int env = -1;
            for (;;) {
                env = list[my];
                if (env >= 0) break;
                if (*units_done >= n_units) { __threadfence(); env = list[my]; break; }
                __nanosleep(200);
            }
            if (env >= 0) list[my] = -1;
            s_env = env;
        }
        __syncthreads();
        const int env = s_env;
        if (env < 0) break;
        __threadfence();
        reset_one_env(S, env, a.obs, a.share, runbuf, rsh);
        __syncthreads();
    }
    if (a.phase_clocks && threadIdx.x == 0) atomicMax(a.phase_clocks + 12, gtime_ns());
}

__global__ void k_build_reset_list(int n_envs, const uint8_t* __restrict__ mask, int32_t* list, int32_t* count) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n_envs) return;
    if (!mask) { list[env] = env; if (env == 0) *count = n_envs; return; }
    if (mask[env]) list[atomicAdd(count, 1)] = env;
}

// =================================================================================================
// k_rebuild: exact brackets from a full sort (prefill / resume / debug cross-check)
// =================================================================================================
constexpr int kSortThreads = 512;
__global__ void __launch_bounds__(kSortThreads) k_rebuild(const sdc::State S) {
    extern __shared__ float buf[];                  // next pow2 >= hist_cap floats
    const int env = blockIdx.x;
    const int n = S.hist_len[env];
    int p2 = 1; while (p2 < n) p2 <<= 1;
    if (p2 < 2) p2 = 2;
    const float* h = S.hist + (size_t)env * S.hist_cap;
    for (int i = threadIdx.x; i < p2; i += kSortThreads) buf[i] = i < n ? h[i] : SDC_INF_F;
    __syncthreads();
    for (int k = 2; k <= p2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < p2; i += kSortThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float x = buf[i], y = buf[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { buf[i] = y; buf[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x < 2) {
        const int j = threadIdx.x;
        float* lst = S.qlist + ((size_t)env * 2 + j) * sdc::kListCap;
        int a = 0, m = 0;
        if (n > 0) {
            const int k = ((j == 0 ? 1 : 3) * (n - 1)) / 4;
            a = k - (sdc::kListCap / 2 - 1);
            if (a + sdc::kListCap > n) a = n - sdc::kListCap;
            if (a < 0) a = 0;
            m = n - a < sdc::kListCap ? n - a : sdc::kListCap;
            for (int i = 0; i < m; ++i) lst[i] = buf[a + i];
        }
        S.q_a[env * 2 + j] = a; S.q_m[env * 2 + j] = m;
    }
}

// =================================================================================================
// launches
// =================================================================================================
static const char* launch_step(Context& c, const sdc::State& S, const StepArgs& a, void* stream) {
    const int U = a.unit_envs;
    size_t smem = (size_t)kWarpsPerBlock * U * kListRow * sizeof(float);
    if (smem < kNormWindow * sizeof(double)) smem = kNormWindow * sizeof(double);      // reset workers reuse the tile region
    smem += kTableBytes;
    const int n_units = (S.n_envs + U - 1) / U;
    const int bps = a.blocks_per_sm > 0 ? a.blocks_per_sm : c.step_blocks_per_sm;
    // All CTAs must be co-resident (reset workers wait for unit CTAs): never more than the resident capacity.
    const int capacity = c.sm_count * (bps < 2 ? bps : 2);
    const int need = (n_units + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int reserve = capacity / 8;                       // CTAs that only do resets (they overlap with the scans)
    if (reserve < 1) reserve = 1;
    int n_unit_ctas = need < capacity - reserve ? need : capacity - reserve;
    if (n_unit_ctas < 1) n_unit_ctas = 1;
    int blocks = n_unit_ctas + reserve;
    if (need < capacity - reserve && blocks < capacity) {
        // small batches: spare CTAs cost nothing; cap so that a handful of envs does not launch a whole grid
        const int want = n_unit_ctas + (S.n_envs < 4096 ? 4 : reserve);
        blocks = want < capacity ? want : capacity;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c.device));
    if (a.unroll == 4) k_step<4><<<blocks, kStepThreads, smem, st>>>(S, a, n_unit_ctas);
    else if (a.unroll == 16) k_step<16><<<blocks, kStepThreads, smem, st>>>(S, a, n_unit_ctas);
    else k_step<8><<<blocks, kStepThreads, smem, st>>>(S, a, n_unit_ctas);
    CU(cudaGetLastError());
    return nullptr;
}

static const char* launch_reset(Context& c, const sdc::State& S, const int32_t* list, const int32_t* count, float* obs, float* share,
                                void* stream) {
    const size_t smem = kNormWindow * sizeof(double);
    CU(cudaSetDevice(c.device));
    int blocks = c.sm_count;
    if (blocks > S.n_envs) blocks = S.n_envs;
    k_reset<<<blocks, kResetThreads, smem, (cudaStream_t)stream>>>(S, list, count, obs, share);
    CU(cudaGetLastError());
    return nullptr;
}

static const char* launch_build_reset_list(Context&, const sdc::State& S, const uint8_t* mask, int32_t* list, int32_t* count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    k_build_reset_list<<<(S.n_envs + 255) / 256, 256, 0, st>>>(S.n_envs, mask, list, count);
    CU(cudaGetLastError());
    return nullptr;
}

static const char* launch_rebuild(Context&, const sdc::State& S, void* stream) {
    int p2 = 1; while (p2 < S.hist_cap) p2 <<= 1;
    const size_t smem = (size_t)p2 * sizeof(float);
    k_rebuild<<<S.n_envs, kSortThreads, smem, (cudaStream_t)stream>>>(S);
    CU(cudaGetLastError());
    return nullptr;
}

static const char* set_kernel_attributes() {
    CU(cudaFuncSetAttribute(k_step<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CU(cudaFuncSetAttribute(k_step<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CU(cudaFuncSetAttribute(k_step<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CU(cudaFuncSetAttribute(k_rebuild, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    return nullptr;
}

}  // namespace backend

#include "sdc_api.inc"
