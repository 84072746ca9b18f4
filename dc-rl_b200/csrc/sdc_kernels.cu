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
static const char* sync(Context& c) { CU(cudaSetDevice(c.device)); CU(cudaDeviceSynchronize()); return nullptr; }

// =================================================================================================
// device helpers
// =================================================================================================
constexpr int kStepThreads = 256;
constexpr int kWarpsPerBlock = kStepThreads / 32;
constexpr int kObsRow = 3 * SDC_OBS_DIM;          // 78 floats per env
constexpr int kObsRowPad = kObsRow + 1;           // odd stride: conflict-free one-lane-per-row writes

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

struct SmemObsSink {
    float* row;
    __device__ __forceinline__ void operator()(int agent, int idx, float v) { row[agent * SDC_OBS_DIM + idx] = v; }
};
struct GlobalInfoSink {
    float* info; int n, env;
    __device__ __forceinline__ void operator()(int col, float v) { if (info) info[(size_t)col * n + env] = v; }
};

// ---- phase B: one warp streams one env's window -------------------------------------------------
struct Acc {
    float s1[4], s2[4];
    int cb[2], ca[2];
    float pred[2], succ[2];
};

template <bool WIDEN>
__device__ __forceinline__ void acc_one(Acc& A, int k, float x, float lo, float hi, float shift, const float* below, const float* above) {
    const float c = fminf(fmaxf(x, lo), hi);
    const float d = c - shift;
    A.s1[k] += d;
    A.s2[k] = fmaf(d, d, A.s2[k]);
    if (WIDEN) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const bool b = x < below[j];
            A.cb[j] += b;
            A.pred[j] = fmaxf(A.pred[j], b ? x : -SDC_INF_F);
            const bool u = x > above[j];
            A.ca[j] += u;
            A.succ[j] = fminf(A.succ[j], u ? x : SDC_INF_F);
        }
    }
}

template <bool WIDEN, int UNROLL>
__device__ __forceinline__ void scan_window(const float* __restrict__ h, int n, float lo, float hi, float shift, const float* below,
                                            const float* above, int lane, sdc::ScanResult& rs) {
    Acc A;
#pragma unroll
    for (int k = 0; k < 4; ++k) { A.s1[k] = 0.f; A.s2[k] = 0.f; }
#pragma unroll
    for (int j = 0; j < 2; ++j) { A.cb[j] = 0; A.ca[j] = 0; A.pred[j] = -SDC_INF_F; A.succ[j] = SDC_INF_F; }
    const float4* p = reinterpret_cast<const float4*>(h);
    const int n4 = n >> 2;
    int c = lane;
    for (; c + (UNROLL - 1) * 32 < n4; c += UNROLL * 32) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(p + c + u * 32);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            acc_one<WIDEN>(A, 0, v[u].x, lo, hi, shift, below, above);
            acc_one<WIDEN>(A, 1, v[u].y, lo, hi, shift, below, above);
            acc_one<WIDEN>(A, 2, v[u].z, lo, hi, shift, below, above);
            acc_one<WIDEN>(A, 3, v[u].w, lo, hi, shift, below, above);
        }
    }
    for (; c < n4; c += 32) {
        const float4 v = __ldcs(p + c);
        acc_one<WIDEN>(A, 0, v.x, lo, hi, shift, below, above);
        acc_one<WIDEN>(A, 1, v.y, lo, hi, shift, below, above);
        acc_one<WIDEN>(A, 2, v.z, lo, hi, shift, below, above);
        acc_one<WIDEN>(A, 3, v.w, lo, hi, shift, below, above);
    }
    const int rem = n & 3;
    if (lane < rem) acc_one<WIDEN>(A, 0, __ldcs(h + (n4 << 2) + lane), lo, hi, shift, below, above);
    rs.s1 = warp_sum((A.s1[0] + A.s1[1]) + (A.s1[2] + A.s1[3]));
    rs.s2 = warp_sum((A.s2[0] + A.s2[1]) + (A.s2[2] + A.s2[3]));
    if (WIDEN) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            rs.cnt_below[j] = warp_sum(A.cb[j]); rs.cnt_above[j] = warp_sum(A.ca[j]);
            rs.pred[j] = warp_max(A.pred[j]); rs.succ[j] = warp_min(A.succ[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) { rs.cnt_below[j] = 0; rs.cnt_above[j] = 0; rs.pred[j] = -SDC_INF_F; rs.succ[j] = SDC_INF_F; }
    }
}

// =================================================================================================
// k_step
// =================================================================================================
template <int UNROLL>
__global__ void __launch_bounds__(kStepThreads, 2) k_step(const sdc::State S, const StepArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int U = a.unit_envs;
    float* tile = smem + (size_t)warp * U * kObsRowPad;          // this warp's [U][79] observation tile
    const int N = S.n_envs;
    const int n_units = (N + U - 1) / U;
    if (blockIdx.x == 0 && threadIdx.x == 0) { *a.ticket_next = 0; *a.reset_count_next = 0; }

    for (;;) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(a.ticket, 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_units) break;
        const int env0 = unit * U;
        const int env = env0 + lane;
        const bool active = lane < U && env < N;
        const int n_here = min(U, N - env0);

        // ---------------- phase A: one lane per env ----------------
        sdc::RewardInputs en;                                       // what phase C needs from phase A
        sdc::ScanRequest rq;
        rq.n = 0; rq.widen = 0; rq.lo = rq.hi = rq.shift = 0.f; rq.degenerate = 0; rq.q1 = rq.q3 = 0.0;
        rq.below[0] = rq.below[1] = -SDC_INF_F; rq.above[0] = rq.above[1] = SDC_INF_F;
        en.energy = 0.0; en.nci_next = 0.0; en.ls_penalty = 0.0;
        {
            double m[13];
#pragma unroll
            for (int k = 0; k < 13; ++k) m[k] = 0.0;
            if (active) {
                if (a.prefetch) {
                    // start pulling this env's window towards L2 while the scalar physics runs
                    const int len = S.hist_len[env];
                    if (len >= 4 && lane < 2) l2_prefetch_bulk(S.hist + (size_t)env * S.hist_cap, (unsigned)((len * 4) & ~15));
                }
                const int a_ls = a.actions[env * 3 + 0], a_dc = a.actions[env * 3 + 1], a_bat = a.actions[env * 3 + 2];
                SmemObsSink obs{tile + lane * kObsRowPad};
                GlobalInfoSink info{a.info, N, env};
                sdc::StepResult st;
                sdc::physics_step(S, env, a_ls, a_dc, a_bat, obs, info, st);
                sdc::reward_prepare(S, env, st.energy, rq);
                en.energy = st.energy; en.nci_next = st.nci_next; en.ls_penalty = st.ls_penalty;
                a.done[env] = (uint8_t)st.terminal;
                if (st.terminal) {
                    a.reset_list[atomicAdd(a.reset_count, 1)] = env;
                    if (a.term_obs) {
                        float* dst = a.term_obs + (size_t)env * kObsRow;
                        for (int k = 0; k < kObsRow; ++k) dst[k] = obs.row[k];
                    }
                }
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
        }
        __syncwarp();
        // observation tile -> global, coalesced (rows of a unit are contiguous in obs[N,3,26])
        {
            float2* dst = reinterpret_cast<float2*>(a.obs + (size_t)env0 * kObsRow);
            const int total2 = n_here * (kObsRow / 2);
            for (int i = lane; i < total2; i += 32) {
                const int e = i / (kObsRow / 2), k = (i - e * (kObsRow / 2)) * 2;
                dst[i] = make_float2(tile[e * kObsRowPad + k], tile[e * kObsRowPad + k + 1]);
            }
            float* sh = a.share + (size_t)env0 * SDC_SHARE_DIM;
            const int total = n_here * SDC_SHARE_DIM;
            for (int i = lane; i < total; i += 32) {
                const int e = i / SDC_SHARE_DIM, k = i - e * SDC_SHARE_DIM;
                // ls[0:26] | dc[11] | dc[13] | padded battery row [25]   (harlsustaindc_env.py:78-85)
                const int src = k < 26 ? k : (k == 26 ? SDC_OBS_DIM + 11 : (k == 27 ? SDC_OBS_DIM + 13 : 2 * SDC_OBS_DIM + 25));
                sh[i] = tile[e * kObsRowPad + src];
            }
        }
        __syncwarp();

        // ---------------- phase B: the warp streams each env's reward window ----------------
        sdc::ScanResult mine;
        mine.s1 = mine.s2 = 0.f;
        mine.cnt_below[0] = mine.cnt_below[1] = mine.cnt_above[0] = mine.cnt_above[1] = 0;
        mine.pred[0] = mine.pred[1] = -SDC_INF_F; mine.succ[0] = mine.succ[1] = SDC_INF_F;
        for (int l = 0; l < n_here; ++l) {
            const int n = __shfl_sync(0xffffffffu, rq.n, l);
            if (a.prefetch && l + 2 < n_here) {
                const int n2 = __shfl_sync(0xffffffffu, rq.n, l + 2);
                if (lane == 0 && n2 >= 4) l2_prefetch_bulk(S.hist + (size_t)(env0 + l + 2) * S.hist_cap, (unsigned)((n2 * 4) & ~15));
            }
            if (n < 2) continue;                                   // z = 0 (utils/reward_creator.py:26-27)
            const float lo = __shfl_sync(0xffffffffu, rq.lo, l), hi = __shfl_sync(0xffffffffu, rq.hi, l);
            const float shift = __shfl_sync(0xffffffffu, rq.shift, l);
            const int widen = __shfl_sync(0xffffffffu, rq.widen, l);
            const float* h = S.hist + (size_t)(env0 + l) * S.hist_cap;
            sdc::ScanResult rs;
            if (widen) {
                float below[2], above[2];
                below[0] = __shfl_sync(0xffffffffu, rq.below[0], l); below[1] = __shfl_sync(0xffffffffu, rq.below[1], l);
                above[0] = __shfl_sync(0xffffffffu, rq.above[0], l); above[1] = __shfl_sync(0xffffffffu, rq.above[1], l);
                scan_window<true, 4>(h, n, lo, hi, shift, below, above, lane, rs);
            } else {
                scan_window<false, UNROLL>(h, n, lo, hi, shift, nullptr, nullptr, lane, rs);
            }
            if (lane == l) mine = rs;
        }

        // ---------------- phase C: one lane per env ----------------
        double m_sum = 0.0, m_ls = 0.0, m_dc = 0.0;
        if (active) {
            float r3[3];
            sdc::reward_finish(S, env, rq, mine, en, r3);
            a.rew[env * 3 + 0] = r3[0]; a.rew[env * 3 + 1] = r3[1]; a.rew[env * 3 + 2] = r3[2];
            m_sum = (double)r3[0] + r3[1] + r3[2]; m_ls = r3[0]; m_dc = r3[1];
        }
        m_sum = warp_sum(m_sum); m_ls = warp_sum(m_ls); m_dc = warp_sum(m_dc);
        if (lane == 0) {
            atomicAdd(a.metrics + sdc::M_REWARD_SUM, m_sum); atomicAdd(a.metrics + sdc::M_REWARD_LS, m_ls);
            atomicAdd(a.metrics + sdc::M_REWARD_DC, m_dc);
        }
        __syncwarp();
    }
}

// =================================================================================================
// k_reset
// =================================================================================================
constexpr int kResetThreads = sdc::kNoiseThreads;   // 256

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

__global__ void __launch_bounds__(kResetThreads) k_reset(const sdc::State S, const int32_t* __restrict__ list,
                                                         const int32_t* __restrict__ count, float* obs, float* share) {
    extern __shared__ float inc[];                  // [256*140] random-walk increments of one env
    __shared__ double red[kResetThreads / 32];
    __shared__ double seg_off[kResetThreads];
    __shared__ int s_start[3];
    __shared__ float s_row[kObsRow];
    const int tid = threadIdx.x;
    const int n = SDC_YEAR_STEPS;
    const int total = *count;
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
        const int env = list[i];
        double* wt = S.weather + (size_t)env * 2 * S.win_len;
        double* ww = wt + S.win_len;
        const bool staged = S.pend_valid && S.pend_valid[env];
        __syncthreads();
        if (staged) {
            if (tid == 0) { s_start[0] = S.pend_day[env]; s_start[1] = S.pend_hour[env]; s_start[2] = 0; }
            const double* src = S.pend_weather + (size_t)env * 2 * S.win_len;
            for (int k = tid; k < 2 * S.win_len; k += kResetThreads) wt[k] = src[k];
            if (tid == 0) { S.t_min[env] = S.pend_tmin[env]; S.t_max[env] = S.pend_tmax[env]; }
        } else {
            const uint32_t ep = S.episode[env];
            const uint64_t seed = S.seed[env];
            if (tid == 0) sdc::draw_episode_start(seed, ep, S.day_lo[env], S.day_hi[env], &s_start[0], &s_start[1], &s_start[2]);
            // pass 1: increments of this thread's segment (utils/managers.py:45-46)
            double seg = 0.0;
            for (int q = 0; q < sdc::kNoiseSeg / 4; ++q) {
                float z[4];
                const int j0 = tid * sdc::kNoiseSeg + q * 4;
                sdc::noise_normals4(seed, ep, (uint32_t)(j0 >> 2), z);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float v = 0.02f * z[k];
                    inc[j0 + k] = v;
                    if (j0 + k < n) seg += (double)v;
                }
            }
            seg_off[tid] = seg;
            __syncthreads();
            if (tid == 0) {                           // serial exclusive prefix, same order as the host statement
                double acc = 0.0;
                for (int k = 0; k < kResetThreads; ++k) { const double s = seg_off[k]; seg_off[k] = acc; acc += s; }
            }
            __syncthreads();
            const double off = seg_off[tid];
            const int j_lo = tid * sdc::kNoiseSeg, j_hi = min(j_lo + sdc::kNoiseSeg, n);
            // pass 2 / 3: mean and population std of the walk
            double run = 0.0, sum = 0.0;
            for (int j = j_lo; j < j_hi; ++j) { run += (double)inc[j]; sum += off + run; }
            const double mean = block_sum(sum, red) / n;
            run = 0.0; double ss = 0.0;
            for (int j = j_lo; j < j_hi; ++j) { run += (double)inc[j]; const double d = (off + run) - mean; ss += d * d; }
            const double scale = 0.75 / sqrt(block_sum(ss, red) / n);          // managers.py:46-48
            // pass 4: roll, clip, window, 30-day min/max (managers.py:598-608)
            const int t0 = s_start[0] * 96 + s_start[1] * 4, roll = s_start[2];
            for (int k = tid; k < 2 * S.win_len; k += kResetThreads) wt[k] = 0.0;
            __syncthreads();
            const sdc::LocTables& L = S.loc[S.loc_id[env]];
            double tmin = INFINITY, tmax = -INFINITY;
            run = 0.0;
            for (int j = j_lo; j < j_hi; ++j) {
                run += (double)inc[j];
                int t = j + 96 * roll; if (t >= n) t -= n;
                if (t < t0) continue;
                const double noise = (off + run) * scale;
                const double vt = fmin(fmax(L.temp_base[j] + noise, 0.0), 45.0);
                if (t < t0 + 2880) { tmin = fmin(tmin, vt); tmax = fmax(tmax, vt); }
                if (t < t0 + S.win_len) {
                    wt[t - t0] = vt;
                    ww[t - t0] = fmin(fmax(L.wetb_base[j] + noise, 0.0), 45.0);
                }
            }
            tmin = block_minmax(tmin, true, red);
            tmax = block_minmax(tmax, false, red);
            if (tid == 0) { S.t_min[env] = tmin; S.t_max[env] = tmax; }
        }
        uint8_t* ring = S.ls_ring + (size_t)env * (S.ls_mask + 1);
        for (int k = tid; k <= S.ls_mask; k += kResetThreads) ring[k] = 0;
        __syncthreads();                               // weather window + norms visible to thread 0
        if (tid == 0) {
            if (staged) S.pend_valid[env] = 0;
            S.episode[env] += 1;
            RowSink sink{s_row};
            sdc::reset_scalar_state(S, env, s_start[0] * 96 + s_start[1] * 4, sink);
        }
        __syncthreads();
        for (int k = tid; k < kObsRow; k += kResetThreads) obs[(size_t)env * kObsRow + k] = s_row[k];
        if (tid < SDC_SHARE_DIM) {
            const int k = tid;
            const int src = k < 26 ? k : (k == 26 ? SDC_OBS_DIM + 11 : (k == 27 ? SDC_OBS_DIM + 13 : 2 * SDC_OBS_DIM + 25));
            share[(size_t)env * SDC_SHARE_DIM + k] = s_row[src];
        }
    }
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
    const size_t smem = (size_t)kWarpsPerBlock * U * kObsRowPad * sizeof(float);
    const int n_units = (S.n_envs + U - 1) / U;
    const int bps = a.blocks_per_sm > 0 ? a.blocks_per_sm : c.step_blocks_per_sm;
    int blocks = c.sm_count * bps;
    const int need = (n_units + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c.device));
    if (a.unroll == 4) k_step<4><<<blocks, kStepThreads, smem, st>>>(S, a);
    else if (a.unroll == 16) k_step<16><<<blocks, kStepThreads, smem, st>>>(S, a);
    else k_step<8><<<blocks, kStepThreads, smem, st>>>(S, a);
    CU(cudaGetLastError());
    return nullptr;
}

static const char* launch_reset(Context& c, const sdc::State& S, const int32_t* list, const int32_t* count, float* obs, float* share,
                                void* stream) {
    const size_t smem = (size_t)sdc::kNoiseThreads * sdc::kNoiseSeg * sizeof(float);
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
    CU(cudaFuncSetAttribute(k_step<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(k_step<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(k_step<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(sdc::kNoiseThreads * sdc::kNoiseSeg * sizeof(float))));
    CU(cudaFuncSetAttribute(k_rebuild, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    return nullptr;
}

}  // namespace backend

#include "sdc_api.inc"
