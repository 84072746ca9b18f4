// sdc_kernels.cu -- CUDA (sm_100a) backend of libsdc_b200.so.
//
// Kernels
//   k_step    ONE cooperative launch per env-step for all N envs (grid = what the device holds at once: 2 CTAs of 256 threads
//             per SM; parameters are __grid_constant__ so that the State struct is read from the constant bank, not copied
//             to every thread's stack).
//             UNITS: a warp takes a unit of U consecutive envs, one lane per env -- load-shifting queue, IT / HVAC model,
//             battery, trace gathers, info row (sdc_core.h, fp64 like the reference), then the reward normaliser
//             INCREMENTALLY (window append, exact rolling quartile brackets, fp64 window moments, sorted tail bands around
//             the IQR fences -> clipped mean / std without touching the 40 KB window; samples that enter / leave inside a
//             bracket or a band are located and edited by the whole warp), rewards, logger sums (one 16-value warp
//             reduction), the 29 distinct observation values through a shared-memory tile of compact rows, and the reset of
//             envs that finish (the look-ahead generation staged the next episode with its observation: buffer flip + copy).
//             No unit warp ever waits for another warp, CTA or job.
//             WORKERS: CTAs without units from the first cycle, and every CTA once its units are done, serve a job queue --
//             window passes (an env whose incremental state runs out of slack: its window is staged in shared memory by one
//             TMA bulk copy, scanned by 256 threads, bucket-sorted, committed; if the step itself cannot be priced without
//             the window, the pass CTA prices it: finish_step) and the look-ahead generation of the episodes that start two
//             steps from now (year-long weather random walk from counter-based RNG streams + the reset observation).
//   k_reset   explicit resets, one CTA per env: start day/hour, weather generation, queue clear, reset observation (or
//             copies a staged episode).
//   k_rebuild one CTA per env: full bitonic sort of the window in shared memory -> fresh brackets (set-up / restore).
//   k_build_reset_list  mask -> env list.
//
// No tensor cores: there is no dense contraction on this path (per-env scalar state machines + streaming passes).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include "sdc_core.h"

namespace backend {

struct Context {
    int device = 0;
    int sm_count = 148;
    bool cooperative = true;      // cudaLaunchCooperativeKernel for k_step (co-residency guaranteed by the driver)
    int occ_per_sm[1] = {0};      // cached occupancy query of k_step ...
    size_t occ_smem[1] = {0};     // ... at this much dynamic shared memory
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
    int coop = 0;
    CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) return "libsdc_b200 needs cooperative kernel launch (worker CTAs wait on unit CTAs of the same grid)";
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
static const char* pinned_alloc(Context&, void** p, size_t bytes) { CU(cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocMapped | cudaHostAllocPortable)); return nullptr; }
static void pinned_free(Context&, void* p) { cudaFreeHost(p); }
static bool host_memory_is_device_visible(Context& c) {
    int uva = 0;
    return cudaDeviceGetAttribute(&uva, cudaDevAttrUnifiedAddressing, c.device) == cudaSuccess && uva != 0;
}
static const char* stream_create(Context&, void** s) { cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); *s = st; return nullptr; }
static const char* stream_sync(Context&, void* s) { CU(cudaStreamSynchronize((cudaStream_t)s)); return nullptr; }
static const char* event_create(Context&, void** ev) { cudaEvent_t e; CU(cudaEventCreate(&e)); *ev = e; return nullptr; }
static const char* event_record(Context&, void* ev, void* st) { CU(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)st)); return nullptr; }
static const char* event_elapsed_ms(Context&, void* a, void* b, double* ms) {
    float f = 0.f; CU(cudaEventElapsedTime(&f, (cudaEvent_t)a, (cudaEvent_t)b)); *ms = f; return nullptr;
}
static void event_destroy(Context&, void* ev) { cudaEventDestroy((cudaEvent_t)ev); }
static const char* dev_fill_bytes(Context&, void* p, int v, size_t bytes) { CU(cudaMemset(p, v, bytes)); return nullptr; }
static const char* sync(Context& c) { CU(cudaSetDevice(c.device)); CU(cudaDeviceSynchronize()); return nullptr; }
static void range_push(const char* name) { nvtxRangePushA(name); }      // NVTX ranges around the C-ABI calls (nsys / ncu --nvtx)
static void range_pop() { nvtxRangePop(); }

// =================================================================================================
// device helpers
// =================================================================================================
constexpr int kStepThreads = 256;
constexpr int kWarpsPerBlock = kStepThreads / 32;
constexpr int kObsRow = 3 * SDC_OBS_DIM;          // 78 floats per env
constexpr int kTileStride = SDC_OBS_COMPACT;      // the shared-memory observation tile holds COMPACT rows (29 floats: odd stride, conflict-free)
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
// Sums of 16 values over the warp with 16 shuffles instead of 16 x 5: every exchange step halves the number of values a lane
// carries (it keeps the half its lane bit selects and receives the partner's copy of that half).  On return lane l holds the
// warp sum of value (l >> 1) bit-reversed over four bits -- warp_sum16_slot(l) -- in both lanes of a pair.
__device__ __forceinline__ double warp_sum16(double (&v)[16], int lane) {
#pragma unroll
    for (int step = 0; step < 4; ++step) {
        const int half = 8 >> step;                    // values kept after this step
        const bool up = (lane >> (4 - step)) & 1;      // lane bit 4, 3, 2, 1
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16 >> step);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// which of the 16 values lane `lane` ends up with: bit 4 of the lane selects the upper 8, bit 3 the upper 4 of those, ...
__device__ __forceinline__ int warp_sum16_slot(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
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
// Called by the whole warp (`active` lanes own an env).  what: bit 0 the reads of the physics, bit 1 those of the normaliser.
__device__ __forceinline__ void prefetch_env(const sdc::State& S, const sdc::Tables& T, int env, bool active, int what) {
    if (!active) env = 0;
    if (what & 1) {
    const int t = S.t[env], t0 = S.t0[env], head = S.ls_head[env], hh = S.hist_head[env];
    const sdc::LocTables& L = T.loc[S.loc_id[env]];
    const double* wt = sdc::weather_cur(S, env) + (t - t0);
    prefetch_line(wt); prefetch_line(wt + 16); prefetch_line(wt + S.win_len);
    const uint8_t* ring = S.ls_ring + (size_t)env * (S.ls_mask + 1);
    prefetch_line(ring + ((t - 24) & S.ls_mask)); prefetch_line(ring + ((t - 48) & S.ls_mask));
    prefetch_line(ring + ((t - 72) & S.ls_mask)); prefetch_line(ring + ((t - 96) & S.ls_mask));
    prefetch_line(ring + (t & S.ls_mask)); prefetch_line(ring + (head & S.ls_mask));
    prefetch_line(S.hist + (size_t)env * S.hist_cap + hh);
    prefetch_line(L.ci + t - 16); prefetch_line(L.ci + t); prefetch_line(L.ci + t + 9);
    prefetch_line(L.workload + t); prefetch_line(L.ns + t); prefetch_line(L.sh + t);
    if (S.step_in_ep[env] + 1 >= S.ep_len) {          // the env finishes in this step: its reset reads the staged episode
        const float* po = S.pend_obs + (size_t)env * kObsRow;
        prefetch_line(po); prefetch_line(po + 32); prefetch_line(po + 64); prefetch_line(po + kObsRow - 1);
        prefetch_line(S.pend_tmin + env); prefetch_line(S.pend_tmax + env); prefetch_line(S.pend_day + env); prefetch_line(S.pend_hour + env);
        prefetch_line(S.pend_valid + env); prefetch_line(S.episode + env);
    }
    }
    if (!(what & 2)) return;
    // reward normaliser: the bracket positions it will look at and the rows of the two tail bands
    const int2 qa = reinterpret_cast<const int2*>(S.q_a)[env], qm = reinterpret_cast<const int2*>(S.q_m)[env];
    const int2 tn = reinterpret_cast<const int2*>(S.tail_n)[env];
    const int n_after = min(S.hist_len[env] + 1, S.hist_cap);
    const float* l0 = S.qlist + (size_t)env * 2 * sdc::kListCap;
    prefetch_line(l0); prefetch_line(l0 + max(qm.x - 1, 0)); prefetch_line(l0 + min(max((n_after - 1) / 4 - qa.x, 0), sdc::kListCap - 1));
    const float* l1 = l0 + sdc::kListCap;
    prefetch_line(l1); prefetch_line(l1 + max(qm.y - 1, 0)); prefetch_line(l1 + min(max((3 * (n_after - 1)) / 4 - qa.y, 0), sdc::kListCap - 1));
    prefetch_line(S.agg_s + 4 * (size_t)env); prefetch_line(S.tail_thr + 4 * (size_t)env); prefetch_line(S.tail_bs + 4 * (size_t)env);
    // the band values at the split (where the fence sits inside each sorted band)
    const int2 nb = reinterpret_cast<const int2*>(S.tail_nb)[env];
    const float* band = S.tails + (size_t)env * 2 * sdc::kTailCap;
    prefetch_line(band + min(max(nb.x, 0), sdc::kTailCap - 1));
    prefetch_line(band + sdc::kTailCap + min(max(tn.y - nb.y, 0), sdc::kTailCap - 1));
}

struct GlobalInfoSink {
    float* info; int n, env;
    __device__ __forceinline__ void operator()(int col, float v) { if (info) info[(size_t)col * n + env] = v; }
};

// ---- window pass: the whole CTA processes one env's window, staged in shared memory by the TMA engine -------
// Parameters and results of the pass in flight (one at a time per CTA).
struct PassJob {
    // request (written by the lane that owns the env)
    int env, n, kind;
    int finish;                             // the step's reward waits for this pass: the CTA that runs it prices the step (finish_step)
    double lo64, hi64;                      // the step's fences in fp64: where the rebuilt bands are split
    float lo, hi, shift, tl, th, tl2, th2;
    int dir[2]; float thr[2];
    int rc[2], k[2]; float ca[2], cb[2];
    int tails, degenerate;
    int q_a[2], q_m[2];
    // results (written by warp 0)
    sdc::ScanResult rs;
    // finish == 1: what reward_finish needs besides the request and the results
    int m_ok; float alt3[3];
    double q1, m_c1, m_c2, m_c0, energy, nci_next, ls_penalty;
    unsigned long long t_pub;               // diagnostics (unit_log): when the record was published (globaltimer ns)
};
static_assert(sizeof(PassJob) <= sdc::kPassJobBytes && sdc::kPassJobBytes % 16 == 0, "maintenance pass record size");
struct PassShared {
    unsigned long long bar;                 // mbarrier of the bulk copy
    PassJob job;
    int fill[5];                            // slots handed out: lower band, upper band, collect list 0, collect list 1, (unused)
    int bkt_cnt[64], bkt_off[64], bkt_fill[64];   // bucket sort of the collected values: population, first slot, next free slot
    int band_pn[kWarpsPerBlock];            // rebuilt bands: per-warp partials of the values beyond the fence (count, sums about
    double band_p1[kWarpsPerBlock], band_p2[kWarpsPerBlock];   // the new centre); warps 0-3 lower band, 4-7 upper band
    float red_f[kWarpsPerBlock][4];         // s1, s2, ext0, ext1
    double red_d[kWarpsPerBlock][6];        // S1, S2, far sums
    int red_i[kWarpsPerBlock][6];           // cnt0, cnt1, below0, below1, far counts
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}
// One thread: global -> shared bulk copy by the TMA engine (UBLKCP), completion counted in bytes on the mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// The pass itself (sdc_core.h, "reward normaliser").  SCAN_PLAIN: clipped moments of this step only.  SCAN_REFRESH:
// additionally the exact unclipped moments about the new centre, the values of the two tail bands (straight into the
// env's band arrays) with the far-tail aggregates, all values inside the re-centring intervals of the brackets (into
// `scr`, then sorted) and the single-rank fallback -- then warp 0 commits the env's new incremental state.
// `win` = hist_cap floats of shared memory.  Called by all threads of the CTA; `phase` = parity of the mbarrier.
__device__ __noinline__ void window_pass(const sdc::State& S, PassShared& ps, float* win, float* scr, float* hits, int hit_cap, unsigned phase,
                                         uint32_t* clk_log = nullptr) {
    float* band_scr = scr + 2 * sdc::kCollectCap;                        // [2][kTailCap] band values, sorted by warps 0 / 1 below
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PassJob& J = ps.job;
    const int n = J.n, env = J.env;
    const bool refresh = J.kind == sdc::SCAN_REFRESH;
    if (tid == 0) {
        const unsigned bytes = ((unsigned)n * 4u + 15u) & ~15u;          // rows are 16-byte multiples (hist_cap % 4 == 0)
        // Generic-proxy accesses that the bulk copy (async proxy) must observe: this CTA's earlier use of `win` in shared
        // memory, and -- for a maintenance pass -- the window's newest sample, a generic store to GLOBAL memory by another CTA
        // of this launch that reached us through __threadfence + the job's ready tag.  fence.proxy.async covers both spaces.
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_expect_tx(&ps.bar, bytes);
        tma_load_1d(win, S.hist + (size_t)env * S.hist_cap, bytes, &ps.bar);
    }
    if (tid < 5) ps.fill[tid] = 0;
    const float lo = J.lo, hi = J.hi, shift = J.shift;
    const float tl = J.tl, th = J.th, tl2 = J.tl2, th2 = J.th2;
    const int dir0 = J.dir[0], dir1 = J.dir[1];
    const float thr0 = J.thr[0], thr1 = J.thr[1];
    const float ca0 = (refresh && J.rc[0]) ? J.ca[0] : SDC_INF_F, cb0 = (refresh && J.rc[0]) ? J.cb[0] : -SDC_INF_F;   // empty when not re-centring
    const float ca1 = (refresh && J.rc[1]) ? J.ca[1] : SDC_INF_F, cb1 = (refresh && J.rc[1]) ? J.cb[1] : -SDC_INF_F;
    const float quiet_lo = refresh ? fmaxf(tl, fmaxf(dir0 == sdc::SCAN_BELOW ? thr0 : -SDC_INF_F, dir1 == sdc::SCAN_BELOW ? thr1 : -SDC_INF_F)) : -SDC_INF_F;
    const float quiet_hi = refresh ? fminf(th, fminf(dir0 == sdc::SCAN_ABOVE ? thr0 : SDC_INF_F, dir1 == sdc::SCAN_ABOVE ? thr1 : SDC_INF_F)) : SDC_INF_F;
    float s1 = 0.f, s2 = 0.f;
    double S1 = 0.0, S2 = 0.0;
    const double c0 = (double)shift;
    int cnt0 = 0, cnt1 = 0, below0 = 0, below1 = 0, far_n0 = 0, far_n1 = 0;
    float ext0 = dir0 == sdc::SCAN_ABOVE ? SDC_INF_F : -SDC_INF_F, ext1 = dir1 == sdc::SCAN_ABOVE ? SDC_INF_F : -SDC_INF_F;
    double far_a0 = 0.0, far_b0 = 0.0, far_a1 = 0.0, far_b1 = 0.0;
    const long long tp0 = clock64();
    if (warp == 0) mbar_wait(&ps.bar, phase);                            // one warp polls; the others sleep at the barrier
    __syncthreads();                                                     // the window is in shared memory, fill[] zeroed
    const long long tp1 = clock64();
    // Per value: accumulate + one combined "is it interesting" test.  The rare hits (a few hundred of 10 000) are only
    // parked in a shared-memory list inside the loop -- with ~5 % hits nearly every warp iteration contains one, and a
    // divergent 80-instruction classification per iteration tripled the cost of the scan -- and classified densely
    // afterwards (shared-memory atomics hand out the slots of the bands / collections; the sets are unordered).
    auto classify = [&](float x) {
        const double y = (double)x - c0;
        if (dir0 | dir1) {
            if (dir0 == sdc::SCAN_BELOW && x < thr0) { cnt0 += 1; ext0 = fmaxf(ext0, x); }
            if (dir0 == sdc::SCAN_ABOVE && x > thr0) { cnt0 += 1; ext0 = fminf(ext0, x); }
            if (dir1 == sdc::SCAN_BELOW && x < thr1) { cnt1 += 1; ext1 = fmaxf(ext1, x); }
            if (dir1 == sdc::SCAN_ABOVE && x > thr1) { cnt1 += 1; ext1 = fminf(ext1, x); }
        }
        if (x < tl2) { far_n0 += 1; far_a0 += y; far_b0 = fma(y, y, far_b0); }
        else if (x < tl) { const int pos = atomicAdd(&ps.fill[0], 1); if (pos < sdc::kTailCap) band_scr[pos] = x; }
        if (x > th2) { far_n1 += 1; far_a1 += y; far_b1 = fma(y, y, far_b1); }
        else if (x > th) { const int pos = atomicAdd(&ps.fill[1], 1); if (pos < sdc::kTailCap) band_scr[sdc::kTailCap + pos] = x; }
        if (x >= ca0 && x <= cb0) { const int pos = atomicAdd(&ps.fill[2], 1); if (pos < sdc::kCollectCap) scr[pos] = x; }
        if (x >= ca1 && x <= cb1) { const int pos = atomicAdd(&ps.fill[3], 1); if (pos < sdc::kCollectCap) scr[sdc::kCollectCap + pos] = x; }
    };
    // Every warp parks its hits in its own slice of the list: the slot is the warp's running count plus the lane's rank among
    // the hitting lanes of this iteration (a ballot and a popcount) -- a shared counter bumped by an atomic per hit made ~80 %
    // of the iterations wait for a shared-memory atomic round trip and was 40 % of a pass.
    const int warp_cap = hit_cap / kWarpsPerBlock;
    float* my_hits = hits + warp * warp_cap;
    int n_mine = 0;                                                      // warp-uniform
    // four consecutive values per thread and trip (one 128-bit shared-memory load; the four values' arithmetic is independent,
    // which is what an in-order warp needs: the scan was bound by ~10 cycles per dependent instruction, not by issue slots)
    const int n_iter = (n + 4 * kStepThreads - 1) / (4 * kStepThreads);
#pragma unroll 2
    for (int it = 0; it < n_iter; ++it) {
        const int i0 = (tid + it * kStepThreads) * 4;
        float xs[4] = {shift, shift, shift, shift};
        if (i0 < n) { const float4 v = *reinterpret_cast<const float4*>(win + i0); xs[0] = v.x; xs[1] = v.y; xs[2] = v.z; xs[3] = v.w; }
        bool hit[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float x = xs[q];
            const bool in = i0 + q < n;
            if (in) {
                const float d = fminf(fmaxf(x, lo), hi) - shift;
                s1 += d; s2 = fmaf(d, d, s2);
                if (refresh) { const double y = (double)x - c0; S1 += y; S2 = fma(y, y, S2); }
                below0 += x < ca0; below1 += x < ca1;
            }
            hit[q] = in && (x < quiet_lo || x > quiet_hi || (x >= ca0 && x <= cb0) || (x >= ca1 && x <= cb1));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned hm = __ballot_sync(0xffffffffu, hit[q]);
            if (hm) {
                if (hit[q]) {
                    const int pos = n_mine + __popc(hm & ((1u << lane) - 1u));
                    if (pos < warp_cap) my_hits[pos] = xs[q];                // beyond the slice: classified by the second loop below
                }
                n_mine += __popc(hm);
            }
        }
    }
    // A warp with more hits than its slice holds (a window in the middle of a regime change: thousands of tail values) walks
    // its values once more and classifies the hits it could not park.  Kept out of the loop above on purpose: the
    // classification is ~100 instructions, and eight inlined copies of it made the scan loop a 19 KB instruction stream.
    if (n_mine > warp_cap) {
        int n_seen = 0;
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it) {
            const int i0 = (tid + it * kStepThreads) * 4;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                const bool in = i0 + q < n;
                const float x = in ? win[i0 + q] : shift;
                const bool hit = in && (x < quiet_lo || x > quiet_hi || (x >= ca0 && x <= cb0) || (x >= ca1 && x <= cb1));
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                if (hit && n_seen + __popc(hm & ((1u << lane) - 1u)) >= warp_cap) classify(x);
                n_seen += __popc(hm);
            }
        }
    }
    __syncwarp();
    const long long tp2 = clock64();
    {
        const int n_hits = min(n_mine, warp_cap);
#pragma unroll 1
        for (int i = lane; i < n_hits; i += 32) classify(my_hits[i]);
    }
    // ---- block reduction: warp shuffles, then the per-warp partials through shared memory ----
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (refresh) {
        S1 = warp_sum(S1); S2 = warp_sum(S2);
        cnt0 = warp_sum(cnt0); cnt1 = warp_sum(cnt1); below0 = warp_sum(below0); below1 = warp_sum(below1);
        far_n0 = warp_sum(far_n0); far_n1 = warp_sum(far_n1);
        far_a0 = warp_sum(far_a0); far_b0 = warp_sum(far_b0); far_a1 = warp_sum(far_a1); far_b1 = warp_sum(far_b1);
        ext0 = dir0 == sdc::SCAN_BELOW ? warp_max(ext0) : warp_min(ext0);
        ext1 = dir1 == sdc::SCAN_BELOW ? warp_max(ext1) : warp_min(ext1);
    }
    if (lane == 0) {
        ps.red_f[warp][0] = s1; ps.red_f[warp][1] = s2; ps.red_f[warp][2] = ext0; ps.red_f[warp][3] = ext1;
        ps.red_d[warp][0] = S1; ps.red_d[warp][1] = S2; ps.red_d[warp][2] = far_a0; ps.red_d[warp][3] = far_b0;
        ps.red_d[warp][4] = far_a1; ps.red_d[warp][5] = far_b1;
        ps.red_i[warp][0] = cnt0; ps.red_i[warp][1] = cnt1; ps.red_i[warp][2] = below0; ps.red_i[warp][3] = below1;
        ps.red_i[warp][4] = far_n0; ps.red_i[warp][5] = far_n1;
    }
    __syncthreads();                                                     // partials + all band / collect stores visible
    const long long tp3 = clock64();
    // Rebuilt bands: rank sort (thread i places value i of its band: the rank is the number of smaller values plus equal ones
    // before it; broadcast reads of shared memory, no barrier) into the free hit list, 128 threads per band, and the split
    // at the requesting step's fence: count and sums about the new centre of the values beyond it (per-warp partials, summed
    // in a fixed order by the committing warp: bit-reproducible whatever CTA runs the pass).
    float* band_sorted = hits;                                           // [2][kTailCap]; the parked hits were consumed above
    static_assert(kStepThreads == 2 * sdc::kTailCap, "one thread per band slot");
    if (refresh && J.tails) {
        const int sd = tid / sdc::kTailCap, me = tid - sd * sdc::kTailCap;
        const int cntb = ps.fill[sd];
        int nbz = 0; double z1 = 0.0, z2 = 0.0;
        if (ps.fill[0] <= sdc::kTailCap && ps.fill[1] <= sdc::kTailCap && me < cntb) {
            const float* b = band_scr + sd * sdc::kTailCap;
            const float x = b[me];
            int r = 0;
#pragma unroll 4
            for (int i = 0; i < cntb; ++i) { const float y = b[i]; r += (y < x) | ((y == x) & (i < me)); }
            band_sorted[sd * sdc::kTailCap + r] = x;
            if (sd == 0 ? (double)x < J.lo64 : (double)x > J.hi64) { const double y = (double)x - c0; nbz = 1; z1 = y; z2 = y * y; }
        }
        nbz = warp_sum(nbz); z1 = warp_sum(z1); z2 = warp_sum(z2);
        if (lane == 0) { ps.band_pn[warp] = nbz; ps.band_p1[warp] = z1; ps.band_p2[warp] = z2; }
    }
    sdc::RefreshRaw raw;
    sdc::ScanResult rs;
    {
        float f[4] = {0.f, 0.f, ps.red_f[0][2], ps.red_f[0][3]};
        double d[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        int c[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) {
            f[0] += ps.red_f[w][0]; f[1] += ps.red_f[w][1];
            f[2] = dir0 == sdc::SCAN_BELOW ? fmaxf(f[2], ps.red_f[w][2]) : fminf(f[2], ps.red_f[w][2]);
            f[3] = dir1 == sdc::SCAN_BELOW ? fmaxf(f[3], ps.red_f[w][3]) : fminf(f[3], ps.red_f[w][3]);
#pragma unroll
            for (int q = 0; q < 6; ++q) { d[q] += ps.red_d[w][q]; c[q] += ps.red_i[w][q]; }
        }
        rs.s1 = f[0]; rs.s2 = f[1];
        rs.ext[0] = dir0 == sdc::SCAN_NONE ? 0.f : f[2]; rs.ext[1] = dir1 == sdc::SCAN_NONE ? 0.f : f[3];
        rs.cnt[0] = c[0]; rs.cnt[1] = c[1]; rs.recentred = 0;
        rs.new_a[0] = J.q_a[0]; rs.new_a[1] = J.q_a[1]; rs.new_m[0] = J.q_m[0]; rs.new_m[1] = J.q_m[1];
        raw.s1 = d[0]; raw.s2 = d[1];
        raw.below[0] = c[2]; raw.below[1] = c[3];
        raw.agg_n[0] = c[4]; raw.agg_n[1] = c[5];
        raw.agg_s1[0] = d[2]; raw.agg_s2[0] = d[3]; raw.agg_s1[1] = d[4]; raw.agg_s2[1] = d[5];
        raw.n_tail[0] = ps.fill[0]; raw.n_tail[1] = ps.fill[1]; raw.c[0] = ps.fill[2]; raw.c[1] = ps.fill[3];
    }
    // (the band partials are read by warp 0 after the barrier that follows the bracket sorts)
    if (refresh) {
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {                                    // uniform across the CTA
            const int c = raw.c[j];
            if (c >= 1 && c <= sdc::kCollectCap) {
                // rank sort of the collected values (typically ~250): every thread places up to two of them; one pass over
                // the c values per element with broadcast reads and no barrier (a bitonic network needs 36-45 CTA barriers,
                // which is what a pass's latency was made of)
                // Bucket sort: 64 equal value buckets over [ca, cb] (the values are spread smoothly over this narrow interval),
                // population count, prefix, scatter into bucket order, then the rank inside the bucket by comparing with the
                // bucket's ~4-10 members only -- a plain rank sort compares with all c (c^2 / 256 per thread: a fifth of a pass).
                const float* buf = scr + j * sdc::kCollectCap;
                float* out = win + j * sdc::kCollectCap;              // the staged window is no longer needed
                float* grp = hits + 2 * sdc::kTailCap;                // [kCollectCap] the values grouped by bucket (after the sorted bands)
                const float ca = J.ca[j], cb = J.cb[j];
                const float scale = cb > ca ? 64.f / (cb - ca) : 0.f;
                if (tid < 64) ps.bkt_cnt[tid] = 0;
                __syncthreads();
                float myx[2] = {0.f, 0.f}; int myb[2] = {-1, -1};
                static_assert(sdc::kCollectCap <= 2 * kStepThreads, "two collected values per thread");
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int me = tid + q * kStepThreads;
                    if (me < c) {
                        const float x = buf[me];
                        myx[q] = x; myb[q] = min(63, max(0, (int)((x - ca) * scale)));
                        atomicAdd(&ps.bkt_cnt[myb[q]], 1);
                    }
                }
                __syncthreads();
                if (warp == 0) {                                      // exclusive prefix over the 64 populations, two per lane
                    const int c0b = ps.bkt_cnt[2 * lane], c1b = ps.bkt_cnt[2 * lane + 1];
                    int incl = c0b + c1b;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
                    const int excl = incl - (c0b + c1b);
                    ps.bkt_off[2 * lane] = excl; ps.bkt_off[2 * lane + 1] = excl + c0b;
                    ps.bkt_fill[2 * lane] = excl; ps.bkt_fill[2 * lane + 1] = excl + c0b;
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 2; ++q) if (myb[q] >= 0) grp[atomicAdd(&ps.bkt_fill[myb[q]], 1)] = myx[q];
                __syncthreads();
                for (int g = tid; g < c; g += kStepThreads) {
                    const float x = grp[g];
                    const int b = min(63, max(0, (int)((x - ca) * scale)));
                    const int b_lo = ps.bkt_off[b], b_hi = ps.bkt_fill[b];
                    int r = b_lo;
                    for (int i = b_lo; i < b_hi; ++i) { const float y = grp[i]; r += (y < x) | ((y == x) & (i < g)); }
                    out[r] = x;
                }
            }
        }
        __syncthreads();
        const long long tp4 = clock64();
        if (clk_log && tid == 0) { clk_log[0] = (uint32_t)(tp1 - tp0); clk_log[1] = (uint32_t)(tp2 - tp1); clk_log[2] = (uint32_t)(tp3 - tp2); clk_log[3] = (uint32_t)(tp4 - tp3); clk_log[5] = (uint32_t)raw.c[0] | ((uint32_t)raw.c[1] << 16); }
        if (warp == 0) {
            sdc::ScanRequest rl;
            rl.n = n; rl.kind = J.kind; rl.shift = shift; rl.tl = tl; rl.th = th; rl.tl2 = tl2; rl.th2 = th2;
            rl.rc[0] = J.rc[0]; rl.rc[1] = J.rc[1]; rl.k[0] = J.k[0]; rl.k[1] = J.k[1];
            rl.tails = J.tails; rl.degenerate = J.degenerate;
            sdc::QView Ql;
            Ql.lst[0] = S.qlist + (size_t)env * 2 * sdc::kListCap; Ql.lst[1] = Ql.lst[0] + sdc::kListCap;
            Ql.a[0] = J.q_a[0]; Ql.a[1] = J.q_a[1]; Ql.m[0] = J.q_m[0]; Ql.m[1] = J.q_m[1];
            const float* sorted[2] = {win, win + sdc::kCollectCap};
            const float* bands[2] = {band_sorted, band_sorted + sdc::kTailCap};
            for (int sd = 0; sd < 2; ++sd) {
                raw.band_nb[sd] = 0; raw.band_b1[sd] = raw.band_b2[sd] = 0.0;
                for (int w = 0; w < kWarpsPerBlock / 2; ++w) {
                    const int q = sd * (kWarpsPerBlock / 2) + w;
                    raw.band_nb[sd] += ps.band_pn[q]; raw.band_b1[sd] += ps.band_p1[q]; raw.band_b2[sd] += ps.band_p2[q];
                }
            }
            sdc::refresh_commit(S, env, rl, raw, sorted, bands, Ql, rs, lane, 32);
            // cursors of the re-centred brackets (a maintenance pass has no owner lane that would store them)
            if (lane < 2 && (rs.recentred & (1 << lane))) { S.q_a[env * 2 + lane] = rs.new_a[lane]; S.q_m[env * 2 + lane] = rs.new_m[lane]; }
        }
    }
    if (tid == 0) ps.job.rs = rs;
    __syncthreads();                                                     // results visible; `win`, `scr`, partials free again
    if (clk_log && tid == 0) { clk_log[7] = (uint32_t)(gtime_ns() - J.t_pub);      // publication -> end of the pass (ns)
                               clk_log[4] = (uint32_t)(clock64() - tp0); clk_log[6] = (uint32_t)J.kind | ((uint32_t)J.tails << 8) | ((uint32_t)J.rc[0] << 16) | ((uint32_t)J.rc[1] << 17); }
}

// The env of a finish job could not price its step without the window (a bracket ran out on one side, or the bands no
// longer contain the fences): the lane that owns it handed over the inputs of reward_finish, and the thread that holds the
// pass results does what that lane would have done after a synchronous pass -- bracket cursors, the three rewards, reward sums.
__device__ __noinline__ void finish_step(const sdc::State& S, const StepArgs& a, const PassJob& J) {
    const int env = J.env;
    sdc::ScanRequest rq;
    rq.kind = J.kind; rq.n = J.n; rq.dir[0] = J.dir[0]; rq.dir[1] = J.dir[1]; rq.degenerate = J.degenerate; rq.q1 = J.q1; rq.shift = J.shift;
    sdc::Moments M; M.c1 = J.m_c1; M.c2 = J.m_c2; M.c0 = J.m_c0; M.ok = J.m_ok;
    sdc::RewardInputs en; en.energy = J.energy; en.nci_next = J.nci_next; en.ls_penalty = J.ls_penalty;
    sdc::QView Q;
    Q.lst[0] = S.qlist + (size_t)env * 2 * sdc::kListCap; Q.lst[1] = Q.lst[0] + sdc::kListCap;
    Q.a[0] = J.q_a[0]; Q.a[1] = J.q_a[1]; Q.m[0] = J.q_m[0]; Q.m[1] = J.q_m[1];
    float r3[3];
    sdc::reward_finish(S, env, rq, J.rs, M, en, J.alt3, Q, r3);
    reinterpret_cast<int2*>(S.q_a)[env] = make_int2(Q.a[0], Q.a[1]);
    reinterpret_cast<int2*>(S.q_m)[env] = make_int2(Q.m[0], Q.m[1]);
    a.rew[env * 3 + 0] = r3[0]; a.rew[env * 3 + 1] = r3[1]; a.rew[env * 3 + 2] = r3[2];
    atomicAdd(a.metrics + sdc::M_REWARD_SUM, (double)r3[0] + r3[1] + r3[2]);
    atomicAdd(a.metrics + sdc::M_REWARD_LS, (double)r3[0]); atomicAdd(a.metrics + sdc::M_REWARD_DC, (double)r3[1]);
}

// =================================================================================================
// episode reset / generation of one env by one CTA (k_reset, and the worker jobs inside k_step)
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

struct OutPtrs { float* obs; float* share; float* obs_c; };       // any of them may be null
// One env's observation rows from its zero-padded [3][26] row: padded obs, HARL shared row, compact row.
__device__ __forceinline__ void store_env_rows(const float* row, int env, const OutPtrs& o, int t, int n_thr) {
    if (o.obs) for (int k = t; k < kObsRow; k += n_thr) o.obs[(size_t)env * kObsRow + k] = row[k];
    if (o.share) for (int k = t; k < SDC_SHARE_DIM; k += n_thr) {
        const int src = k < 26 ? k : (k == 26 ? SDC_OBS_DIM + 11 : (k == 27 ? SDC_OBS_DIM + 13 : 2 * SDC_OBS_DIM + 25));
        o.share[(size_t)env * SDC_SHARE_DIM + k] = row[src];
    }
    if (o.obs_c) for (int k = t; k < SDC_OBS_COMPACT; k += n_thr) o.obs_c[(size_t)env * SDC_OBS_COMPACT + k] = row[sdc::compact_to_padded(k)];
}

struct RowSink {
    float* row;
    __device__ __forceinline__ void operator()(int agent, int idx, float v) { row[agent * SDC_OBS_DIM + idx] = v; }
};
// Keeps only the 29 distinct values of the three rows (agent_ls[26] | agent_dc[11] | agent_dc[13] | agent_bat[12], see
// sdc::compact_to_padded): every other dc / bat entry repeats an agent_ls entry or is padding.  The call sites pass constant
// (agent, idx), so the dropped stores vanish at compile time.
struct CompactSink {
    float* row;
    __device__ __forceinline__ void operator()(int agent, int idx, float v) {
        if (agent == 0) row[idx] = v;
        else if (agent == 1 && idx == 11) row[26] = v;
        else if (agent == 1 && idx == 13) row[27] = v;
        else if (agent == 2 && idx == 12) row[28] = v;
    }
};
// column of the zero-padded [3][26] row -> compact column, -1 = padding zero (inverse of sdc::compact_to_padded plus the repeats)
__host__ __device__ constexpr int padded_to_compact(int col) {
    return col < 26 ? col
         : col < 52 ? (col - 26 < 10 ? col - 26 : (col - 26 == 10 ? 13 : (col - 26 == 11 ? 26 : (col - 26 == 12 ? 14 : (col - 26 == 13 ? 27 : -1)))))
                    : (col - 52 < 10 ? col - 52 : (col - 52 == 10 ? 13 : (col - 52 == 11 ? 14 : (col - 52 == 12 ? 28 : -1))));
}
// One env's rows from its COMPACT row (shared memory): padded obs, HARL shared row, compact row.  p2c: padded_to_compact as a table.
__device__ __forceinline__ void store_env_rows_compact(const float* crow, int env, const OutPtrs& o, int t, int n_thr, const signed char* p2c) {
    if (o.obs) for (int k = t; k < kObsRow; k += n_thr) { const int c = p2c[k]; o.obs[(size_t)env * kObsRow + k] = c >= 0 ? crow[c] : 0.f; }
    if (o.share) for (int k = t; k < SDC_SHARE_DIM; k += n_thr) o.share[(size_t)env * SDC_SHARE_DIM + k] = k < 28 ? crow[k] : 0.f;
    if (o.obs_c) for (int k = t; k < SDC_OBS_COMPACT; k += n_thr) o.obs_c[(size_t)env * SDC_OBS_COMPACT + k] = crow[k];
}

struct ResetShared {
    double red[kResetThreads / 32];
    double seg_off[sdc::kNoiseSegs];
    int start[3];
    float row[kObsRow];
    sdc::CiFeat ci;                 // the two halves of a staged reset observation (pregen_one_env)
    sdc::TempFeat tf;
};

constexpr int kNormWindow = 2880;                 // 30 days of quarter-hours (utils/managers.py:435,606)

// Episode reset of one env by one CTA.  `runbuf` = max(kNormWindow, win_len) doubles of shared memory (run_buf_doubles).
// Weather noise (utils/managers.py:35-48,596-613): the year-long random walk is generated ONCE (Philox, 140 samples per
// thread); its mean / variance come from per-thread partial sums combined with the segment offsets, and only the walk
// values that land in the episode window or in the 30-day normalisation slice after the start are kept (in shared
// memory).  The emit pass then runs over window positions, so its trace reads are coalesced and independent.  No global
// scratch: a dependent global access costs ~2 us while the other CTAs saturate HBM with window scans.
// Generates the next episode of `env` (start day / hour, realised weather window, 30-day temperature range) from the
// env's counter-based RNG stream into (wt, ww, *tmin_out, *tmax_out, sh.start).  All threads of the CTA.
__device__ __forceinline__ void generate_episode(const sdc::State& S, int env, double* wt, double* ww, double* tmin_out, double* tmax_out,
                                                 double* runbuf, ResetShared& sh) {
    const int tid = threadIdx.x;
    const int n = SDC_YEAR_STEPS;
    const uint32_t ep = S.episode[env];
    const uint64_t seed = S.seed[env];
    __syncthreads();
    if (tid == 0) sdc::draw_episode_start(seed, ep, S.day_lo[env], S.day_hi[env], &sh.start[0], &sh.start[1], &sh.start[2]);
    __syncthreads();
    const int t0 = sh.start[0] * 96 + sh.start[1] * 4, roll = sh.start[2];
    const int k_norm = min(kNormWindow, n - t0);           // the reference's 30-day slice is truncated at the year end
    const int k_win = min(S.win_len, n - t0);              // episode window (a 30-day episode is 2898 samples: longer than the slice)
    const int k_keep = max(k_norm, k_win);
    // pass 1: this thread's segment of the walk (utils/managers.py:45-46), normals from the segment's own PCG32 stream.
    // Walk sample j lands at trace index (j + 96 roll) mod n, i.e. at window position k = j + c_lo (c_hi past the wrap).
    // (Two interleaved half-segments per thread were measured: 7 % slower -- registers, not latency, bound this loop.)
    double run = 0.0, sum_run = 0.0, sum_run2 = 0.0;
    const int j0 = tid * sdc::kNoiseSeg;
    const int cnt = max(0, min(sdc::kNoiseSeg, n - j0));    // even for every thread (n and the segment length are even)
    const int jw = n - 96 * roll, c_lo = 96 * roll - t0, c_hi = c_lo - n;
    sdc::Pcg32 g = sdc::noise_stream(seed, ep, (uint32_t)tid);
    // Only ~1 thread in 12 owns a segment with samples inside the kept window: the others run the walk without the position
    // arithmetic and the conditional store (a third of the loop's instructions).
    bool keeps = false;
    {
        const int lo_a = j0, hi_a = min(j0 + cnt, jw);                 // samples before the wrap: k = j + c_lo
        const int lo_b = max(j0, jw), hi_b = j0 + cnt;                 // samples after it: k = j + c_hi
        if (lo_a < hi_a && lo_a + c_lo < k_keep && hi_a - 1 + c_lo >= 0) keeps = true;
        if (lo_b < hi_b && lo_b + c_hi < k_keep && hi_b - 1 + c_hi >= 0) keeps = true;
        keeps = __any_sync(0xffffffffu, keeps);                        // per warp: a warp that ran both loops would take twice as long
    }
    // The kept samples of a thread are one contiguous range of its segment (the window is contiguous in trace index, and the
    // two pieces a wrapped window has in walk index lie at opposite ends of the walk): sample q goes to runbuf[q + k_off] for
    // q in [q_lo, q_hi).  Should both mappings ever apply to one segment, the warp takes the per-sample form.
    int q_lo = 0, q_hi = 0, k_off = 0;
    bool generic = false;
    {
        const int a_lo = max(0, -c_lo - j0), a_hi = min(cnt, min(jw, k_keep - c_lo) - j0);
        const int b_lo = max(0, max(jw, -c_hi) - j0), b_hi = min(cnt, k_keep - c_hi - j0);
        if (a_lo < a_hi && b_lo < b_hi) generic = true;
        else if (a_lo < a_hi) { q_lo = a_lo; q_hi = a_hi; k_off = j0 + c_lo; }
        else if (b_lo < b_hi) { q_lo = b_lo; q_hi = b_hi; k_off = j0 + c_hi; }
        generic = __any_sync(0xffffffffu, generic);
    }
    if (keeps && !generic) {
#pragma unroll 2
        for (int q = 0; q < cnt; q += 2) {
            float z[2];
            sdc::noise_pair(g, z);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                run += (double)(0.02f * z[u]);
                sum_run += run; sum_run2 = fma(run, run, sum_run2);
                if (q + u >= q_lo && q + u < q_hi) runbuf[q + u + k_off] = run;
            }
        }
    } else if (keeps) {
#pragma unroll 2
        for (int q = 0; q < cnt; q += 2) {
            float z[2];
            sdc::noise_pair(g, z);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = j0 + q + u;
                run += (double)(0.02f * z[u]);
                sum_run += run; sum_run2 = fma(run, run, sum_run2);
                const int k = j + (j >= jw ? c_hi : c_lo);
                if ((unsigned)k < (unsigned)k_keep) runbuf[k] = run;
            }
        }
    } else {
#pragma unroll 2
        for (int q = 0; q < cnt; q += 2) {
            float z[2];
            sdc::noise_pair(g, z);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                run += (double)(0.02f * z[u]);
                sum_run += run; sum_run2 = fma(run, run, sum_run2);
            }
        }
    }
    sh.seg_off[tid] = run;
    __syncthreads();
    if (tid < 32) {                           // exclusive prefix over the segment totals: 8 per lane + a warp scan
        constexpr int kPer = sdc::kNoiseSegs / 32;
        double v[kPer], tot = 0.0;
#pragma unroll
        for (int i = 0; i < kPer; ++i) { v[i] = sh.seg_off[tid * kPer + i]; tot += v[i]; }
        double incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double up = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += up; }
        double acc = incl - tot;
#pragma unroll
        for (int i = 0; i < kPer; ++i) { sh.seg_off[tid * kPer + i] = acc; acc += v[i]; }
    }
    __syncthreads();
    // walk_j = off + run_j  ->  sums of w and w^2 from the partial sums
    const double off = sh.seg_off[tid];
    const double pw = (double)cnt * off + sum_run;
    const double pw2 = (double)cnt * off * off + 2.0 * off * sum_run + sum_run2;
    const double sw = block_sum(pw, sh.red);
    const double sw2 = block_sum(pw2, sh.red);
    const double mean = sw / n;
    const double scale = 0.75 / sqrt(sw2 / n - mean * mean);              // managers.py:46-48
    // emit: roll, clip, window, 30-day min/max (managers.py:598-608), one window position per thread and trip
    const sdc::LocTables& L = S.loc[S.loc_id[env]];
    double tmin = INFINITY, tmax = -INFINITY;
    // four positions per thread and trip, all trace loads of a trip in flight together (the loop was a chain of ~12 L2 round
    // trips per thread)
    constexpr int kEmitUnroll = 4;
    const int k_max = max(k_keep, S.win_len);
    for (int kb = tid; kb < k_max; kb += kEmitUnroll * kResetThreads) {
        double tb[kEmitUnroll], wb[kEmitUnroll], nz[kEmitUnroll];
#pragma unroll
        for (int u = 0; u < kEmitUnroll; ++u) {
            const int k = kb + u * kResetThreads;
            tb[u] = wb[u] = nz[u] = 0.0;
            if (k < k_keep) {
                int j = t0 + k - 96 * roll; if (j < 0) j += n;
                tb[u] = L.temp_base[j];
                if (k < k_win) wb[u] = L.wetb_base[j];
                nz[u] = (sh.seg_off[j / sdc::kNoiseSeg] + runbuf[k]) * scale;     // segment offset + position inside it
            }
        }
#pragma unroll
        for (int u = 0; u < kEmitUnroll; ++u) {
            const int k = kb + u * kResetThreads;
            if (k < k_keep) {
                const double vt = fmin(fmax(tb[u] + nz[u], 0.0), 45.0);
                if (k < k_norm) { tmin = fmin(tmin, vt); tmax = fmax(tmax, vt); }
                if (k < k_win) { wt[k] = vt; ww[k] = fmin(fmax(wb[u] + nz[u], 0.0), 45.0); }
            }
            if (k >= k_win && k < S.win_len) { wt[k] = 0.0; ww[k] = 0.0; }    // beyond the year end (flagged domain) / padding
        }
    }
    tmin = block_minmax(tmin, true, sh.red);
    tmax = block_minmax(tmax, false, sh.red);
    if (tid == 0) { *tmin_out = tmin; *tmax_out = tmax; }
}

// Look-ahead: an env that will finish two steps from now gets its next episode generated into its staging buffers (the
// other weather buffer; the ones sdc_stage_episode fills in replay mode) TOGETHER WITH the observation its reset will
// return, while the other CTAs step -- so that the reset itself is a buffer flip and a 428-byte copy that the env's own
// warp does at the end of its step (k_step), not a job.
__device__ __forceinline__ void pregen_one_env(const sdc::State& S, int env, double* runbuf, ResetShared& sh) {
    if (S.pend_valid[env]) return;                      // uniform: a host-staged (or already generated) episode is waiting
    double* wt = sdc::weather_pend(S, env);
    generate_episode(S, env, wt, wt + S.win_len, S.pend_tmin + env, S.pend_tmax + env, runbuf, sh);
    __syncthreads();                                    // window, range and start are in place (global / shared memory)
    // the reset observation (sustaindc_env.py:488-494): its two halves on two warps, then one thread writes the rows
    const int t0 = sh.start[0] * 96 + sh.start[1] * 4;
    const sdc::LocTables& L = S.loc[S.loc_id[env]];
    if (threadIdx.x == 0) {
        const double cmin = L.ci_min30[t0];
        sdc::ci_features(L, t0, cmin, L.ci_max30[t0] - cmin, sh.ci);
    } else if (threadIdx.x == 32) {
        const double tmin = S.pend_tmin[env];
        sdc::temp_features(wt, tmin, S.pend_tmax[env] - tmin, sh.tf);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        S.pend_day[env] = sh.start[0]; S.pend_hour[env] = sh.start[1];
        sdc::LsStats ls;
        ls.oldest = 0.0; ls.avg = 0.0; ls.norm_q = 0.0;
        for (int i = 0; i < 5; ++i) ls.hist[i] = 0.0;
        RowSink sink{S.pend_obs + (size_t)env * kObsRow};
        sdc::emit_obs_rows(S.hour_cos[t0 % 96], S.hour_sin[t0 % 96], L.workload[t0], L.workload[t0 + 1], sh.ci, sh.tf, ls, 0.0, sink);
        __threadfence();
        S.pend_valid[env] = 3;
    }
}

// Episode reset of one env by one CTA (k_reset, and the worker jobs of k_step for envs whose next episode was not staged
// with its observation: first resets, host-staged replays).  `runbuf` = run_buf_bytes(S) of shared memory.
__device__ __forceinline__ void reset_one_env(const sdc::State& S, int env, const OutPtrs& out, double* runbuf, ResetShared& sh) {
    const int tid = threadIdx.x;
    const int staged = S.pend_valid[env];
    double* wt = sdc::weather_pend(S, env);             // the next episode's window: staged, or generated right here
    __syncthreads();
    if (staged & 1) {
        if (tid == 0) { sh.start[0] = S.pend_day[env]; sh.start[1] = S.pend_hour[env]; sh.start[2] = 0; }
    } else {
        generate_episode(S, env, wt, wt + S.win_len, S.pend_tmin + env, S.pend_tmax + env, runbuf, sh);
    }
    uint32_t* ring = reinterpret_cast<uint32_t*>(S.ls_ring + (size_t)env * (S.ls_mask + 1));
    for (int k = tid; k < (S.ls_mask + 1) / 4; k += kResetThreads) ring[k] = 0u;
    __syncthreads();                               // weather window + range visible to thread 0
    if (tid == 0) {
        const int t0 = sh.start[0] * 96 + sh.start[1] * 4;
        const double tmin = S.pend_tmin[env], tmax = S.pend_tmax[env];
        S.t_min[env] = tmin; S.t_max[env] = tmax;
        S.cur_buf[env] ^= 1;                       // the staged window becomes the current one
        S.pend_valid[env] = 0;
        S.episode[env] += 1;
        sdc::reset_scalars(S, env, t0);
        if (staged & 2) {
            for (int k = 0; k < kObsRow; ++k) sh.row[k] = S.pend_obs[(size_t)env * kObsRow + k];
        } else {
            RowSink sink{sh.row};
            sdc::reset_observation(S, env, t0, wt, tmin, tmax, sink);
        }
    }
    __syncthreads();
    store_env_rows(sh.row, env, out, tid, kResetThreads);
}

// =================================================================================================
// k_reset: explicit resets (sdc_reset), one CTA per listed env
// =================================================================================================
__global__ void __launch_bounds__(kResetThreads) k_reset(const __grid_constant__ sdc::State S, const int32_t* __restrict__ list,
                                                         const int32_t* __restrict__ count, float* obs, float* share) {
    extern __shared__ double runbuf[];              // [run_buf_doubles] walk values inside the episode window / 30-day slice
    __shared__ ResetShared sh;
    const int total = *count;
    const OutPtrs out{obs, share, nullptr};
    for (int i = blockIdx.x; i < total; i += gridDim.x) reset_one_env(S, list[i], out, runbuf, sh);
}

// =================================================================================================
// k_step
// =================================================================================================
__global__ void __launch_bounds__(kStepThreads, 2) k_step(const __grid_constant__ sdc::State S, const __grid_constant__ StepArgs a, const int n_unit_ctas, const int hit_cap,
                                                          const int table_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int U = a.unit_envs;
    // location / dc-parameter tables -> shared memory (removes one level of pointer chasing per env)
    sdc::Tables T{S.loc, S.dc};
    __shared__ PassShared ps;
    __shared__ signed char p2c[kObsRow + 2];               // padded column -> compact column (-1: padding zero)
    if (threadIdx.x < kObsRow) p2c[threadIdx.x] = (signed char)padded_to_compact(threadIdx.x);
    // The first unit of this warp is known here: its normaliser-side reads (which need no table) are put in flight before
    // the CTA waits for the table copy below.
    bool early_prefetch = false;
    if (blockIdx.x < n_unit_ctas) {
        const int unit = blockIdx.x * kWarpsPerBlock + warp;
        const int env = unit * U + lane;
        if (unit * U < S.n_envs) { prefetch_env(S, T, env, lane < U && env < S.n_envs, 2); early_prefetch = true; }
    }
    {
        const int loc_bytes = S.n_loc * (int)sizeof(sdc::LocTables), dc_bytes = S.n_cfg * (int)sizeof(sdc_dc_params);
        if (loc_bytes + dc_bytes <= table_bytes) {
            int* dst = reinterpret_cast<int*>(smem_raw);
            const int* src_loc = reinterpret_cast<const int*>(S.loc);
            const int* src_dc = reinterpret_cast<const int*>(S.dc);
            for (int i = threadIdx.x; i < loc_bytes / 4; i += kStepThreads) dst[i] = src_loc[i];
            for (int i = threadIdx.x; i < dc_bytes / 4; i += kStepThreads) dst[loc_bytes / 4 + i] = src_dc[i];
            T.loc = reinterpret_cast<const sdc::LocTables*>(smem_raw);
            T.dc = reinterpret_cast<const sdc_dc_params*>(smem_raw + loc_bytes);
        }
        if (threadIdx.x == 0) {
            mbar_init(&ps.bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    float* scr = reinterpret_cast<float*>(smem_raw + table_bytes);      // [2][kCollectCap] + [2][kTailCap] collect scratch of the window pass
    float* win = scr + 2 * sdc::kCollectCap + 2 * sdc::kTailCap;        // [hist_cap] the staged window
    const int win_floats = S.hist_cap > 2 * sdc::kCollectCap ? S.hist_cap : 2 * sdc::kCollectCap;
    float* hits = win + win_floats;                                     // parked hits of a pass: the rest of the region the obs tiles use

    unsigned pass_phase = 0;
    const int N = S.n_envs;
    const int n_units = (N + U - 1) / U;
    if (blockIdx.x == 0 && threadIdx.x < 16) a.ctr_next[threadIdx.x] = 0;

    // ------------------------------------------------------------------------------------------------------------
    // Units.  A warp takes a unit of U consecutive envs, one lane per env:
    //   physics (load shifting, data centre, battery, traces) -> the step's energy
    //   reward normaliser, incremental: window append, quartile brackets, moments, tail bands (sdc_core.h)
    //   rewards, observations, logger sums, resets of finished envs.
    // An env whose incremental state ran out of slack publishes a window pass to the job queue (served by the worker
    // CTAs: TMA bulk copy of the window into shared memory, all 256 threads scan, sort and commit); the rare env that
    // cannot price this step without its window leaves the pricing to the CTA that runs the pass (finish_step).
    // ------------------------------------------------------------------------------------------------------------
    if (a.phase_clocks && threadIdx.x == 0) atomicMin(a.phase_clocks + 13, gtime_ns());
    // Warps run their units independently: nothing inside a unit meets the rest of the CTA.  The first unit of a warp is
    // its slot in the grid (no round trip to the ticket counter at the head of the launch); further ones -- batches larger
    // than one resident round -- come from the counter.
    const int n_first = min(n_unit_ctas * kWarpsPerBlock, n_units);
    bool first_round = true;
    if (blockIdx.x < n_unit_ctas) for (;;) {
        int unit = blockIdx.x * kWarpsPerBlock + warp;
        if (!first_round) {
            if (lane == 0) unit = n_first + atomicAdd(a.ctr + 0, 1);
            unit = __shfl_sync(0xffffffffu, unit, 0);
        }
        first_round = false;
        if (unit >= n_units) break;                                     // warp-uniform
        constexpr bool have_unit = true;
        const int env0 = unit * U;
        const int env = env0 + lane;
        const bool active = have_unit && lane < U && env < N;
        const long long tk0 = clock64();
        const unsigned long long t_unit0 = a.unit_log ? gtime_ns() : 0ull;
        unsigned n_edits = 0;                           // diagnostics: bracket edits | band edits << 8 | passes asked << 16
        sdc::StepResult st;
        sdc::ObsDeferred od;
        sdc::RewardInputs en;
        sdc::ScanRequest rq;
        sdc::ScanResult rs;
        sdc::Moments M;
        sdc::QView Q;
        sdc::PrepState prep; prep.search[0] = prep.search[1] = 0; prep.err = 0; prep.fc = 0u;
        prep.first[0] = prep.first[1] = prep.last[0] = prep.last[1] = 0.f;
        sdc::ListEdit edits[2];
        edits[0].rm = edits[1].rm = -1; edits[0].drop = edits[1].drop = 0; edits[0].ins = edits[1].ins = -1;
        edits[0].val = edits[1].val = 0.f; edits[0].m0 = edits[1].m0 = 0;
        st.terminal = 0;
        en.energy = 0.0; en.nci_next = 0.0; en.ls_penalty = 0.0;
        rq.kind = sdc::SCAN_SKIP; rq.n = 0; rq.lo = rq.hi = rq.shift = 0.f; rq.tl = rq.th = rq.tl2 = rq.th2 = 0.f; rq.tails = 0;
        rq.lo64 = rq.hi64 = rq.q1 = 0.0; rq.e = rq.o = 0.f; rq.evict = 0;
        rq.dir[0] = rq.dir[1] = 0; rq.thr[0] = rq.thr[1] = 0.f; rq.rc[0] = rq.rc[1] = 0; rq.k[0] = rq.k[1] = 0;
        rq.ca[0] = rq.ca[1] = rq.cb[0] = rq.cb[1] = 0.f; rq.degenerate = 0;
        rs.s1 = rs.s2 = 0.f; rs.cnt[0] = rs.cnt[1] = 0; rs.ext[0] = rs.ext[1] = 0.f; rs.recentred = 0;
        rs.new_a[0] = rs.new_a[1] = rs.new_m[0] = rs.new_m[1] = 0;
        M.ok = 0; M.c0 = M.c1 = M.c2 = 0.0;
        Q.lst[0] = S.qlist + (size_t)(active ? env : 0) * 2 * sdc::kListCap; Q.lst[1] = Q.lst[0] + sdc::kListCap;
        Q.a[0] = Q.a[1] = Q.m[0] = Q.m[1] = 0;
        long long tk1 = tk0;
        float alt3[3] = {0.f, 0.f, 0.f};
        sdc::BandPlan bp; bp.rm[0] = bp.rm[1] = bp.ins[0] = bp.ins[1] = 0;
        sdc::BandDone bd; bd.rm[0] = bd.rm[1] = bd.pos[0] = bd.pos[1] = -1;
        // the unit's 96 action ids as three contiguous 128-byte rows (they may sit in pinned host memory: a host call without
        // an H2D copy), handed to their lanes by shuffles
        int a_ls = 0, a_dc = 0, a_bat = 0;
        {
            int raw[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { const size_t i = (size_t)env0 * 3 + lane + 32 * k; raw[k] = i < (size_t)N * 3 ? a.actions[i] : 0; }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int e = 3 * lane + c, from = e & 31, k = e >> 5;
                const int v0 = __shfl_sync(0xffffffffu, raw[0], from), v1 = __shfl_sync(0xffffffffu, raw[1], from), v2 = __shfl_sync(0xffffffffu, raw[2], from);
                const int v = k == 0 ? v0 : (k == 1 ? v1 : v2);
                if (c == 0) a_ls = v; else if (c == 1) a_dc = v; else a_bat = v;
            }
        }
        prefetch_env(S, T, env, active, early_prefetch ? 1 : 3);
        early_prefetch = false;
        if (active) {
            const int2 qa = reinterpret_cast<const int2*>(S.q_a)[env], qm = reinterpret_cast<const int2*>(S.q_m)[env];
            Q.a[0] = qa.x; Q.a[1] = qa.y; Q.m[0] = qm.x; Q.m[1] = qm.y;
            if (S.append_history) sdc::load_list_ends(Q, prep);     // in flight during the physics; nothing writes the lists before reward_prepare_c
            {
                GlobalInfoSink info{a.info, N, env};
                sdc::physics_step(S, T, env, a_ls, a_dc, a_bat, info, st, od, false);
                st.evicted = S.hist[(size_t)env * S.hist_cap + st.hist_head];     // prefetched into L2 at the start of the unit
            }
            en.energy = st.energy; en.nci_next = st.nci_next; en.ls_penalty = st.ls_penalty;
            if (sdc::any_alt_reward(S)) {
                const sdc::AltInputs ai{st.ite_kw, st.total_kw, st.water, od.tn % 96};
                sdc::alt_rewards(S, env, st.energy, ai, alt3);
            }
            tk1 = clock64();
            if (S.append_history)                      // utils/reward_creator.py:62-63: only default_ls_reward grows the window
                sdc::reward_prepare_a(S, env, en.energy, st.hist_len, st.hist_head, st.evicted, Q, rq, prep, true);   // en.energy -> relative
        }
        __syncwarp();
        // A sample that leaves / enters INSIDE a bracket (~2.4 % per env and list) needs its index / position in the sorted
        // list: for the env's lane alone that is a linear search in batches plus a binary search, ~25 dependent round
        // trips (7.5 k clocks per edit while 31 lanes wait).  The whole warp does it instead: one coalesced load of the
        // list (four values per lane), a ballot for the evicted value, a ballot count of the values <= the new one.
        sdc::ListHints hints; hints.rm[0] = hints.rm[1] = hints.pos[0] = hints.pos[1] = sdc::kNoHint;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            unsigned need = __ballot_sync(0xffffffffu, active && S.append_history && prep.search[j] != 0);
            while (need) {
                const int l = __ffs(need) - 1;
                need &= need - 1;
                const int want_rm = __shfl_sync(0xffffffffu, prep.search[j], l) & 1;
                const float o_l = __shfl_sync(0xffffffffu, rq.o, l), e_l = __shfl_sync(0xffffffffu, rq.e, l);
                const int m0 = __shfl_sync(0xffffffffu, Q.m[j], l);
                const float* B = S.qlist + ((size_t)(env0 + l) * 2 + j) * sdc::kListCap;
                float v[sdc::kListCap / 32];
#pragma unroll
                for (int t = 0; t < sdc::kListCap / 32; ++t) { const int i = lane + 32 * t; v[t] = i < m0 ? B[i] : SDC_INF_F; }
                int rm = -1;
                if (want_rm) {
                    rm = m0;                           // not found: list_remove flags the bracket
#pragma unroll
                    for (int t = sdc::kListCap / 32 - 1; t >= 0; --t) {
                        const unsigned hit = __ballot_sync(0xffffffffu, lane + 32 * t < m0 && v[t] == o_l);
                        if (hit) rm = 32 * t + __ffs(hit) - 1;
                    }
                }
                int pos = 0;
#pragma unroll
                for (int t = 0; t < sdc::kListCap / 32; ++t) {
                    const int i = lane + 32 * t;
                    pos += __popc(__ballot_sync(0xffffffffu, i < m0 && i != rm && v[t] <= e_l));
                }
                if (lane == l) { hints.rm[j] = want_rm ? rm : sdc::kNoHint; hints.pos[j] = pos; }
            }
        }
        if (active && S.append_history) {
            sdc::reward_prepare_c(S, env, Q, rq, edits, prep, hints);
            sdc::reward_plan_a(S, env, rq, M, bp);
        }
        __syncwarp();
        // Planned bracket edits (a value entered / left inside a bracket: ~10 % of the env-steps) are applied by the whole
        // warp, one env at a time: every lane computes its elements of the edited list from the old one, then stores them.
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            unsigned need = __ballot_sync(0xffffffffu, active && !sdc::edit_trivial(edits[j]));
            n_edits += __popc(need);
            while (need) {
                const int l = __ffs(need) - 1;
                need &= need - 1;
                sdc::ListEdit ed;
                ed.rm = __shfl_sync(0xffffffffu, edits[j].rm, l); ed.drop = __shfl_sync(0xffffffffu, edits[j].drop, l);
                ed.ins = __shfl_sync(0xffffffffu, edits[j].ins, l); ed.val = __shfl_sync(0xffffffffu, edits[j].val, l);
                ed.m0 = __shfl_sync(0xffffffffu, edits[j].m0, l);
                float* B = S.qlist + ((size_t)(env0 + l) * 2 + j) * sdc::kListCap;
                const int len = sdc::edit_len(ed);
                float v[sdc::kListCap / 32];
#pragma unroll
                for (int t = 0; t < sdc::kListCap / 32; ++t) { const int i = lane + 32 * t; v[t] = i < len ? sdc::edit_at(B, ed, i) : 0.f; }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < sdc::kListCap / 32; ++t) { const int i = lane + 32 * t; if (i < len) B[i] = v[t]; }
                __syncwarp();
            }
        }
        // The step's sample entering / the evicted one leaving a tail band (~2 % of the env-steps): sorted removal / insertion
        // by the whole warp -- every lane holds four band slots, ballots find the evicted value and count the values <= the
        // new one (insertion after ties), then the edited band is written back like a bracket.
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            unsigned need = __ballot_sync(0xffffffffu, active && (bp.rm[sd] | bp.ins[sd]));
            n_edits += __popc(need) << 8;
            while (need) {
                const int l = __ffs(need) - 1;
                need &= need - 1;
                const int do_rm = __shfl_sync(0xffffffffu, bp.rm[sd], l), do_ins = __shfl_sync(0xffffffffu, bp.ins[sd], l);
                const float o_l = __shfl_sync(0xffffffffu, rq.o, l), e_l = __shfl_sync(0xffffffffu, rq.e, l);
                float* B = sdc::tail_ptr(S, env0 + l, sd);
                const int nB = S.tail_n[2 * (env0 + l) + sd];              // same address for the warp: one broadcast load
                float v[sdc::kTailCap / 32];
#pragma unroll
                for (int t = 0; t < sdc::kTailCap / 32; ++t) { const int i = lane + 32 * t; v[t] = i < nB ? B[i] : SDC_INF_F; }
                int rm = -1, pos = -1;
                if (do_rm) {
#pragma unroll
                    for (int t = 0; t < sdc::kTailCap / 32; ++t) {
                        const unsigned hit = __ballot_sync(0xffffffffu, lane + 32 * t < nB && v[t] == o_l);
                        if (hit && rm < 0) rm = 32 * t + __ffs(hit) - 1;
                    }
                }
                if (do_ins && nB - (rm >= 0) < sdc::kTailCap) {
                    pos = 0;
#pragma unroll
                    for (int t = 0; t < sdc::kTailCap / 32; ++t) {
                        const int i = lane + 32 * t;
                        pos += __popc(__ballot_sync(0xffffffffu, i < nB && i != rm && v[t] <= e_l));
                    }
                }
                if (rm >= 0 || pos >= 0) {
                    sdc::ListEdit ed; ed.rm = rm; ed.drop = 0; ed.ins = pos; ed.val = e_l; ed.m0 = nB;
                    const int len = sdc::edit_len(ed);
                    float w[sdc::kTailCap / 32];
#pragma unroll
                    for (int t = 0; t < sdc::kTailCap / 32; ++t) { const int i = lane + 32 * t; w[t] = i < len ? sdc::edit_at(B, ed, i) : 0.f; }
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < sdc::kTailCap / 32; ++t) { const int i = lane + 32 * t; if (i < len) B[i] = w[t]; }
                    __syncwarp();
                }
                if (lane == l) { bd.rm[sd] = rm; bd.pos[sd] = pos; }
            }
        }
        if (active && S.append_history) sdc::reward_plan_c(S, env, rq, M, bp, bd);
        __syncwarp();
        const long long tk2 = clock64();
        // A pass whose result this step's reward does not need (the brackets still hold the quartile ranks and the tail
        // bands still contain the fences: the refresh only restores slack for FUTURE steps) is a maintenance pass: its
        // record goes to a global queue served by whichever CTA is free.  The rare env that cannot price this step
        // without the window (slow_lane) publishes the same record with the inputs of reward_finish: the pass CTA prices it.
        const bool wants_pass = active && rq.kind != sdc::SCAN_SKIP;
        // the window of an env that asks for a pass is on its way into L2 before a worker CTA picks the job up
        if (wants_pass) l2_prefetch_bulk(S.hist + (size_t)env * S.hist_cap, ((unsigned)rq.n * 4u + 15u) & ~15u);
        const bool moments_needed = rq.n >= 2 && !rq.degenerate;
        const bool slow_lane = wants_pass && ((moments_needed && !M.ok) || rq.dir[0] || rq.dir[1]);
        {
            const unsigned slow = __ballot_sync(0xffffffffu, wants_pass);
            n_edits += __popc(slow) << 16;
            const unsigned n_refresh = __popc(__ballot_sync(0xffffffffu, active && rq.kind == sdc::SCAN_REFRESH));
            const unsigned n_lists = __popc(__ballot_sync(0xffffffffu, active && (rq.rc[0] || rq.rc[1])));
            const unsigned n_tails = __popc(__ballot_sync(0xffffffffu, active && rq.kind == sdc::SCAN_REFRESH && !M.ok && rq.tails));
            if (lane == 0) {
                if (slow) {          // statistics only
                    atomicAdd(a.ctr + 4, __popc(slow) - (int)n_refresh); atomicAdd(a.ctr + 5, (int)n_refresh);
                    atomicAdd(a.ctr + 6, (int)n_lists); atomicAdd(a.ctr + 7, (int)n_tails);
                    atomicAdd(a.pass_total + 0, (unsigned long long)(__popc(slow) - (int)n_refresh)); atomicAdd(a.pass_total + 1, (unsigned long long)n_refresh);
                    atomicAdd(a.pass_total + 2, (unsigned long long)n_lists); atomicAdd(a.pass_total + 3, (unsigned long long)n_tails);
                }
            }
        }
        float r3[3] = {0.f, 0.f, 0.f};
        if (active) {
            const int kind = rq.kind;
            if (!slow_lane) {
                // the bracket cursors are final here (pricing from the incremental state does not move them) and must be in
                // memory before the pass is published: a re-centring pass overwrites them
                reinterpret_cast<int2*>(S.q_a)[env] = make_int2(Q.a[0], Q.a[1]);
                reinterpret_cast<int2*>(S.q_m)[env] = make_int2(Q.m[0], Q.m[1]);
            }
            if (wants_pass) {
                // Publish the pass: record first, then its tag.  The pass CTA rewrites q_a / q_m / brackets / bands / moments of
                // this env, all of which this lane (and, for the edits, its warp: __syncwarp above) has finished writing.  A
                // slow lane also hands over what reward_finish needs: the pass CTA prices the step, stores the rewards and the
                // bracket cursors and adds the reward sums -- this warp does not wait for it.
                const int idx = atomicAdd(a.ctr + 10, 1);
                PassJob J;
                J.env = env; J.n = rq.n; J.kind = kind; J.finish = slow_lane ? 1 : 0; J.lo64 = rq.lo64; J.hi64 = rq.hi64;
                J.lo = rq.lo; J.hi = rq.hi; J.shift = rq.shift; J.tl = rq.tl; J.th = rq.th; J.tl2 = rq.tl2; J.th2 = rq.th2;
                J.dir[0] = slow_lane ? rq.dir[0] : 0; J.dir[1] = slow_lane ? rq.dir[1] : 0;
                J.thr[0] = slow_lane ? rq.thr[0] : 0.f; J.thr[1] = slow_lane ? rq.thr[1] : 0.f;
                J.rc[0] = rq.rc[0]; J.rc[1] = rq.rc[1]; J.k[0] = rq.k[0]; J.k[1] = rq.k[1];
                J.ca[0] = rq.ca[0]; J.ca[1] = rq.ca[1]; J.cb[0] = rq.cb[0]; J.cb[1] = rq.cb[1];
                J.tails = rq.tails; J.degenerate = rq.degenerate;
                J.q_a[0] = Q.a[0]; J.q_a[1] = Q.a[1]; J.q_m[0] = Q.m[0]; J.q_m[1] = Q.m[1];
                J.m_ok = M.ok; J.alt3[0] = alt3[0]; J.alt3[1] = alt3[1]; J.alt3[2] = alt3[2];
                J.q1 = rq.q1; J.m_c1 = M.c1; J.m_c2 = M.c2; J.m_c0 = M.c0;
                J.energy = en.energy; J.nci_next = en.nci_next; J.ls_penalty = en.ls_penalty;
                J.t_pub = a.unit_log ? gtime_ns() : 0ull;
                reinterpret_cast<PassJob*>(reinterpret_cast<unsigned char*>(a.pass_jobs) + (size_t)idx * sdc::kPassJobBytes)[0] = J;
                __threadfence();
                *reinterpret_cast<volatile int32_t*>(a.pass_ready + idx) = a.seq;
            }
            if (!slow_lane) {
                rq.kind = sdc::SCAN_SKIP;          // no pass results to apply: price the step from the incremental state
                sdc::reward_finish(S, env, rq, rs, M, en, alt3, Q, r3);
            }
        }
        {
            // rewards leave as three contiguous 128-byte rows per unit (the rows may be pinned host memory, where 4-byte
            // stores at a 12-byte stride are three partial writes per sector); the slots of envs priced by a pass CTA are skipped
            const unsigned own = __ballot_sync(0xffffffffu, active && !slow_lane);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int idx = lane + 32 * k, src = idx / 3, comp = idx - 3 * src;
                const float v0 = __shfl_sync(0xffffffffu, r3[0], src), v1 = __shfl_sync(0xffffffffu, r3[1], src), v2 = __shfl_sync(0xffffffffu, r3[2], src);
                if ((own >> src) & 1u) a.rew[(size_t)env0 * 3 + idx] = comp == 0 ? v0 : (comp == 1 ? v1 : v2);
            }
        }
        const long long tk2b = clock64();
        const int finished = st.terminal;
        if (have_unit) {
            // logger sums (harl/envs/sustaindc/sustaindc_logger.py:86-101): warp reduce, one atomic per metric.  Done while
            // the step's results are still in registers (after the long observation code they come back from spills).
            double m[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) m[k] = 0.0;
            if (active) {
                m[0] = st.energy; m[1] = st.co2; m[2] = st.water; m[3] = st.tasks_in_queue; m[4] = st.tasks_dropped;
                m[5] = st.ite_kw; m[6] = st.ct_kw; m[7] = st.comp_kw; m[8] = st.hvac_kw; m[9] = 1.0; m[10] = st.terminal;
                m[11] = st.overdue; m[12] = st.total_kw;
                // reward sums ride along (the envs priced by a pass CTA contribute zero here and add theirs there)
                m[13] = (double)r3[0] + r3[1] + r3[2]; m[14] = (double)r3[0]; m[15] = (double)r3[1];
                if (st.hvac_kw > 0.0) {             // fire-and-forget reduction; the logger's p90 comes from these bins
                    int bin = (int)(st.hvac_kw * (double)a.hvac_bins_per_kw);
                    bin = bin < 0 ? 0 : (bin >= SDC_HVAC_BINS ? SDC_HVAC_BINS - 1 : bin);
                    atomicAdd(a.hvac_hist + bin, 1ull);
                }
            }
            constexpr int slot[16] = {sdc::M_ENERGY, sdc::M_CO2, sdc::M_WATER, sdc::M_TASKS_IN_QUEUE, sdc::M_TASKS_DROPPED,
                                      sdc::M_ITE_KW, sdc::M_CT_KW, sdc::M_COMP_KW, sdc::M_HVAC_KW, sdc::M_STEPS, sdc::M_EPISODES,
                                      sdc::M_OVERDUE, sdc::M_TOTAL_KW, sdc::M_REWARD_SUM, sdc::M_REWARD_LS, sdc::M_REWARD_DC};
            const double tot = warp_sum16(m, lane);              // lane l: the sum of metric warp_sum16_slot(l)
            const int k = warp_sum16_slot(lane);
            if (!(lane & 1)) atomicAdd(a.metrics + slot[k], tot);
        }
        {
            // The 29 distinct observation values of every env go through a shared-memory tile of compact rows (odd row stride:
            // conflict-free) and leave as contiguous 128-bit stores: the padded [3][26] rows and the shared row are column
            // selections of the compact row (p2c table), the compact output IS the tile.  A 29-column tile instead of a
            // 79-column one is what lets the kernel run with 54 KB of shared memory per CTA, i.e. with twice the L1.
            float* tile = reinterpret_cast<float*>(smem_raw + table_bytes) + (size_t)warp * 32 * kTileStride;
            if (active) {
                CompactSink sink{tile + lane * kTileStride};
                sdc::emit_obs(S, T, env, od, sink);
                a.done[env] = (uint8_t)finished;
            }
            __syncwarp();
            if (have_unit) {
                const int n_here = min(U, N - env0);
                if (a.obs) {
                    const int total = n_here * kObsRow;
                    float4* dst4 = reinterpret_cast<float4*>(a.obs + (size_t)env0 * kObsRow);      // env0 is a multiple of 8
                    for (int i = lane; i < total / 4; i += 32) {
                        float v[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int f = 4 * i + q; const int e = f / kObsRow; const int c = p2c[f - e * kObsRow];
                            v[q] = c >= 0 ? tile[e * kTileStride + c] : 0.f;
                        }
                        dst4[i] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                    for (int f = (total & ~3) + lane; f < total; f += 32) {
                        const int e = f / kObsRow; const int c = p2c[f - e * kObsRow];
                        a.obs[(size_t)env0 * kObsRow + f] = c >= 0 ? tile[e * kTileStride + c] : 0.f;
                    }
                }
                if (a.share) {
                    float* dsh = a.share + (size_t)env0 * SDC_SHARE_DIM;
                    for (int f = lane; f < n_here * SDC_SHARE_DIM; f += 32) {
                        const int e = f / SDC_SHARE_DIM, k = f - e * SDC_SHARE_DIM;
                        dsh[f] = k < 28 ? tile[e * kTileStride + k] : 0.f;      // ls[0:26] | dc[11] | dc[13] | the battery row's padding zero
                    }
                }
                if (a.obs_c) {
                    // compact rows: 116 B instead of the 428 B of obs + share, which is what a host caller pays for over the host
                    // link -- the tile's own layout, copied as it is
                    const int total = n_here * SDC_OBS_COMPACT;
                    float4* dst4 = reinterpret_cast<float4*>(a.obs_c + (size_t)env0 * SDC_OBS_COMPACT);   // env0 * 29 * 4 B: 16-byte aligned for env0 % 4 == 0
                    const float4* src4 = reinterpret_cast<const float4*>(tile);                             // warp * 32 * 29 * 4 B: 16-byte aligned
                    for (int i = lane; i < total / 4; i += 32) dst4[i] = src4[i];
                    for (int f = (total & ~3) + lane; f < total; f += 32) a.obs_c[(size_t)env0 * SDC_OBS_COMPACT + f] = tile[f];
                }
                unsigned fin = __ballot_sync(0xffffffffu, finished != 0);
                while (fin && (a.term_obs || a.term_c)) {
                    const int l = __ffs(fin) - 1;
                    fin &= fin - 1;
                    store_env_rows_compact(tile + l * kTileStride, env0 + l, OutPtrs{a.term_obs, nullptr, a.term_c}, lane, 32, p2c);
                }
            }
        }
        const long long tk2c = clock64();
        const long long tk3 = tk2c;
        // Finished envs.  The normal case: the look-ahead generation of the previous launch staged the next episode WITH its
        // reset observation, and the reset is done right here by the env's own warp -- flip the weather buffer, clear the
        // queue ring, copy the staged observation over the terminal one (which already went to term_obs), reset the
        // scalars.  Everything else (first resets, episodes staged by the host without an observation) goes to the reset
        // workers (other CTAs of this launch); there order matters: the terminal observation is in global memory before the
        // env is published, because the worker overwrites obs/share.  Only the ~5 % of units with a finished env get here.
        if (__any_sync(0xffffffffu, finished)) {
            const bool fast = finished && S.pend_valid[env] == 3;
            unsigned todo = __ballot_sync(0xffffffffu, fast);
            __syncwarp();                               // the owner lanes' state stores of this step precede the reset's
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                const int e = env0 + l;
                uint4* ring = reinterpret_cast<uint4*>(S.ls_ring + (size_t)e * (S.ls_mask + 1));
                for (int k = lane; k < (S.ls_mask + 1) / 16; k += 32) ring[k] = make_uint4(0u, 0u, 0u, 0u);
                store_env_rows(S.pend_obs + (size_t)e * kObsRow, e, OutPtrs{a.obs, a.share, a.obs_c}, lane, 32);
                const int loc_l = __shfl_sync(0xffffffffu, od.loc, l);
                if (lane == 0) {
                    // every load first (one round trip to lines the unit prefetched), then the stores: loads interleaved with
                    // stores to arrays the compiler cannot prove distinct run one after the other, and the envs that finish
                    // are the stragglers of a step
                    const double tmin = S.pend_tmin[e], tmax = S.pend_tmax[e];
                    const int pd = S.pend_day[e], ph = S.pend_hour[e];
                    const uint8_t cb = S.cur_buf[e];
                    const uint32_t ep = S.episode[e];
                    S.t_min[e] = tmin; S.t_max[e] = tmax;
                    S.cur_buf[e] = cb ^ 1;
                    S.pend_valid[e] = 0;
                    S.episode[e] = ep + 1;
                    sdc::reset_scalars(S, e, pd * 96 + ph * 4, &T.loc[loc_l]);
                }
            }
            if (__any_sync(0xffffffffu, finished && !fast)) {
                __threadfence();
                if (finished && !fast) a.reset_list[atomicAdd(a.ctr + 1, 1)] = env;
                __threadfence();
            }
            __syncwarp();
        }
        // look-ahead: envs that will finish two steps from now -> pre-generation list consumed by the next launch
        if (active && st.step_after + 2 == S.ep_len) a.pre_list[atomicAdd(a.ctr + 8, 1)] = env;
        __syncwarp();
        if (lane == 0 && have_unit) { __threadfence(); atomicAdd(a.ctr + 2, 1); }     // after every publication of this unit
        if (a.phase_clocks && lane == 0 && have_unit) {
            const long long tk4 = clock64();
            atomicAdd(a.phase_clocks + 0, (unsigned long long)(tk1 - tk0));   // load shifting + data centre + battery
            atomicAdd(a.phase_clocks + 1, (unsigned long long)(tk2 - tk1));   // window append, brackets, moments, tail bands
            atomicAdd(a.phase_clocks + 8, (unsigned long long)(tk2b - tk2));  // rewards of the incremental lanes
            atomicAdd(a.phase_clocks + 9, (unsigned long long)(tk2c - tk2b)); // observations + sums
            atomicAdd(a.phase_clocks + 2, (unsigned long long)(tk3 - tk2c));  // waiting for the CTA + window passes
            atomicAdd(a.phase_clocks + 3, (unsigned long long)(tk4 - tk3));   // rewards after passes, hand-overs
            atomicAdd(a.phase_clocks + 4, 1ull);                              // units
            atomicMax(a.phase_clocks + 14, gtime_ns());                       // timeline: last unit done
            atomicMax(a.phase_clocks + 10, (unsigned long long)(tk4 - tk0));  // slowest unit (clocks)
            atomicMax(a.phase_clocks + 11, (unsigned long long)(tk2c - tk0)); // slowest unit up to its observations
            if (a.unit_log) {
                unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                uint32_t* L = a.unit_log + (size_t)unit * 8;
                L[0] = (uint32_t)(tk1 - tk0); L[1] = (uint32_t)(tk2 - tk1); L[2] = (uint32_t)(tk2b - tk2); L[3] = (uint32_t)(tk2c - tk2b);
                L[4] = (uint32_t)(tk4 - tk2c); L[5] = smid | ((uint32_t)blockIdx.x << 16); L[6] = (uint32_t)(t_unit0 & 0xffffffffu); L[7] = n_edits;
            }
        }
    }

    // ---------------- workers: every CTA becomes one once it has no unit left ----------------
    // CTAs beyond n_unit_ctas start here immediately, so episode generation, maintenance passes and resets overlap
    // with the other CTAs' units.
    __shared__ ResetShared rsh;
    __shared__ int s_env;
    __syncthreads();
    double* runbuf = reinterpret_cast<double*>(smem_raw + table_bytes);   // the window-pass buffers are free by now
    // A worker always holds one ticket of the reset queue and one of the maintenance-pass queue (a ticket is a slot
    // index; it is served as soon as the slot is filled) and takes look-ahead generation jobs -- which were published by
    // the PREVIOUS launch and are therefore available from the first cycle on -- when neither is ready.  It leaves when
    // every unit is past its hand-overs and its tickets lie beyond what was published.
    {
        const int n_pre = *reinterpret_cast<volatile const int32_t*>(a.ctr_prev + 8);
        __shared__ int s_kind;                  // 0 exit, 1 reset, 2 maintenance pass, 3 generation
        int my_reset = -1, my_pass = -1;        // thread 0 only
        bool pre_left = n_pre > 0;
        for (;;) {
            if (threadIdx.x == 0) {
                volatile int32_t* list = a.reset_list;
                volatile int32_t* ready = a.pass_ready;
                volatile int32_t* units_done = a.ctr + 2;
                int kind = -1, env = -1;
                while (kind < 0) {
                    if (my_reset < 0) my_reset = atomicAdd(a.ctr + 3, 1);
                    if (my_pass < 0) my_pass = atomicAdd(a.ctr + 11, 1);
                    const bool all_done = *units_done >= n_units;       // read BEFORE the slots: nothing is published after it
                    if (all_done) __threadfence();
                    const int e = list[my_reset];
                    if (e >= 0) { list[my_reset] = -1; my_reset = -1; kind = 1; env = e; break; }
                    // longest jobs first: a generation (44 us) started late would be the tail of the launch
                    if (pre_left) {
                        const int idx = atomicAdd(a.ctr + 9, 1);
                        if (idx < n_pre) { kind = 3; env = a.pre_list_prev[idx]; break; }
                        pre_left = false;
                    }
                    if (my_pass < N && ready[my_pass] == a.seq) { __threadfence(); env = my_pass; my_pass = -1; kind = 2; break; }
                    if (all_done) { kind = 0; break; }
                    __nanosleep(200);
                }
                s_kind = kind; s_env = env;
            }
            __syncthreads();
            const int kind = s_kind, env = s_env;
            __syncthreads();
            if (kind == 0) break;
            const long long tw0 = clock64();
            if (kind == 1) {
                __threadfence();
                reset_one_env(S, env, OutPtrs{a.obs, a.share, a.obs_c}, runbuf, rsh);
                __syncthreads();
            } else if (kind == 2) {
                const int* src = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(a.pass_jobs) + (size_t)env * sdc::kPassJobBytes);
                int* dst = reinterpret_cast<int*>(&ps.job);
                for (int i = threadIdx.x; i < (int)(sizeof(PassJob) / 4); i += kStepThreads) dst[i] = __ldcg(src + i);
                __syncthreads();
                // diagnostics: the phases of pass job `env` (a ticket index) go to the upper half of the unit log
                uint32_t* clk_log = (a.unit_log && env < 4096) ? a.unit_log + ((size_t)(((S.n_envs + 7) / 8) / 2) + env) * 8 : nullptr;
                if (clk_log && (size_t)(clk_log - a.unit_log) + 8 > (size_t)((S.n_envs + 7) / 8) * 8) clk_log = nullptr;
                window_pass(S, ps, win, scr, hits, hit_cap, pass_phase, clk_log);
                pass_phase ^= 1u;
                if (threadIdx.x == 0 && ps.job.finish) finish_step(S, a, ps.job);
            } else {
                pregen_one_env(S, env, runbuf, rsh);
                __syncthreads();
            }
            if (a.phase_clocks && threadIdx.x == 0) {
                const int slot = kind == 1 ? 5 : (kind == 2 ? 7 : 6);            // clocks: resets, generation, passes
                atomicAdd(a.phase_clocks + slot, (unsigned long long)(clock64() - tw0));
                atomicAdd(a.phase_clocks + 12, 1ull << (16 * (kind - 1)));       // packed job counts (16 bits each)
            }
        }
    }
    if (a.phase_clocks && threadIdx.x == 0) atomicMax(a.phase_clocks + 15, gtime_ns());   // timeline: CTA done
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
__global__ void __launch_bounds__(kSortThreads) k_rebuild(const __grid_constant__ sdc::State S) {
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
        S.tail_n[env * 2 + j] = -1;                // moments / tail sets are rebuilt by the env's next step
    }
}

// =================================================================================================
// launches
// =================================================================================================
constexpr int kMaxDynSmem = 100 * 1024;           // opt-in dynamic shared memory of k_step / k_obs / k_reset
constexpr int kMaxWinLen = (kMaxDynSmem - kTableBytes) / (int)sizeof(double);   // walk buffer of the longest supported episode
static int max_window_len() { return kMaxWinLen; }
static size_t run_buf_bytes(const sdc::State& S) { return (size_t)(S.win_len > kNormWindow ? S.win_len : kNormWindow) * sizeof(double); }

static const char* launch_step(Context& c, const sdc::State& S, const StepArgs& a, void* stream) {
    const int U = a.unit_envs;
    // Dynamic shared memory of k_step after the tables: collect scratch + the staged window (which also receives the sorted
    // collections: at least 2 x kCollectCap floats) + the parked-hit list (worker role); the unit role uses the start of the
    // same region for its compact observation tiles (30 KB).  54 KB per CTA: two CTAs leave the SM a 124 KB L1.
    const size_t pass_floats = (size_t)2 * sdc::kCollectCap + 2 * sdc::kTailCap + (S.hist_cap > 2 * sdc::kCollectCap ? S.hist_cap : 2 * sdc::kCollectCap);
    const size_t tile_floats = (size_t)kWarpsPerBlock * 32 * kTileStride;
    constexpr size_t kHitFloats = 2048;            // parked hits of a pass (256 per warp; overflow is classified in place), then the sorted
                                                   // bands + the bucket order of the collected values
    size_t smem_floats = pass_floats + kHitFloats > tile_floats ? pass_floats + kHitFloats : tile_floats;
    const int hit_cap = (int)(smem_floats - pass_floats);
    if (hit_cap < 2 * sdc::kTailCap + sdc::kCollectCap) return "k_step: shared memory layout leaves no room for the sorted bands and the bucket order";
    size_t smem = smem_floats * sizeof(float);
    if (smem < run_buf_bytes(S)) smem = run_buf_bytes(S);                              // episode generation reuses the region
    // shared-memory copy of the location / dc parameter tables: only what this handle needs
    int table_bytes = (int)(S.n_loc * sizeof(sdc::LocTables) + S.n_cfg * sizeof(sdc_dc_params));
    table_bytes = table_bytes <= kTableBytes ? (table_bytes + 127) & ~127 : 0;
    smem += table_bytes;
    if (smem > (size_t)kMaxDynSmem) return "k_step: shared memory budget exceeded";
    const int n_units = (S.n_envs + U - 1) / U;
    // All CTAs must be co-resident: workers spin on counters that unit CTAs advance.  The grid is sized from the occupancy
    // the runtime reports for this kernel / block size / shared memory, and the launch is COOPERATIVE: the driver then
    // gang-schedules the grid (or refuses the launch), also when other kernels -- an NCCL collective, a policy forward on
    // another stream -- hold part of the device.
    const void* fn = (const void*)k_step;
    CU(cudaSetDevice(c.device));
    const int which = 0;
    if (c.occ_per_sm[which] == 0 || c.occ_smem[which] != smem) {
        int q = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, fn, kStepThreads, smem));
        c.occ_per_sm[which] = q; c.occ_smem[which] = smem;
    }
    int per_sm = c.occ_per_sm[which];
    if (per_sm < 1) return "k_step: no resident CTA fits on an SM";
    const int design = 2;                             // __launch_bounds__ of the kernel
    if (per_sm > design) per_sm = design;
    const int bps = a.blocks_per_sm > 0 ? a.blocks_per_sm : per_sm;
    const int capacity = c.sm_count * (bps < per_sm ? bps : per_sm);
    const int need = (n_units + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int reserve = capacity / 8;                       // at least this many CTAs are workers from the first cycle on
    if (reserve < 1) reserve = 1;
    int n_unit_ctas = need < capacity - reserve ? need : capacity - reserve;
    if (n_unit_ctas < 1) n_unit_ctas = 1;
    // large batches fill the chip (every CTA beyond the unit CTAs is a worker); small ones do not launch a whole grid
    int blocks = S.n_envs < 4096 ? n_unit_ctas + 4 : capacity;
    if (blocks > capacity) blocks = capacity;
    cudaStream_t st = (cudaStream_t)stream;
    int hc = hit_cap, tb = table_bytes, nuc = n_unit_ctas;
    void* args[] = {(void*)&S, (void*)&a, (void*)&nuc, (void*)&hc, (void*)&tb};
    if (c.cooperative) CU(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(kStepThreads), args, smem, st));
    else CU(cudaLaunchKernel(fn, dim3(blocks), dim3(kStepThreads), args, smem, st));
    CU(cudaGetLastError());
    return nullptr;
}

static const char* launch_reset(Context& c, const sdc::State& S, const int32_t* list, const int32_t* count, float* obs, float* share,
                                void* stream) {
    const size_t smem = run_buf_bytes(S);
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
    CU(cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    CU(cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    CU(cudaFuncSetAttribute(k_rebuild, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    return nullptr;
}

}  // namespace backend

#include "sdc_api.inc"
