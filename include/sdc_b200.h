/* sdc_b200.h -- C ABI of libsdc_b200.so: N batched SustainDC environments stepped on one B200 (sm_100a).
 *
 * Drop-in boundary.  The reference is pure Python and has no FFI; what it has is
 * the vec-env call pair that `harl.runners` drive:
 *     ShareSubprocVecEnv.reset()/step(actions)        harl/envs/env_wrappers.py:257-280
 *       -> HARLSustainDCEnv.reset()/step()             harl/envs/sustaindc/harlsustaindc_env.py:89-131
 *         -> SustainDC.reset()/step()                  sustaindc_env.py:436-531, 533-621
 * Every entry point below names the reference call it replaces.  Plain pointers and sizes only; the
 * caller (PyTorch / ctypes) owns all I/O buffers, the handle owns all persistent env state.
 *
 * All functions return 0 on success or a negative SDC_E_* code; sdc_last_error() gives the message.
 * No C++ exception crosses this boundary.  One caller thread per handle.
 */
#ifndef SDC_B200_H
#define SDC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDC_ABI_VERSION 1

/* fixed sizes of the reference path */
#define SDC_N_AGENTS 3          /* agent_ls, agent_dc, agent_bat            sustaindc_env.py:107 */
#define SDC_OBS_DIM 26          /* rows zero-padded to the widest agent      harlsustaindc_env.py:25-26 */
#define SDC_SHARE_DIM 29        /* nonoverlapping shared obs                 harlsustaindc_env.py:78-85 */
#define SDC_OBS_COMPACT 29      /* the DISTINCT observation values of an env: agent_ls[26] | workload(t+1) | norm T(t+1) | SoC */
#define SDC_INFO_STRIDE 64      /* info table rows (59 used, see info_layout.py) */
#define SDC_YEAR_STEPS 35040    /* 365 d x 96 quarter-hours                  utils/managers.py:183-185 */
#define SDC_TRACE_PAD 64        /* readable slack after the last trace sample */
#define SDC_MAX_RACK_CLASSES 32
#define SDC_HIST_CAP 10000      /* reward window                             utils/reward_creator.py:5 */
#define SDC_N_METRICS 16
#define SDC_HVAC_BINS 4096

/* error codes */
#define SDC_OK 0
#define SDC_E_ARG (-1)
#define SDC_E_CUDA (-2)
#define SDC_E_STATE (-3)
#define SDC_E_NOMEM (-4)

/* per-env error flag bits (sdc_error_flags); the env clamps and continues, SURVEY.md section 4 */
#define SDC_F_WORKLOAD_RANGE 0x1   /* envs/carbon_ls.py:333-336  workload outside [0,1]            */
#define SDC_F_CPU_LOAD_RANGE 0x2   /* envs/dc_gym.py:288-290     shifted workload outside [0,1]    */
#define SDC_F_OUTLET_DELTA 0x4     /* envs/datacenter.py:295-300 rack outlet - inlet < 2 C          */
#define SDC_F_TRACE_DOMAIN 0x8     /* trace index outside [16, 35040-17] (reference crashes there) */
#define SDC_F_BRACKET 0x10         /* internal: quartile bracket invariant broken (bug guard)      */
#define SDC_F_NONFINITE 0x20       /* non-finite energy value                                      */
#define SDC_F_BATTERY 0x40         /* envs/bat_env_fwd_view.py:237 discharge > DC energy           */
#define SDC_F_REWARD_DOMAIN 0x80   /* utils/reward_creator.py:191 tou_reward off the full hour (KeyError there) */

/* reward methods (utils/reward_creator.py:322-334, selected per agent at sustaindc_env.py:137-144) */
#define SDC_R_DEFAULT_LS 0         /* default_ls_reward  :48-82   (agent_ls only: it alone appends to the window, :62-63) */
#define SDC_R_DEFAULT_DC 1         /* default_dc_reward / default_bat_reward :85-130 (the same function)  */
#define SDC_R_CUSTOM 2             /* custom_agent_reward :133-146 -> 0.0                                 */
#define SDC_R_TOU 3                /* tou_reward :154-202                                                  */
#define SDC_R_ENERGY_EFFICIENCY 4  /* energy_efficiency_reward :227-243                                    */
#define SDC_R_PUE 5                /* energy_PUE_reward :246-268                                           */
#define SDC_R_WATER 6              /* water_usage_efficiency_reward :297-318                               */

typedef struct sdc_env sdc_env;

typedef struct {
    int32_t n_envs;        /* N environments on this device */
    int32_t device;        /* CUDA device ordinal */
    int32_t ep_len;        /* steps per episode = days_per_episode*96   utils/managers.py:113 */
    int32_t n_loc;         /* number of location trace sets */
    int32_t n_cfg;         /* number of (dc_config, location) parameter sets */
    int32_t hist_cap;      /* reward window length, <= SDC_HIST_CAP, multiple of 4 (reference: 10000) */
    int32_t unit_envs;     /* envs handled per warp in the step kernel: 8, 16 or 32 (0 = default) */
    int32_t reserved;
} sdc_config;

/* One location's exogenous traces at 15-min resolution (utils/managers.py Workload/CI/Weather managers).
 * All arrays are HOST pointers of SDC_YEAR_STEPS + SDC_TRACE_PAD entries, copied at the call. */
typedef struct {
    const double* workload;   /* rescaled + smoothed cpu load           managers.py:220-244,268-271 */
    const uint8_t* ns_tasks;  /* ceil(w*0.8*100) evaluated in fp64      carbon_ls.py:194 */
    const uint8_t* sh_tasks;  /* floor(w*0.2*100) evaluated in fp64     carbon_ls.py:195 */
    const double* ci;         /* carbon intensity, clipped >= 0         managers.py:417 */
    const double* ci_min30;   /* min of ci[t:t+2880]                    managers.py:435-437 */
    const double* ci_max30;   /* max of ci[t:t+2880] */
    const double* temp_base;  /* dry bulb before noise                  managers.py:550 */
    const double* wetb_base;  /* wet bulb before noise                  managers.py:547 */
} sdc_location;

/* One sized data-centre description: rack classes (racks with identical parameters merged, with a
 * multiplicity), curve coefficients (envs/datacenter.py:31-49) and the init-time sizing results
 * (utils/make_envs_pyenv.py:149-218). */
typedef struct {
    int32_t n_classes;
    int32_t n_racks;
    double cls_full[SDC_MAX_RACK_CLASSES];    /* full-load W per CPU */
    double cls_idle[SDC_MAX_RACK_CLASSES];    /* idle W per CPU */
    double cls_ncpu[SDC_MAX_RACK_CLASSES];    /* CPUs per rack */
    double cls_supply[SDC_MAX_RACK_CLASSES];  /* supply approach temp, clamped [3.8,5.3] */
    double cls_mult[SDC_MAX_RACK_CLASSES];    /* number of racks in the class */
    double ret_mean;                          /* mean of RACK_RETURN_APPROACH_TEMP_LIST */
    double m_cpu, c_cpu, shift_cpu;           /* datacenter.py:36-39 */
    double m_fan, c_fan, shift_fan;           /* datacenter.py:46-49 */
    double itfan_ref_p, itfan_ref_v_ratio, itfan_full_load_v;
    double c_air, rho_air, crac_supply_flow_pu;
    double cw_pump_w, ct_pump_w;              /* datacenter.py:455-458 */
    double ctafr, ct_fan_ref_p;               /* sized: CT_REFRENCE_AIR_FLOW_RATE, CT_FAN_REF_P */
    double power_lb_kw, power_ub_kw;          /* dc_gym.py:86-87 */
    double bat_capacity_mwh;                  /* make_envs_pyenv.py:190-197 */
} sdc_dc_params;

/* ---- lifetime ------------------------------------------------------------------------------ */
int sdc_abi_version(void);
/* replaces make_train_env(...) constructing N HARLSustainDCEnv workers  harl/utils/envs_tools.py:49-74 */
int sdc_create(const sdc_config* cfg, sdc_env** out);
void sdc_destroy(sdc_env* env);
const char* sdc_last_error(sdc_env* env /* may be NULL: last create error */);

/* ---- init-time tables (host pointers, copied) ---------------------------------------------- */
/* replaces Workload_Manager/CI_Manager/Weather_Manager construction    utils/managers.py:153-197,330-387,504-569 */
int sdc_set_location(sdc_env* env, int32_t loc, const sdc_location* data);
/* replaces make_dc_pyeplus_env sizing + DC_Config                      utils/make_envs_pyenv.py:75-242 */
int sdc_set_dc_params(sdc_env* env, int32_t cfg, const sdc_dc_params* params);
/* 96-entry (cos, sin) table of sc_obs(round(hour/24,3))                utils/managers.py:66-88 */
int sdc_set_hour_table(sdc_env* env, const double* cos96, const double* sin96);
/* per-env assignment: trace set, parameter set, admissible start-day range (month) and RNG seed
 * (make_train_env month/seed rules, harl/utils/envs_tools.py:56-67; sustaindc_env.py:197-198). */
int sdc_assign(sdc_env* env, const uint8_t* loc_id, const uint8_t* cfg_id, const int16_t* day_lo,
               const int16_t* day_hi, const uint64_t* seed);

/* Reward method of each agent (SDC_R_*; default: DEFAULT_LS, DEFAULT_DC, DEFAULT_DC).  Replaces get_reward_method
 * (utils/reward_creator.py:336-349).  When agent_ls does not use default_ls_reward nothing appends to the reward window
 * (utils/reward_creator.py:62-63): the default dc / bat rewards are then computed against the empty window, i.e. 0. */
int sdc_set_reward_methods(sdc_env* env, int32_t ls, int32_t dc, int32_t bat);

/* ---- episodes ------------------------------------------------------------------------------ */
/* Replay ("injection") mode: stage the NEXT episode of `count` envs: start (day, hour), the realised
 * weather windows temp/wetb[count][win_len] (win_len = ep_len + 18, fp64, host) and the 30-day min / max
 * of the dry bulb.  Replaces the RNG draws of SustainDC.reset / Weather_Manager.reset
 * (sustaindc_env.py:454-455, utils/managers.py:594-613).  Envs without a staged episode draw their
 * start and weather noise from the device Philox generator. */
int sdc_stage_episode(sdc_env* env, int32_t count, const int32_t* env_ids, const int32_t* day,
                      const int32_t* hour, const double* temp_win, const double* wetb_win,
                      const double* t_min30, const double* t_max30);
int sdc_window_len(sdc_env* env);

/* SustainDC.reset / HARLSustainDCEnv.reset for the envs whose mask byte is non-zero (NULL = all).
 * mask, obs[N,3,26], share[N,29] are DEVICE pointers; work is enqueued on `stream` (cudaStream_t). */
int sdc_reset(sdc_env* env, const uint8_t* mask_dev, float* obs_dev, float* share_dev, void* stream);

/* SustainDC.step + HARL view + vec-env auto-reset (env_wrappers.py:173-192) for all N envs.
 * actions[N,3] int32; obs[N,3,26]; share[N,29]; rew[N,3]; done[N]; info[SDC_INFO_STRIDE][N]
 * (column-major table, may be NULL); term_obs[N,3,26] receives the pre-reset obs of finished envs
 * (may be NULL).  All DEVICE pointers. */
int sdc_step(sdc_env* env, const int32_t* actions_dev, float* obs_dev, float* share_dev, float* rew_dev,
             uint8_t* done_dev, float* info_dev, float* term_obs_dev, void* stream);

/* Same call with HOST buffers: H2D of actions, step, D2H of obs/share/rew/done (+info if non-NULL)
 * through the handle's pinned staging and its own stream; returns after the results are in place.
 * This is the reference-facing vec-env `step` (numpy in, numpy out). */
int sdc_step_host(sdc_env* env, const int32_t* actions, float* obs, float* share, float* rew,
                  uint8_t* done, float* info, float* term_obs);
int sdc_reset_host(sdc_env* env, const uint8_t* mask, float* obs, float* share);
/* The same call in two halves -- what ShareVecEnv.step_async / step_wait are (harl/envs/env_wrappers.py:104-127): _begin
 * enqueues the copies and the step on the handle's stream and returns, sdc_step_host_end waits and delivers. */
int sdc_step_host_begin(sdc_env* env, const int32_t* actions, float* obs, float* share, float* rew, uint8_t* done, float* info,
                        float* term_obs);
int sdc_step_host_end(sdc_env* env);

/* Compact outputs: obsc[N, SDC_OBS_COMPACT] holds the 29 DISTINCT values of an env's three observations.  The reference builds
 * agent_dc's 14 and agent_bat's 13 entries from the same time / carbon-intensity / workload / temperature values as agent_ls's
 * 26 (sustaindc_env.py:302-433): dc = ls[0:10] | ls[13] | workload(t+1) | ls[14] | norm T(t+1), bat = ls[0:10] | ls[13] | ls[14] | SoC.
 * So obsc = agent_ls[26] | workload(t+1) | norm T(t+1) | SoC -- the HARL shared row (harlsustaindc_env.py:78-85) with the
 * battery's SoC in its last slot instead of the padding zero.  The padded [3,26] rows and the shared row are column
 * selections of these 29 floats: sdc_expand_obs rebuilds both on the host, bit for bit.  termc likewise for finished envs.
 * 129 bytes per env-step cross PCIe instead of 441. */
int sdc_step_compact(sdc_env* env, const int32_t* actions_dev, float* obsc_dev, float* rew_dev, uint8_t* done_dev, float* info_dev,
                     float* termc_dev, void* stream);
int sdc_step_compact_host(sdc_env* env, const int32_t* actions, float* obsc, float* rew, uint8_t* done, float* info, float* termc);
int sdc_step_compact_host_begin(sdc_env* env, const int32_t* actions, float* obsc, float* rew, uint8_t* done, float* info, float* termc);
int sdc_host_buffers_compact(sdc_env* env, float** obsc, float** termc);      /* page-locked, like sdc_host_buffers */
void sdc_expand_obs(const float* obsc, int64_t n, float* obs /*[n,3,26] or NULL*/, float* share /*[n,29] or NULL*/);
/* With sdc_set_tuning(env, "lazy_info", 1) a host step keeps its info table on the device; this copies columns
 * [first_col, first_col + n_cols) of the LAST host step to out[n_cols][N] (host).  The runners read a dozen of the 59
 * columns per step (harl/envs/sustaindc/sustaindc_logger.py:86-101), not 16 MB. */
int sdc_fetch_info(sdc_env* env, int32_t first_col, int32_t n_cols, float* out);
/* The handle's own page-locked HOST buffers, laid out like the sdc_step_host arguments (actions[N,3] int32,
 * obs[N,3,26], share[N,29], rew[N,3], done[N], info[SDC_INFO_STRIDE][N], term_obs[N,3,26]); valid until
 * sdc_destroy.  Passing these same pointers to sdc_step_host / sdc_reset_host makes the transfers go
 * straight between them and the device (no staging copy on the host): this is how the Python vec-env hands
 * numpy views to the runner.  Any of the out-pointers may be NULL. */
int sdc_host_buffers(sdc_env* env, int32_t** actions, float** obs, float** share, float** rew, uint8_t** done,
                     float** info, float** term_obs);

/* ---- metrics / state ----------------------------------------------------------------------- */
/* Running sums since the last call with clear!=0 (SustainDCLogger.per_step,
 * harl/envs/sustaindc/sustaindc_logger.py:86-101): out[SDC_N_METRICS] doubles, HOST pointer. */
int sdc_metrics(sdc_env* env, double* out, int32_t clear);
/* Histogram of every positive dc_HVAC_total_power_kW sample since the last clear: SDC_HVAC_BINS equal bins over
 * [0, *range_kw) (range = the largest power upper bound of the handle's dc configs), samples beyond the range in the
 * last bin.  Replaces the list of all samples the reference logger keeps for its mean / max / 90th percentile
 * (harl/envs/sustaindc/sustaindc_logger.py:98-99,152-155); ranks sum their histograms (one all-reduce of 32 KB) instead
 * of gathering the samples.  counts[SDC_HVAC_BINS] uint64 and range_kw: HOST pointers. */
int sdc_hvac_histogram(sdc_env* env, uint64_t* counts, double* range_kw, int32_t clear);
/* Fill every env's reward window with `count` values each (host fp32 [N][count], or [count] shared by
 * all envs when per_env==0) and rebuild the quartile brackets -- benchmark / resume helper. */
int sdc_prefill_history(sdc_env* env, const float* values, int32_t count, int32_t per_env);
/* Recompute the quartile brackets of every env from its window with a full sort (debug / resume). */
int sdc_rebuild_brackets(sdc_env* env, void* stream);
/* Copies named per-env state arrays to host for tests and checkpointing. name: "t","step_in_ep",
 * "setpoint","bat_load","hist_len","hist","ls_len","err","qlist","q_a","q_m"; returns bytes written
 * or a negative code. */
int64_t sdc_read_state(sdc_env* env, const char* name, void* out, int64_t capacity_bytes);
/* Overwrites one named per-env state array from host memory (exact size required) -- resume / tests /
 * de-synchronising episode phases.  Writing "hist" or "hist_len" requires sdc_rebuild_brackets afterwards. */
int64_t sdc_write_state(sdc_env* env, const char* name, const void* src, int64_t bytes);
size_t sdc_state_bytes(sdc_env* env);
int sdc_get_state(sdc_env* env, void* blob, size_t bytes);
int sdc_set_state(sdc_env* env, const void* blob, size_t bytes);
/* device pointer to int32 error flags [N] */
const int32_t* sdc_error_flags(sdc_env* env);
/* number of kernels this handle has launched (bench gpu_launches) */
int64_t sdc_launch_count(sdc_env* env);
/* Per-kernel device timing: after sdc_set_tuning(env, "timing", 1) every sdc_step brackets its k_step launch
 * with CUDA events recorded on the launch stream.  sdc_kernel_times synchronises, writes
 * out[0] = number of timed steps, out[1] = sum of k_step ms, out[2] = 0 (episode resets run inside k_step),
 * out[3] = max k_step ms, and clears the accumulators. */
int sdc_kernel_times(sdc_env* env, double* out4);
/* knobs: "unit_envs" (8 / 16 / 32), "blocks_per_sm", "timing", "phases" (diagnostic clocks + per-unit / per-pass log),
 * "direct_host" (host calls on the handle's pinned buffers: bit 0 obs / compact obs, 1 share, 2 terminal rows, 3 rewards + dones
 * written by the kernel straight into them, 4 actions read from there; default 31 = a host step is one launch and one stream
 * synchronisation, no copy operation), "lazy_info", "clear_pass_total" */
int sdc_set_tuning(sdc_env* env, const char* key, int32_t value);

#ifdef __cplusplus
}
#endif
#endif /* SDC_B200_H */
