"""ctypes binding of the C ABI in include/sdc_b200.h (libsdc_b200.so, CUDA sm_100a).

There is no CPU fallback: if the shared library is missing or cannot be loaded, importing the engine
raises.  ``load(path)`` exists so that tests can bind the same prototypes to the test-only host build of
the device logic (tests/hostsim); the package itself always loads ``csrc/libsdc_b200.so``.
"""
import ctypes as C
import os

MAX_RACK_CLASSES = 32
N_AGENTS, OBS_DIM, SHARE_DIM, INFO_STRIDE, OBS_COMPACT = 3, 26, 29, 64, 29
YEAR_STEPS, TRACE_PAD, HIST_CAP, N_METRICS = 35040, 64, 10000, 16
LIST_CAP, TAIL_CAP = 128, 128
HVAC_BINS = 4096          # sdc_core.h kListCap / kTailCap (state inspection only)
ABI_VERSION = 1

F_WORKLOAD_RANGE, F_CPU_LOAD_RANGE, F_OUTLET_DELTA, F_TRACE_DOMAIN, F_BRACKET, F_NONFINITE, F_BATTERY, F_REWARD_DOMAIN = (
    0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40, 0x80)
# reward method ids (SDC_R_*), keyed by the reference's names (utils/reward_creator.py:322-334)
REWARD_METHOD_IDS = {"default_ls_reward": 0, "default_dc_reward": 1, "default_bat_reward": 1, "custom_agent_reward": 2, "tou_reward": 3,
                     "energy_efficiency_reward": 4, "energy_PUE_reward": 5, "water_usage_efficiency_reward": 6}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libsdc_b200.so")


class Config(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("device", C.c_int32), ("ep_len", C.c_int32), ("n_loc", C.c_int32),
                ("n_cfg", C.c_int32), ("hist_cap", C.c_int32), ("unit_envs", C.c_int32), ("reserved", C.c_int32)]


class Location(C.Structure):
    _fields_ = [("workload", C.c_void_p), ("ns_tasks", C.c_void_p), ("sh_tasks", C.c_void_p), ("ci", C.c_void_p),
                ("ci_min30", C.c_void_p), ("ci_max30", C.c_void_p), ("temp_base", C.c_void_p), ("wetb_base", C.c_void_p)]


_D32 = C.c_double * MAX_RACK_CLASSES


class DcParams(C.Structure):
    _fields_ = [("n_classes", C.c_int32), ("n_racks", C.c_int32),
                ("cls_full", _D32), ("cls_idle", _D32), ("cls_ncpu", _D32), ("cls_supply", _D32), ("cls_mult", _D32),
                ("ret_mean", C.c_double),
                ("m_cpu", C.c_double), ("c_cpu", C.c_double), ("shift_cpu", C.c_double),
                ("m_fan", C.c_double), ("c_fan", C.c_double), ("shift_fan", C.c_double),
                ("itfan_ref_p", C.c_double), ("itfan_ref_v_ratio", C.c_double), ("itfan_full_load_v", C.c_double),
                ("c_air", C.c_double), ("rho_air", C.c_double), ("crac_supply_flow_pu", C.c_double),
                ("cw_pump_w", C.c_double), ("ct_pump_w", C.c_double),
                ("ctafr", C.c_double), ("ct_fan_ref_p", C.c_double),
                ("power_lb_kw", C.c_double), ("power_ub_kw", C.c_double), ("bat_capacity_mwh", C.c_double)]


_P = C.c_void_p
_PROTOS = {
    "sdc_abi_version": (C.c_int, []),
    "sdc_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "sdc_destroy": (None, [_P]),
    "sdc_last_error": (C.c_char_p, [_P]),
    "sdc_set_location": (C.c_int, [_P, C.c_int32, C.POINTER(Location)]),
    "sdc_set_dc_params": (C.c_int, [_P, C.c_int32, C.POINTER(DcParams)]),
    "sdc_set_hour_table": (C.c_int, [_P, _P, _P]),
    "sdc_assign": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "sdc_set_reward_methods": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "sdc_stage_episode": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "sdc_window_len": (C.c_int, [_P]),
    "sdc_reset": (C.c_int, [_P, _P, _P, _P, _P]),
    "sdc_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "sdc_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "sdc_reset_host": (C.c_int, [_P, _P, _P, _P]),
    "sdc_step_host_begin": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "sdc_step_host_end": (C.c_int, [_P]),
    "sdc_step_compact": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "sdc_step_compact_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "sdc_step_compact_host_begin": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "sdc_host_buffers_compact": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "sdc_expand_obs": (None, [_P, C.c_int64, _P, _P]),
    "sdc_fetch_info": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "sdc_host_buffers": (C.c_int, [_P] + [C.POINTER(_P)] * 7),
    "sdc_metrics": (C.c_int, [_P, _P, C.c_int32]),
    "sdc_hvac_histogram": (C.c_int, [_P, _P, _P, C.c_int32]),
    "sdc_prefill_history": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "sdc_rebuild_brackets": (C.c_int, [_P, _P]),
    "sdc_read_state": (C.c_int64, [_P, C.c_char_p, _P, C.c_int64]),
    "sdc_write_state": (C.c_int64, [_P, C.c_char_p, _P, C.c_int64]),
    "sdc_state_bytes": (C.c_size_t, [_P]),
    "sdc_get_state": (C.c_int, [_P, _P, C.c_size_t]),
    "sdc_set_state": (C.c_int, [_P, _P, C.c_size_t]),
    "sdc_error_flags": (_P, [_P]),
    "sdc_launch_count": (C.c_int64, [_P]),
    "sdc_kernel_times": (C.c_int, [_P, _P]),
    "sdc_set_tuning": (C.c_int, [_P, C.c_char_p, C.c_int32]),
}
EXPORTS = tuple(_PROTOS)


def load(path=None):
    """Loads the shared library and attaches prototypes. Raises if it is missing (no fallback)."""
    path = path or os.environ.get("SDC_B200_LIB") or LIB_PATH
    if not os.path.isfile(path):
        raise ImportError(
            "dc_rl_b200: CUDA library %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError here == missing export
        fn.restype, fn.argtypes = res, args
    if lib.sdc_abi_version() != ABI_VERSION:
        raise ImportError("dc_rl_b200: ABI version mismatch in %s" % path)
    return lib
