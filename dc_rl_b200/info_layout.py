"""Column layout of the per-env info row ``info[N, INFO_K]`` written by the step kernel.

The columns are the scalar keys of the reference's merged info dict
(reference sustaindc_env.py:676-710: ``{**dc_info, **ls_info, **bat_info, **reward_params}``; sub-env
dicts at envs/bat_env_fwd_view.py:111-122, envs/carbon_ls.py:291-308, envs/dc_gym.py:213-229).  Vector-valued
keys are flattened (``ls_task_age_histogram`` -> 5 columns, ``forecast_CI`` -> 8 columns); the one
string-valued key ``bat_a_t`` is derived from ``bat_action`` on the host.
The same order is used by csrc/sdc_layout.h (enum InfoCol) -- keep the two in sync.
"""

INFO_SCALAR_KEYS = [
    # battery (envs/bat_env_fwd_view.py:111-122)
    "bat_action", "bat_SOC", "bat_CO2_footprint", "bat_avg_CI",
    "bat_total_energy_without_battery_KWh", "bat_total_energy_with_battery_KWh",
    "bat_max_bat_cap", "bat_dcload_min", "bat_dcload_max",
    # load shifting (envs/carbon_ls.py:291-308)
    "ls_original_workload", "ls_shifted_workload", "ls_action", "ls_norm_load_left",
    "ls_unasigned_day_load_left", "ls_penalty_flag", "ls_queue_max_len", "ls_tasks_in_queue",
    "ls_norm_tasks_in_queue", "ls_tasks_dropped", "ls_current_hour", "ls_tasks_processed",
    "ls_enforced", "ls_oldest_task_age", "ls_average_task_age", "ls_overdue_penalty",
    "ls_computed_tasks",
]
INFO_HIST_KEY = "ls_task_age_histogram"          # 5 columns
INFO_DC_KEYS = [
    # data centre (envs/dc_gym.py:213-229)
    "dc_ITE_total_power_kW", "dc_CT_total_power_kW", "dc_Compressor_total_power_kW",
    "dc_HVAC_total_power_kW", "dc_total_power_kW", "dc_crac_setpoint_delta", "dc_crac_setpoint",
    "dc_cpu_workload_fraction", "dc_int_temperature", "dc_exterior_ambient_temp", "dc_power_lb_kW",
    "dc_power_ub_kW", "dc_CW_pump_power_kW", "dc_CT_pump_power_kW", "dc_water_usage",
]
INFO_COMMON_KEYS = ["outside_temp", "day", "hour", "norm_CI"]   # sustaindc_env.py:678-683
INFO_FORECAST_KEY = "forecast_CI"                # 8 columns
INFO_TERMINAL_KEY = "isterminal"

BAT_ACTION_NAMES = {0: "charge", 1: "discharge", 2: "idle"}      # envs/bat_env_fwd_view.py:28


def _build():
    cols = list(INFO_SCALAR_KEYS)
    cols += ["%s[%d]" % (INFO_HIST_KEY, i) for i in range(5)]
    cols += INFO_DC_KEYS + INFO_COMMON_KEYS
    cols += ["%s[%d]" % (INFO_FORECAST_KEY, i) for i in range(8)]
    cols.append(INFO_TERMINAL_KEY)
    return cols


INFO_COLUMNS = _build()
INFO_K_USED = len(INFO_COLUMNS)     # 59
INFO_K = 64                         # row stride in floats (256-byte rows)
COL = {name: i for i, name in enumerate(INFO_COLUMNS)}
HIST0 = COL[INFO_HIST_KEY + "[0]"]
FORECAST0 = COL[INFO_FORECAST_KEY + "[0]"]


def info_dict_to_row(info):
    """Flattens one reference-style info dict into a list of INFO_K_USED python floats."""
    row = []
    for name in INFO_COLUMNS:
        if "[" in name:
            key, idx = name[:-1].split("[")
            row.append(float(info[key][int(idx)]))
        else:
            row.append(float(info[name]))
    return row
