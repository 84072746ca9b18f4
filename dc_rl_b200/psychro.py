"""Wet-bulb temperature from dry-bulb, relative humidity and pressure (SI units).

Host-side ingest helper: the reference derives the hourly wet-bulb trace from the EPW columns with
``psychrolib.GetTWetBulbFromRelHum(T, RH/100, P)`` (reference utils/managers.py:10,530;
psychrolib==2.5.0 pinned in requirements.txt:53).  psychrolib is a third-party dependency that is
absent from the reference tree, so this module restates the published ASHRAE Handbook-Fundamentals
(2017) ch.1 formulation that psychrolib 2.x implements: saturation pressure (eqs. 5/6), humidity
ratio from wet-bulb (eqs. 33/35), dew point by Newton-Raphson on ln(p_ws) and wet bulb by bisection on
[T_dew, T_db] to 1e-3 degC.  The library's own answer is only defined to that tolerance.

Only ``dc_water_usage`` depends on the wet-bulb trace (reference envs/datacenter.py:343).
"""
import math

ZERO_C_K = 273.15
TRIPLE_POINT_C = 0.01
FREEZING_C = 0.0
TOLERANCE_C = 0.001
MAX_ITER = 100
MIN_HUM_RATIO = 1e-7
_BOUNDS_C = (-100.0, 200.0)


def _ln_sat_vap_pres(t_c):
    tk = t_c + ZERO_C_K
    if t_c <= TRIPLE_POINT_C:
        return (-5.6745359e3 / tk + 6.3925247 - 9.677843e-3 * tk + 6.2215701e-7 * tk ** 2
                + 2.0747825e-9 * tk ** 3 - 9.484024e-13 * tk ** 4 + 4.1635019 * math.log(tk))
    return (-5.8002206e3 / tk + 1.3914993 - 4.8640239e-2 * tk + 4.1764768e-5 * tk ** 2
            - 1.4452093e-8 * tk ** 3 + 6.5459673 * math.log(tk))


def sat_vap_pres(t_c):
    """Saturation vapour pressure over ice / liquid water, Pa."""
    return math.exp(_ln_sat_vap_pres(t_c))


def _dln_sat_vap_pres(t_c):
    tk = t_c + ZERO_C_K
    if t_c <= TRIPLE_POINT_C:
        return (5.6745359e3 / tk ** 2 - 9.677843e-3 + 2 * 6.2215701e-7 * tk + 3 * 2.0747825e-9 * tk ** 2
                - 4 * 9.484024e-13 * tk ** 3 + 4.1635019 / tk)
    return (5.8002206e3 / tk ** 2 - 4.8640239e-2 + 2 * 4.1764768e-5 * tk
            - 3 * 1.4452093e-8 * tk ** 2 + 6.5459673 / tk)


def hum_ratio_from_vap_pres(vap_pres, pressure):
    return max(0.621945 * vap_pres / (pressure - vap_pres), MIN_HUM_RATIO)


def sat_hum_ratio(t_c, pressure):
    return hum_ratio_from_vap_pres(sat_vap_pres(t_c), pressure)


def hum_ratio_from_wet_bulb(t_db, t_wb, pressure):
    ws = sat_hum_ratio(t_wb, pressure)
    if t_wb >= FREEZING_C:
        w = ((2501.0 - 2.326 * t_wb) * ws - 1.006 * (t_db - t_wb)) / (2501.0 + 1.86 * t_db - 4.186 * t_wb)
    else:
        w = ((2830.0 - 0.24 * t_wb) * ws - 1.006 * (t_db - t_wb)) / (2830.0 + 1.86 * t_db - 2.1 * t_wb)
    return max(w, MIN_HUM_RATIO)


def dew_point_from_vap_pres(t_db, vap_pres):
    if vap_pres < sat_vap_pres(_BOUNDS_C[0]) or vap_pres > sat_vap_pres(_BOUNDS_C[1]):
        raise ValueError("partial pressure of water vapour outside the range of validity")
    t_dp = t_db
    ln_vp = math.log(vap_pres)
    for _ in range(MAX_ITER + 1):
        t_it = t_dp
        t_dp = t_it - (_ln_sat_vap_pres(t_it) - ln_vp) / _dln_sat_vap_pres(t_it)
        t_dp = min(max(t_dp, _BOUNDS_C[0]), _BOUNDS_C[1])
        if abs(t_dp - t_it) <= TOLERANCE_C:
            break
    else:
        raise ValueError("dew point iteration did not converge")
    return min(t_dp, t_db)


def wet_bulb_from_hum_ratio(t_db, hum_ratio, pressure):
    w = max(hum_ratio, MIN_HUM_RATIO)
    vap_pres = pressure * w / (0.621945 + w)
    lo = dew_point_from_vap_pres(t_db, vap_pres)
    hi = t_db
    t_wb = (lo + hi) / 2
    it = 1
    while (hi - lo) > TOLERANCE_C:
        if hum_ratio_from_wet_bulb(t_db, t_wb, pressure) > w:
            hi = t_wb
        else:
            lo = t_wb
        t_wb = (hi + lo) / 2
        if it >= MAX_ITER:
            raise ValueError("wet bulb bisection did not converge")
        it += 1
    return t_wb


def wet_bulb_from_rel_hum(t_db, rel_hum, pressure):
    """T_wb (degC) from dry bulb (degC), relative humidity in [0, 1] and pressure (Pa)."""
    if rel_hum < 0 or rel_hum > 1:
        raise ValueError("relative humidity is outside range [0, 1]")
    w = hum_ratio_from_vap_pres(rel_hum * sat_vap_pres(t_db), pressure)
    return wet_bulb_from_hum_ratio(t_db, w, pressure)
