"""jgb200 — B200-native Newton-Raphson power flow and Gauss-Newton WLS state estimation behind JuliaGrid's
operator surface. Host mirror in Python (Julia is not available in this image; `julia/JuliaGridB200.jl` is the
ccall shim a JuliaGrid user loads), numerics in hand-written sm_100a CUDA behind the C ABI of include/jgb200.h.
"""
from ._lib import Context, JgbError, load, LIB_PATH, exported_symbols  # noqa: F401
from .cases import PowerSystem, power_system, synthetic_grid  # noqa: F401
from .model import AcModel, ac_model, apply_branch_status  # noqa: F401
from .ac_power_flow import (AcPowerFlow, newton_raphson, mismatch, solve, power_flow, set_initial_point,  # noqa: F401
                            set_voltage, update_branch, power_device, newtonRaphson, powerFlow, setInitialPoint, updateBranch,
                            generator_power, reactive_limit, adjust_angle, generatorPower, reactiveLimit, adjustAngle)
from .measurement import (Measurement, measurement, power, add_voltmeter, add_ammeter, add_wattmeter,  # noqa: F401
                          add_varmeter, add_pmu, ac_wls, WlsTables, load_measurement, measurement_from_arrays,
                          measurement_to_arrays)
from .ac_state_estimation import (AcStateEstimation, gauss_newton, increment, solve_se, state_estimation,  # noqa: F401
                                  set_mean, set_voltage_se, gaussNewton, stateEstimation, chi_test,
                                  residual_test, ChiTest, ResidualTest, chiTest, residualTest, update_voltmeter,
                                  update_ammeter, update_wattmeter, update_varmeter, update_pmu, update_branch_se)
from .batch import BatchResult, eligible_outages, outage_arrays, nr_batch, wls_batch  # noqa: F401
from .linear_solver import LinearSolver  # noqa: F401
from .dc_power_flow import DcModel, DcPowerFlow, dc_model, dc_power_flow, solve_dc, dc_batch, power_dc  # noqa: F401
from .dc_state_estimation import (DcStateEstimation, LinearWls, dc_wls_tables, dc_state_estimation,  # noqa: F401
                                  solve_dc_se, dc_se_batch)
from .pmu_state_estimation import (PmuStateEstimation, pmu_wls_tables, pmu_state_estimation, solve_pmu_se,  # noqa: F401
                                   pmu_se_batch)
from .fast_newton_raphson import (AcPowerFlowFast, fast_jacobians, fast_newton_raphson_bx,  # noqa: F401
                                  fast_newton_raphson_xb, mismatch_fnr, solve_fnr, power_flow_fnr, fnr_batch,
                                  fastNewtonRaphsonBX, fastNewtonRaphsonXB)
from . import dist  # noqa: F401
