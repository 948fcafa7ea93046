#!/bin/bash
# compute-sanitizer passes over the small parity cases (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards)
set -o pipefail
SEL='case14test or case30test or synthetic20 or all_codes or monte_carlo or outage_batch or islanding'
NEW='test_dc_power_flow_goldens or test_dc_state_estimation_recovers or test_pmu_state_estimation_recovers or test_linear_solver_refactor or test_one_outlier or test_two_outliers or test_rectangular_pmu'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_nr_gpu.py tests/test_wls_gpu.py tests/test_batch_gpu.py -q -x -k "$SEL and not 10k and not ACTIVS" 2>&1 | tail -15
echo "memcheck rc=$?"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_linear_gpu.py tests/test_baddata_gpu.py -q -x -k "$NEW" 2>&1 | tail -15
echo "memcheck (linear solves, bad data) rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_nr_gpu.py tests/test_batch_gpu.py tests/test_wls_gpu.py -q -x -k "(test_power_flow_golden or outage_batch or all_codes or monte_carlo) and not 10k" 2>&1 | tail -25
echo "racecheck rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_linear_gpu.py tests/test_baddata_gpu.py -q -x -k "$NEW" 2>&1 | tail -25
echo "racecheck (linear solves, bad data) rc=$?"
