#!/bin/bash
# compute-sanitizer passes over the small parity cases (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards)
set -o pipefail
SEL='case14test or case30test or synthetic20 or all_codes or monte_carlo or outage_batch or islanding'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_nr_gpu.py tests/test_wls_gpu.py tests/test_batch_gpu.py -q -x -k "$SEL and not 10k and not ACTIVS" 2>&1 | tail -15
echo "memcheck rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_nr_gpu.py tests/test_batch_gpu.py tests/test_wls_gpu.py -q -x -k "(test_power_flow_golden or outage_batch or all_codes or monte_carlo) and not 10k" 2>&1 | tail -25
echo "racecheck rc=$?"
