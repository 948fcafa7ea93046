#!/bin/bash
# tuning sweep of the supernode amalgamation for batches (JGB_RELAX overrides both presets)
run() { echo "== relax $1"; JGB_RELAX="$1" python scripts/time_nr.py 2048 2>&1 | grep -E "batch S"; }
run "2,8,0.3,24,0.1,0.02"
run "3,8,0.3,24,0.1,0.02"
run "4,8,0.3,24,0.1,0.02"
run "2,6,0.2,16,0.08,0.02"
run "2,12,0.4,32,0.12,0.03"
run "1,4,0.2,16,0.05,0.0"
run "4,12,0.4,32,0.15,0.03"
