#!/bin/bash
# A/B of the staged (cp.async.bulk ring) extend-add of the scenario-tile LU kernel against the gather in rounds
run() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 2>&1 | grep -E "batch S|check scen|rror"; }
run JGB_STAGED_EA=0
run JGB_STAGED_EA=1
run JGB_STAGED_EA=0
run JGB_STAGED_EA=1
