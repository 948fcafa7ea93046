#!/bin/bash
# A/B of the DAG schedule of the batch factor phase (one stream per launch class) against one stream
run() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 2>&1 | grep -E "batch S|check scen|rror"; env "$@" python scripts/time_wls.py 1000 2>&1 | grep -E "batch WLS|rror"; }
run JGB_LANES=1
run JGB_LANES=0
run JGB_LANES=2
run JGB_LANES=4
