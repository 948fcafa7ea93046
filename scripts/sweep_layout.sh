#!/bin/bash
# timings of the NR and WLS paths (single case + batch), used for the update-storage layout A/B of round 2

run() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 single 2>&1 | grep -E "batch S|single NR|rror"; env "$@" python scripts/time_wls.py 1000 2>&1 | grep -E "single WLS|batch WLS|rror"; }

run JGB_DEFAULT=1
