#!/bin/bash
wls() { echo "== $*"; env "$@" python scripts/time_wls.py 1000 2>&1 | grep -E "single WLS|batch WLS|rror"; }
wls JGB_X=1
