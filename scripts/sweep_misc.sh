#!/bin/bash
wls() { echo "== $*"; env "$@" python scripts/time_wls.py 1000 2>&1 | grep -E "batch WLS|rror"; }
wls JGB_BS_TS1_TILE=1
wls JGB_BS_TS1_TILE=0
wls JGB_BS_TS1_TILE=0 JGB_BS_THREADS=256
wls JGB_BS_TS1_TILE=0 JGB_BS_THREADS=128
