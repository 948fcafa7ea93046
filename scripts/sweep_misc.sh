#!/bin/bash
wls() { echo "== $*"; env "$@" python scripts/time_wls.py 1000 2>&1 | grep -E "single WLS|batch WLS|rror"; }
wls JGB_DENSE_THREADS=256
wls JGB_DENSE_THREADS=512
wls JGB_DENSE_THREADS=1024
wls JGB_DENSE_THREADS=768
