#!/bin/bash
# batch preset: supernode amalgamation against the current kernels (JGB_RELAX = small,mid,midfrac,big,bigfrac,anyfrac)
nr() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 2>&1 | grep -E "batch S|rror" | sed 's/; status.*//'; }
wls() { env "$@" python scripts/time_wls.py 1000 2>&1 | grep -E "batch WLS|rror"; }
nr JGB_RELAX=2,6,0.2,16,0.05,0.01
wls JGB_RELAX=2,6,0.2,16,0.05,0.01
nr JGB_RELAX=2,6,0.2,24,0.05,0.01
wls JGB_RELAX=2,6,0.2,24,0.05,0.01
nr JGB_RELAX=2,8,0.2,24,0.05,0.01
nr JGB_RELAX=2,6,0.25,16,0.08,0.01
nr JGB_RELAX=2,6,0.15,16,0.04,0.01
nr JGB_RELAX=1,6,0.2,16,0.05,0.01
wls JGB_RELAX=4,8,0.3,24,0.1,0.02
