#!/bin/bash
# single case with the dense kernels: supernode amalgamation (JGB_RELAX = small,mid,midfrac,big,bigfrac,anyfrac) and ordering
run() { echo "== $*"; env "$@" python scripts/time_nr.py 32 single 2>&1 | grep -E "single NR|rror"; env "$@" python scripts/time_wls.py 32 2>&1 | grep -E "single WLS|rror"; }
run JGB_X=1
run JGB_RELAX=4,16,0.5,96,0.3,0.05
run JGB_RELAX=4,24,0.5,128,0.5,0.1
run JGB_RELAX=8,24,0.6,128,0.4,0.1
run JGB_RELAX=6,32,0.6,150,0.5,0.15
run JGB_RELAX=4,16,0.5,64,0.25,0.05
