#!/bin/bash
# single-case Newton-Raphson: the dense LU kernel (three-barrier panel + DMMA trailing update) against the scenario-tile
# kernel, staged (cp.async.bulk ring) against round-wise gather, chains of fronts walked by one CTA (JGB_SEQ = widest
# level that still chains, 0 = off), and the WLS single case with wider dense LDL^T CTAs
run() { echo "== $*"; env "$@" python scripts/time_nr.py 32 single 2>&1 | grep -E "single NR|rror"; }
wls() { echo "== $*"; env "$@" python scripts/time_wls.py 32 2>&1 | grep -E "single WLS|rror"; }
run JGB_DENSE_LU_MIN=0
run JGB_STAGED_EA=0
run JGB_STAGED_EA=1
run JGB_STAGED_EA=1 JGB_DENSE_LU_THREADS=512
run JGB_STAGED_EA=1 JGB_SEQ=4
wls JGB_X=0
wls JGB_DENSE_THREADS=512
wls JGB_DENSE_THREADS=512 JGB_DENSE_MIN=1
