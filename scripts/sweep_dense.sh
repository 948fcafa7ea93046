#!/bin/bash
# A/B of the LDL^T big-front kernel with the FP64 tensor-core trailing update (JGB_DENSE_MIN = smallest front order it takes, 0 = off)
S=${1:-1000}
run() { echo "== $*"; env "$@" python scripts/time_wls.py $S 2>&1 | grep -E "single WLS|batch WLS|rror" ; }
run JGB_DENSE_MIN=0
run JGB_DEFAULT=1
run JGB_DENSE_MIN=16
run JGB_DENSE_MIN=32
run JGB_DENSE_MIN=48
