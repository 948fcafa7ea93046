#!/bin/bash
# A/B of the LDL^T big-front kernel with the FP64 tensor-core trailing update
#   JGB_DENSE_MIN    smallest front order it takes (0 = off; default 64 in batches, 16 for a single case)
#   JGB_DENSE_BUCKET launch buckets of this many rows (default 32)
#   JGB_DENSE_SMALL  fronts up to this order run 128-thread CTAs (default 96)
S=${1:-1000}
run() { echo "== $*"; env "$@" python scripts/time_wls.py $S 2>&1 | grep -E "single WLS|batch WLS|rror" ; }
run JGB_DENSE_MIN=0
run JGB_DENSE_BUCKET=1000 JGB_DENSE_SMALL=0
run JGB_DENSE_BUCKET=32 JGB_DENSE_SMALL=0
run JGB_DENSE_BUCKET=32 JGB_DENSE_SMALL=96
run JGB_DENSE_BUCKET=16 JGB_DENSE_SMALL=128
run JGB_DENSE_BUCKET=16 JGB_DENSE_SMALL=128 JGB_DENSE_MIN=48
