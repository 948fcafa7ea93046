#!/bin/bash
S=${1:-10016}
run() { echo "== $*"; env "$@" python scripts/time_nr.py $S single 2>&1 | grep -E "batch S|single NR|check scenario|rror" ; }
run JGB_NO_ASYNC_EA=1
run JGB_X=0
run JGB_TASKS=1
