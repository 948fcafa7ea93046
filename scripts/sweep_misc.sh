#!/bin/bash
nr() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 2>&1 | grep -E "batch S|check scen|rror" | sed 's/; status.*//'; }
nr JGB_X=1
nr JGB200_LIB=$PWD/build/alt/libjgb200_mb4.so
