#!/bin/bash
run() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 2>&1 | grep -E "batch S|check scen|rror"; }
run JGB_BULK_RING=1 JGB_STAGED_EA=1
run JGB_BULK_RING=1 JGB_STAGED_EA=0
