#!/bin/bash
# A/B of the ring-mode children staging of the small-front (bulk) kernel
run() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 2>&1 | grep -E "batch S|check scen|rror"; env "$@" python scripts/time_wls.py 1000 2>&1 | grep -E "batch WLS|rror"; }
run JGB_BULK_RING=0
run JGB_BULK_RING=1
JGB_PLAN_DEBUG=1 python scripts/time_nr.py 10016 2>&1 | grep "plan S=10016" | grep bulk | sort | uniq -c | sort -k1,1nr | head -40
