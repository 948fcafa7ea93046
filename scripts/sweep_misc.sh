#!/bin/bash
nr() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 2>&1 | grep -E "batch S|rror" | sed 's/; status.*//'; }
wls() { echo "== $*"; env "$@" python scripts/time_wls.py 1000 2>&1 | grep -E "batch WLS|rror"; }
nr JGB_X=1
T="20:16:256,24:8:256,32:8:256,48:4:256,64:2:256,96:1:256,150:1:512,208:1:1024"
wls JGB_X=1
wls JGB_FPLAN_BATCH="6:32:128,8:32:128,10:32:128,12:32:128,16:32:256,$T"
