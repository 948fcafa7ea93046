#!/bin/bash
nr() { echo "== $*"; env "$@" python scripts/time_nr.py 10016 2>&1 | grep -E "batch S|rror" | sed 's/; status.*//'; }
nr JGB_BULK_LANES=4,4,8
nr JGB_BULK_LANES=4,8,8
nr JGB_BULK_LANES=8,8,8
nr JGB_BULK_LANES=4,8,16
nr JGB_BULK_LANES=4,4,16
