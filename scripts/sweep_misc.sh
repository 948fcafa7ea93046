#!/bin/bash
b() { echo "== $*"; env "$@" python bench.py --workload wls --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:round(v['ms'],2) for k,v in d['roofline']['per_phase'].items()})
"; }
b JGB_LANES=1
b JGB_LANES=0
b JGB_LANES=1 JGB_BULK_RING=0
