#!/bin/bash
run() { echo "== relax $1"; JGB_RELAX="$1" python scripts/time_nr.py 1024 single 2>&1 | grep -E "batch S|single NR"; JGB_RELAX="$1" python -c "
import jgb200
ps=jgb200.synthetic_grid(); ctx=jgb200.Context(0); a=jgb200.newton_raphson(ps,ctx)
print({k:ctx.stat('nr.'+k) for k in ('fronts','levels','nnz_lu','flops','max_front','u_size','upd_size')})"; }
run "4,16,0.5,48,0.15,0.05"
run "0,0,0,0,0,0"
run "2,8,0.3,24,0.1,0.02"
run "4,8,0.3,16,0.1,0.0"
run "8,32,0.6,64,0.3,0.1"
