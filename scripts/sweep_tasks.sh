#!/bin/bash
# A/B of the batch factor phase (scripts/time_nr.py S): old per-front kernels, child-loop extend-add, task kernel variants
S=${1:-10016}
run() { echo "== $*"; env "$@" python scripts/time_nr.py $S 2>&1 | grep -E "batch S|check scenario|Error|error" ; }
run JGB_NO_TASKS=1 JGB_NO_CHILD_LOOP=1
run JGB_NO_TASKS=1
run JGB_X=0
run JGB_TASK_MAXNF=12
run JGB_TASK_MAXNF=8
run JGB_TASK_STACK=128
run JGB_TASK_STACK=400
run JGB_TASK_BUNDLE=8
run JGB_TASK_BUNDLE=48 JGB_TASK_META=8192
