#!/bin/bash
# diagnostic visit: suspect test modules one by one, verbose, each under its own kill timer; logs survive a kill
OUT=gpurun_out
mkdir -p $OUT
run() { # name, seconds, command...
  name=$1; secs=$2; shift 2
  echo "== $name" | tee -a $OUT/diag.log
  timeout -s KILL $secs "$@" > $OUT/diag_$name.log 2>&1
  echo "rc=$? $(tail -1 $OUT/diag_$name.log | cut -c1-200)" | tee -a $OUT/diag.log
}
rm -f $OUT/diag*.log
export PYTHONUNBUFFERED=1
PPE_DEEP_WALKER=0 run expand_nodeep 400 python -m pytest tests/test_gpu_expand.py -m gpu -x -v --timeout 120
PPE_DEEP_WALKER=0 run harness_nodeep 500 python -m pytest tests/test_gpu_harness.py -m gpu -x -v --timeout 200
PPE_DEEP_WALKER=0 run plan_nodeep 700 python -m pytest tests/test_gpu_plan.py -m gpu -x -v --timeout 300
run parity_deep 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 300
for w in c2 c5; do python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-plan --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['thread_walked_edge_fraction'])" | tee -a $OUT/diag.log; done
