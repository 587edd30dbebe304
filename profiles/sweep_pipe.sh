#!/bin/bash
# profiles/sweep_pipe.sh — register prefetch vs TMA bulk-copy particle pipeline (run under gpurun)
set -u
out=gpurun_out/sweep_pipe.txt; : > $out
run() {
  PICSP_NVCC_DEFINES="$1" python -m picsp_b200.build --force > /dev/null 2>&1 || { echo "$1 BUILD FAILED" >> $out; return; }
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or midsize or sort_period or determin or tiny or rect" 2>&1 | tail -1 >> $out
  python bench.py --no-cpu-baseline --no-e2e --particles 4e8 --steps 12 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'push %.3f sort %.3f ms' % (d['phases_ms_per_step']['push'], d['phases_ms_per_step']['sort']))" >> $out
}
run "-DPICSP_BULK_PIPE=0"
run "-DPICSP_BULK_PIPE=1 -DPICSP_STAGES=3"
run "-DPICSP_BULK_PIPE=1 -DPICSP_STAGES=4"
run "-DPICSP_BULK_PIPE=1 -DPICSP_STAGES=2"
run "-DPICSP_BULK_PIPE=1 -DPICSP_STAGES=4 -DPICSP_MOVER_MIN_CTAS=3"
run "-DPICSP_BULK_PIPE=1 -DPICSP_STAGES=3 -DPICSP_REPL=4"
python -m picsp_b200.build --force > /dev/null 2>&1
cat $out
