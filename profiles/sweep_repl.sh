#!/bin/bash
# profiles/sweep_repl.sh — plain window vs 4 bank-staggered copies (run under gpurun)
set -u
out=gpurun_out/sweep_repl.txt; : > $out
run() {
  PICSP_NVCC_DEFINES="$1" python -m picsp_b200.build --force > /dev/null 2>&1 || { echo "$1 BUILD FAILED" >> $out; return; }
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or midsize or sort_period or determin" 2>&1 | tail -1 >> $out
  python bench.py --no-cpu-baseline --no-e2e --particles 4e8 --steps 12 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'push %.3f sort %.3f ms' % (d['phases_ms_per_step']['push'], d['phases_ms_per_step']['sort']))" >> $out
}
run "-DPICSP_REPL=1"
run "-DPICSP_REPL=4"
run "-DPICSP_REPL=4 -DPICSP_HALO=4 -DPICSP_MOVER_THREADS=256 -DPICSP_MOVER_MIN_CTAS=3"
run "-DPICSP_REPL=4 -DPICSP_HALO=4"
run "-DPICSP_REPL=4 -DPICSP_MOVER_THREADS=1024 -DPICSP_MOVER_MIN_CTAS=1 -DPICSP_CHUNK=8192"
python -m picsp_b200.build --force > /dev/null 2>&1
cat $out
