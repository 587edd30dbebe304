#!/bin/bash
set -u
out=gpurun_out/sweep_pipe5.txt; : > $out
run() {
  PICSP_NVCC_DEFINES="$1" python -m picsp_b200.build --force --verbose 2>&1 | grep -A2 "k_tile_moverILi0" | grep -E "Used" | sed 's/ptxas info    : //' | tr '\n' ' ' >> $out
  python bench.py --no-cpu-baseline --no-e2e --particles 4e8 --steps 12 --warmup 3 $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2', 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'push %.3f sort %.3f ms' % (d['phases_ms_per_step']['push'], d['phases_ms_per_step']['sort']))" >> $out
}
run "-DPICSP_BULK_PIPE=0" ""
run "-DPICSP_BULK_PIPE=0 -DPICSP_MOVER_MIN_CTAS=7" ""
run "-DPICSP_BULK_PIPE=1" ""
python -m picsp_b200.build --force > /dev/null 2>&1
cat $out
