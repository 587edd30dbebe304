#!/bin/bash
set -u
out=gpurun_out/sweep_mover2.txt; : > $out
run() {
  PICSP_NVCC_DEFINES="$1" python -m picsp_b200.build --force > /dev/null 2>&1 || { echo "$1 BUILD FAILED" >> $out; return; }
  python bench.py --no-cpu-baseline --no-e2e --particles 4e8 --steps 16 --warmup 3 $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2', 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'push %.3f sort %.3f ms' % (d['phases_ms_per_step']['push'], d['phases_ms_per_step']['sort']))" >> $out
}
run "-DPICSP_CHUNK=4096" ""
run "-DPICSP_CHUNK=8192" ""
run "-DPICSP_CHUNK=16384" ""
run "-DPICSP_CHUNK=4096 -DPICSP_HALO=6" "--sort-period-e 16"
run "-DPICSP_CHUNK=8192 -DPICSP_HALO=6" "--sort-period-e 16"
run "-DPICSP_CHUNK=8192 -DPICSP_HALO=8" "--sort-period-e 24"
run "-DPICSP_CHUNK=8192 -DPICSP_MOVER_THREADS=512 -DPICSP_MOVER_MIN_CTAS=2" ""
run "-DPICSP_CHUNK=4096 -DPICSP_MOVER_THREADS=128 -DPICSP_MOVER_MIN_CTAS=8" ""
python -m picsp_b200.build --force > /dev/null 2>&1
cat $out
