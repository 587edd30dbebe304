#!/bin/bash
set -u
out=gpurun_out/sweep_pipe3.txt; : > $out
run() {
  PICSP_NVCC_DEFINES="-DPICSP_BULK_PIPE=1 $1" python -m picsp_b200.build --force --verbose 2>&1 | grep -A2 "k_tile_moverILi0" | grep -E "Used" | sed 's/ptxas info    : //' | tr '\n' ' ' >> $out
  python bench.py --no-cpu-baseline --no-e2e --particles 4e8 --steps 12 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'push %.3f sort %.3f ms' % (d['phases_ms_per_step']['push'], d['phases_ms_per_step']['sort']))" >> $out
}
B="-DPICSP_MOVER_THREADS=128 -DPICSP_STAGES=2"
run "$B -DPICSP_HALO=4 -DPICSP_MOVER_MIN_CTAS=8 -DPICSP_CHUNK=2048"
run "$B -DPICSP_HALO=4 -DPICSP_MOVER_MIN_CTAS=8 -DPICSP_CHUNK=8192"
run "$B -DPICSP_HALO=5 -DPICSP_MOVER_MIN_CTAS=8"
run "$B -DPICSP_HALO=4 -DPICSP_MOVER_MIN_CTAS=9"
run "$B -DPICSP_HALO=3 -DPICSP_MOVER_MIN_CTAS=8"
run "-DPICSP_MOVER_THREADS=96 -DPICSP_STAGES=2 -DPICSP_HALO=4 -DPICSP_MOVER_MIN_CTAS=10"
run "-DPICSP_MOVER_THREADS=64 -DPICSP_STAGES=2 -DPICSP_HALO=4 -DPICSP_MOVER_MIN_CTAS=11"
python -m picsp_b200.build --force > /dev/null 2>&1
cat $out
