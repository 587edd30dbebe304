#!/bin/bash
# profiles/sweep_mover.sh — rebuild the library with different mover tuning macros and bench each (run under gpurun).
set -u
out=gpurun_out/sweep_mover.txt; : > $out
run() {
  PICSP_NVCC_DEFINES="$1" python -m picsp_b200.build --force > /dev/null 2>&1 || { echo "$1 BUILD FAILED" >> $out; return; }
  python bench.py --no-cpu-baseline --no-e2e --particles 4e8 --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'push %.3f sort %.3f ms' % (d['phases_ms_per_step']['push'], d['phases_ms_per_step']['sort']))" >> $out
}
for cta in 3 4 5 6; do run "-DPICSP_MOVER_MIN_CTAS=$cta"; done
for chunk in 1024 4096; do run "-DPICSP_CHUNK=$chunk"; done
for halo in 2 3 6; do run "-DPICSP_HALO=$halo"; done
run "-DPICSP_CHUNK=4096 -DPICSP_MOVER_MIN_CTAS=5"
python -m picsp_b200.build --force > /dev/null 2>&1
cat $out
