#!/bin/bash
# same-box A/B of library builds: bash profiles/ab_variants.sh "<label>:<lib>:<flags>" ...   (lib '' = in-tree build)
run() {
  python - "$2" "$3" <<'PY' > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
import sys, runpy
import picsp_b200.lib as l
if sys.argv[1]: l.LIB_PATH = sys.argv[1]
flags = sys.argv[2]
sys.argv = ["bench.py", "--no-e2e", "--no-cpu-baseline", "--steps", "24", "--warmup", "3", "--flags", flags]
runpy.run_path("bench.py", run_name="__main__")
PY
  python -c "
import json;d=json.load(open('gpurun_out/ab_tmp.json'));print('$1'.ljust(24), '%.4g'%d['value'], round(d['ms_per_step'],3), 'push', round(d['phases_ms_per_step']['push'],3), 'sort', round(d['phases_ms_per_step']['sort'],3), d['clocks']['sm_mhz'])"
}
for rep in 1 2; do
  for spec in "$@"; do
    IFS=: read -r label lib flags <<< "$spec"
    [ -n "$lib" ] && lib="$PWD/$lib"
    run "$label" "$lib" "$flags"
  done
done
