#!/bin/bash
# Round 2 same-box A/B of the mover variants (run on the GPU box through gpurun): rebuilds the library with different
# -D switches and prints per-species mover launch times of the bench workload.
#   AGG0      no warp aggregation (PICSP_AGG_ROUNDS=0): the round-1 deposit
#   default   MATCH + REDUX aggregation, ions cell-ordered every 64 steps
#   nocell    aggregation compiled in, no cell ordering
set -u
mkdir -p gpurun_out
run() {  # tag, defines, bench args...
    tag=$1; defs=$2; shift 2
    PICSP_NVCC_DEFINES="$defs" python -m picsp_b200.build --force > /dev/null 2>&1 || { echo "$tag: build failed"; return; }
    timeout 300 python bench.py --steps 24 --warmup 8 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{tag}.json"))
    p = d["phases_ms_per_step"]
    print(f"{tag:12s} value {d['value']:.4g} ms/step {d['ms_per_step']:.3f} frac {d['roofline']['frac']:.3f} push_i {p.get('push_ions', 0):.3f} "
          f"push_e {p.get('push_electrons', 0):.3f} sort {p['sort']:.3f} sm {d['clocks']['sm_mhz']}")
except Exception as e:
    print(tag, "failed", e)
PY
}
run AGG0 "-DPICSP_AGG_ROUNDS=0" --cell-period-i 0
run default ""
run nocell "" --cell-period-i 0
run cell_e2 "" --cell-period-e 2
run AGG0b "-DPICSP_AGG_ROUNDS=0" --cell-period-i 0
python -m picsp_b200.build --force > /dev/null 2>&1
