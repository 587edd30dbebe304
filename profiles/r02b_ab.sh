#!/bin/bash
# Round 2 (second session) same-box A/B on the GPU box: bank order inside the chunks + the mover's instruction diet.
#   variants are pre-built here (profiles/build_variant.sh) and selected with PICSP_B200_LIB
set -u
mkdir -p gpurun_out
run() {  # tag, lib ("" = in-tree), bench args...
    tag=$1; lib=$2; shift 2
    PICSP_B200_LIB=$lib timeout 300 python bench.py --steps 24 --warmup 8 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ab2_$tag.json 2> gpurun_out/ab2_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab2_{tag}.json"))
    p = d["phases_ms_per_step"]
    print(f"{tag:12s} value {d['value']:.4g} ms/step {d['ms_per_step']:.3f} frac {d['roofline']['frac']:.3f} push_i {p.get('push_ions', 0):.3f} "
          f"push_e {p.get('push_electrons', 0):.3f} sort {p['sort']:.3f} sm {d['clocks']['sm_mhz']} probe {d.get('parity_probe')}")
except Exception as e:
    print(tag, "failed", e)
PY
}
"$@"
