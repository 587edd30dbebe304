#!/bin/bash
# Round 2 (second session), after the own DFT took over the 1025^2 solve: the bench configuration on 8 / 4 / 2 GPUs again
mkdir -p gpurun_out
run() {  # tag nproc args...
    tag=$1; n=$2; shift 2
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
        bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/r02b_mg2_$tag.json 2> gpurun_out/r02b_mg2_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02b_mg2_{tag}.json"))
    p = d["phases_ms_per_step"]
    print(f"{tag:10s} N={d['n_gpus']} value {d['value']:.4g} ms/step {d['ms_per_step']:.3f} frac {d['roofline']['frac']:.3f} push {p['push']:.3f} "
          f"allreduce {p['allreduce']:.3f} solve {p['solve']:.3f} rho {p['rho']:.3f} ef {p['ef']:.3f} "
          f"probe {d['parity_probe']['sum_abs_rho']:.12g} {d['parity_probe']['l2_phi']:.12g} {d['parity_probe']['ke_e']:.12g}")
except Exception as ex:
    print(tag, "failed", ex)
PY
}
run n8 8 --steps 20 --warmup 4 --no-e2e
run n8_cufft 8 --steps 20 --warmup 4 --no-e2e --flags 256
run n4 4 --steps 20 --warmup 4 --no-e2e
run n2 2 --steps 20 --warmup 4 --no-e2e
run n1 1 --steps 20 --warmup 4 --no-e2e
true
