#!/bin/bash
# Round 2 (second session) multi-GPU lines (run on an 8-GPU box through gpurun --gpus 8)
mkdir -p gpurun_out
run() {  # tag nproc args...
    tag=$1; n=$2; shift 2
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
        bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/r02b_mg_$tag.json 2> gpurun_out/r02b_mg_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02b_mg_{tag}.json"))
    p = d["phases_ms_per_step"]; e = d.get("e2e") or {}
    print(f"{tag:14s} N={d['n_gpus']} value {d['value']:.4g} ms/step {d['ms_per_step']:.3f} frac {d['roofline']['frac']:.3f} push {p['push']:.3f} "
          f"allreduce {p['allreduce']:.3f} solve {p['solve']:.3f} rho {p['rho']:.3f} ef {p['ef']:.3f} e2e {e.get('value', 0):.4g} steady {e.get('steady_state_value', 0):.4g} "
          f"probe {d['parity_probe']['sum_abs_rho']:.12g} {d['parity_probe']['l2_phi']:.12g} {d['parity_probe']['ke_e']:.12g}")
except Exception as ex:
    print(tag, "failed", ex)
PY
    grep -m1 "picsp_b200 error" gpurun_out/r02b_mg_$tag.err
}






nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/gputests_multi.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/gputests_multi.log
run n8 8 --steps 20 --warmup 4
run c5_n8 8 --cells 2048 --particles 4e9 --steps 10 --warmup 4 --no-e2e
run c5_n8_cufft 8 --cells 2048 --particles 4e9 --steps 10 --warmup 4 --no-e2e --flags 256
run n4 4 --steps 20 --warmup 4 --no-e2e
run c5_n4 4 --cells 2048 --particles 4e9 --steps 10 --warmup 4 --no-e2e
run n2 2 --steps 20 --warmup 4 --no-e2e
run c5_n2 2 --cells 2048 --particles 4e9 --steps 10 --warmup 4 --no-e2e
