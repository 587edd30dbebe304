mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gputests_c10.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gputests_c10.log
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_1gpu_c10.json 2> gpurun_out/bench_1gpu_c10.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_1gpu_c10.err; head -c 2500 gpurun_out/bench_1gpu_c10.json
timeout 600 python bench.py --cells 2048 --particles 4e9 --steps 10 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_n1_v2.json 2> gpurun_out/bench_c5_n1_v2.err; echo "c5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_n1_v2.json')); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['parity_probe'], d['clocks'])"
