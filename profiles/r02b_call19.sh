mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gputests_c19.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gputests_c19.log
timeout 600 python profiles/fft_only.py 32 64 128 256 512 1024 2048 2>&1 | tee gpurun_out/fft_only_v4.log
timeout 900 python bench.py --steps 20 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/bench_1gpu_c19.json 2> gpurun_out/bench_1gpu_c19.err; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_c19.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['phases_ms_per_step'], d['config']['solver'], d['parity_probe'])"
