mkdir -p gpurun_out
nvidia-smi --query-gpu=memory.total,memory.free --format=csv
timeout 900 python bench.py --cells 2048 --particles 4e9 --steps 16 --warmup 8 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "c5 n1 rc=$?"; tail -5 gpurun_out/bench_c5_n1.err; cat gpurun_out/bench_c5_n1.json | head -c 3000
