mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_baseline.py -x -q -m gpu -k "sor or SOR or Sor or solve or golden or chained or rect or baseline" 2>&1 | tail -4
for n in 128 256 512 1024 2048; do timeout 120 python profiles/sor_only.py $n 30; done
