mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_baseline.py -x -q -m gpu -k "sor or SOR or Sor or solve or golden or chained or rect" 2>&1 | tail -4
for n in 64 512 2048; do timeout 120 python profiles/sor_only.py $n 30; done
V=$PWD/picsp_b200/variants
for n in 64 512 2048; do PICSP_B200_LIB=$V/libpicsp_b200_head.so timeout 120 python profiles/sor_only.py $n 30 | sed 's/^/HEAD /'; done
