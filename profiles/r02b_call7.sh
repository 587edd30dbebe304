mkdir -p gpurun_out
timeout 600 python profiles/fft_only.py 64 128 256 512 1024 2048 4096 2>&1 | tee gpurun_out/fft_only_v1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fft -c 4 -o gpurun_out/ncu_fft_v1 python profiles/fft_only.py 2048 > gpurun_out/ncu_fft_v1.log 2>&1; echo "ncu rc=$?"
