mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fft -c 4 -o gpurun_out/ncu_fft_1025 python profiles/fft_only.py 1024 > gpurun_out/ncu_fft_1025.log 2>&1; echo "ncu rc=$?"
