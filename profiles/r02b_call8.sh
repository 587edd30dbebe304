mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_own_fft.py -x -q -m gpu -s > gpurun_out/gputests_fft2.log 2>&1; echo "fft tests rc=$?"; grep -E "nodes:|passed|failed|Error|error" gpurun_out/gputests_fft2.log | head -40
timeout 600 python profiles/fft_only.py 64 128 256 512 1024 2048 2>&1 | tee gpurun_out/fft_only_v2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fft -c 4 -o gpurun_out/ncu_fft_v2 python profiles/fft_only.py 2048 > gpurun_out/ncu_fft_v2.log 2>&1; echo "ncu rc=$?"
