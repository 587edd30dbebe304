mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_own_fft.py -x -q -m gpu -s > gpurun_out/gputests_fft3.log 2>&1; echo "fft tests rc=$?"; grep -E "nodes:|passed|failed|Error|error" gpurun_out/gputests_fft3.log | head -40
timeout 600 python profiles/fft_only.py 32 64 128 256 512 1024 2048 2>&1 | tee gpurun_out/fft_only_v3.log
