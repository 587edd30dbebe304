mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_own_fft.py -x -q -m gpu -s > gpurun_out/gputests_fft.log 2>&1; echo "fft tests rc=$?"; grep -E "nodes:|passed|failed|Error|error" gpurun_out/gputests_fft.log | head -40
