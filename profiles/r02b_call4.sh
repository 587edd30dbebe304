mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parts.py -x -q -m gpu > gpurun_out/gputests_parts.log 2>&1; echo "parts tests rc=$?"; tail -30 gpurun_out/gputests_parts.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gputests_c4.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/gputests_c4.log
