mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gputests_c16.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/gputests_c16.log
