mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gputests_c20.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gputests_c20.log
for n in 32 64 100; do timeout 120 python profiles/sor_only.py $n 50; PICSP_SOR_NO_SMEM=1 timeout 120 python profiles/sor_only.py $n 50 | sed 's/^/pipelined: /'; done
timeout 600 python profiles/configs_bench.py 2>/dev/null | head -1 | cut -c1-400
time ./picsp_b200/picsp_b200_run tests/golden/input_ini_shipped.ini --out /tmp/data.h5 > /tmp/run.log 2>&1; tail -2 /tmp/run.log
