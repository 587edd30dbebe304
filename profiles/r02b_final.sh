#!/bin/bash
# Final verification of the round's last build (one GPU): whole GPU suite, smoke, sanitizer, default bench line,
# BASELINE config 5 on this one GPU with the default flags (the e2e leg skips itself: 128 GB of host buffers)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gputests_final.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gputests_final.log
python __graft_entry__.py smoke 2>&1 | tail -2
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_case.py > gpurun_out/r02b_final_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02b_final_sanitize_$tool.log
done
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d['phases_ms_per_step']['solve'])"
timeout 900 python bench.py --cells 2048 --particles 4e9 --no-cpu-baseline > gpurun_out/bench_c5_n1_final.json 2> gpurun_out/bench_c5_n1_final.err; echo "c5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_n1_final.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['e2e_skipped'], d['store_parts_per_species'], d['parity_probe']['sum_abs_rho'])"
