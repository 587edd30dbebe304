mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_case.py > gpurun_out/r02b_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02b_sanitize_$tool.log; grep -c "^r2b\|^fft" gpurun_out/r02b_sanitize_$tool.log
done
