#!/bin/bash
# electron re-binning cadence with the re-binning mover (same box, back to back)
for rep in 1 2; do
for pe in 4 6 8 12 16 24; do
  python bench.py --no-e2e --no-cpu-baseline --steps 48 --warmup 3 --sort-period-e $pe > gpurun_out/sw_tmp.json 2> gpurun_out/sw_tmp.err
  python -c "
import json;d=json.load(open('gpurun_out/sw_tmp.json'));print('period_e', $pe, '%.4g'%d['value'], round(d['ms_per_step'],3), 'push', round(d['phases_ms_per_step']['push'],3), d['clocks']['sm_mhz'])"
done; done
