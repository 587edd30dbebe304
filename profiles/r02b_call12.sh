mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_mover -s 12 -c 4 -o gpurun_out/r02b_mover_final python bench.py --particles 2e8 --sort-period-e 2 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_mover_final.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_final.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_launches_final.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:k_bank_order -c 1 -o gpurun_out/r02b_bank_order python bench.py --particles 2e8 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02b_bank_order.log 2>&1; echo "ncu bank rc=$?"
timeout 600 python profiles/configs_bench.py > gpurun_out/r02b_configs.log 2>&1; echo "configs rc=$?"; tail -12 gpurun_out/r02b_configs.log
