source profiles/r02b_ab.sh true
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gputests_c1.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/gputests_c1.log
run head $PWD/picsp_b200/variants/libpicsp_b200_head.so
run cur ""
run nobank "" --bank-order-i 0
run banke "" --bank-order-e 1
run head2 $PWD/picsp_b200/variants/libpicsp_b200_head.so
run cur2 ""
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_mover -s 12 -c 2 -o gpurun_out/ncu_bank_c1 python bench.py --particles 2e8 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bank_c1.log 2>&1; echo "ncu rc=$?"
