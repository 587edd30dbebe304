source profiles/r02b_ab.sh true
V=$PWD/picsp_b200/variants
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gputests_c2.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/gputests_c2.log
run head $V/libpicsp_b200_head.so
run base ""
run hi $V/libpicsp_b200_hi.so
run near $V/libpicsp_b200_near.so
run both $V/libpicsp_b200_both.so
run base_nobank "" --bank-order-i 0
run head2 $V/libpicsp_b200_head.so
run base2 ""
run hi2 $V/libpicsp_b200_hi.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_mover -s 12 -c 2 -o gpurun_out/ncu_bank_c2 python bench.py --particles 2e8 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bank_c2.log 2>&1; echo "ncu rc=$?"
