source profiles/r02b_ab.sh true
V=$PWD/picsp_b200/variants
run head $V/libpicsp_b200_head.so
run p25_nobank $V/libpicsp_b200_p25.so --bank-order-i 0
run p26_nobank "" --bank-order-i 0
run p25_bank $V/libpicsp_b200_p25.so
run p26_bank ""
run head2 $V/libpicsp_b200_head.so
run p25_nobank2 $V/libpicsp_b200_p25.so --bank-order-i 0
run p26_nobank2 "" --bank-order-i 0
run p25_bank2 $V/libpicsp_b200_p25.so
run p26_bank2 ""
# decay of the order: the same 24 timed steps after 60 warm-up steps (order established at upload, 60-84 steps old)
run head_late $V/libpicsp_b200_head.so --warmup 60
run p26_bank_late "" --warmup 60
run p25_bank_late $V/libpicsp_b200_p25.so --warmup 60
