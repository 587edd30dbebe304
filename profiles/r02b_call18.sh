V=$PWD/picsp_b200/variants
for v in kp3 kp2 kp3t224 kp3t256 kp2t256; do echo $v; PICSP_B200_LIB=$V/libpicsp_b200_$v.so timeout 200 python profiles/fft_only.py 512 1024 | tail -2; done
