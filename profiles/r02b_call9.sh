mkdir -p gpurun_out
for t in 128 192 256 384 512; do echo "threads $t"; PICSP_FFT_THREADS=$t timeout 300 python profiles/fft_only.py 256 2048 2>&1 | tail -2; done
echo default; timeout 300 python profiles/fft_only.py 32 64 128 256 2048 2>&1 | tail -5
