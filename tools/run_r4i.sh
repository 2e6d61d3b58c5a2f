mkdir -p gpurun_out/r4i
timeout 900 python tools/fuzz_fft_many.py 200 3 > gpurun_out/r4i/fuzz_fft.log 2>&1
