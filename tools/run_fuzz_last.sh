mkdir -p gpurun_out/fuzz_last
timeout 1000 python tools/fuzz_matsubara.py 300 2026 > gpurun_out/fuzz_last/matsubara.log 2>&1
timeout 500 python tools/fuzz_fft_many.py 150 2026 > gpurun_out/fuzz_last/fft.log 2>&1
