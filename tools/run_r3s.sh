mkdir -p gpurun_out/r3s
timeout 800 python tools/fuzz_matsubara.py 150 7 > gpurun_out/r3s/fuzz.log 2>&1
