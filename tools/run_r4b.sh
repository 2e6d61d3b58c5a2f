mkdir -p gpurun_out/r4b
timeout 900 python tools/fuzz_matsubara.py 200 23 > gpurun_out/r4b/fuzz.log 2>&1
