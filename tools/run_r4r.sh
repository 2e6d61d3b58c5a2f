mkdir -p gpurun_out/r4r
python -m pytest tests -m gpu -q > gpurun_out/r4r/gputests.log 2>&1
python tools/len_probe.py 4999,833,1 19999,3333,1 9999,1000,1 > gpurun_out/r4r/len.log 2>&1
timeout 500 python tools/fuzz_matsubara.py 100 77 > gpurun_out/r4r/fuzz.log 2>&1
