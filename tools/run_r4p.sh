mkdir -p gpurun_out/r4p
python -m pytest tests -m gpu -q > gpurun_out/r4p/gputests.log 2>&1
python tools/len_probe.py 4999,833,1 4999,833,2 4999,833,4 19999,3333,1 4999,833,25 > gpurun_out/r4p/narrow_blue.log 2>&1
timeout 600 python tools/fuzz_matsubara.py 120 41 > gpurun_out/r4p/fuzz.log 2>&1
