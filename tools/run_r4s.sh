mkdir -p gpurun_out/r4s
python -m pytest tests -m gpu -q > gpurun_out/r4s/gputests.log 2>&1
python tools/len_probe.py 9999,1000,25 19999,3333,25 4999,833,9 12345,2000,25 > gpurun_out/r4s/len.log 2>&1
timeout 500 python tools/fuzz_matsubara.py 100 91 > gpurun_out/r4s/fuzz.log 2>&1
