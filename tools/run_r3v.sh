mkdir -p gpurun_out/r3v
python -m pytest tests -m gpu -q > gpurun_out/r3v/gputests.log 2>&1
python tools/len_probe.py 9999,1000,25 8633,1400,25 7979,1300,25 5353,890,25 6150,1025,25 9999,1000,2 2>&1 | cut -c1-150 > gpurun_out/r3v/len.log
timeout 500 python tools/fuzz_matsubara.py 80 11 > gpurun_out/r3v/fuzz.log 2>&1
