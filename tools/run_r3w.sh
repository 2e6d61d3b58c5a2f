mkdir -p gpurun_out/r3w
python -m pytest tests -m gpu -q -k "bluestein or large_prime or legendre or planned" > gpurun_out/r3w/gputests.log 2>&1
python tools/len_probe.py 9999,1000,25 8633,1400,25 7979,1300,25 5353,890,25 12345,2000,25 2>&1 | cut -c1-400 > gpurun_out/r3w/len.log
timeout 600 python tools/blue_probe.py > gpurun_out/r3w/blue.jsonl 2>&1
