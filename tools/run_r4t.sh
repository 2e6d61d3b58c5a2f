mkdir -p gpurun_out/r4t
python tools/len_probe.py 4999,833,9 4999,833,8 4999,833,12 4999,833,16 > gpurun_out/r4t/len2.log 2>&1
python -m pytest tests -m gpu -q -k "bluestein or large_prime or randomised or narrow" > gpurun_out/r4t/t.log 2>&1
