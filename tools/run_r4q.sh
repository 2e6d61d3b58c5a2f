mkdir -p gpurun_out/r4q
python tools/len_probe.py 9999,1000,1 9999,1000,2 8633,1400,1 > gpurun_out/r4q/len.log 2>&1
python -m pytest tests -m gpu -q -k "planned or narrow or bluestein or randomised" > gpurun_out/r4q/t.log 2>&1
