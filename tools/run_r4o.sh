mkdir -p gpurun_out/r4o
python tools/len_probe.py 4999,833,1 4999,833,2 4999,833,4 19999,3333,1 5000,833,1 > gpurun_out/r4o/narrow_blue.log 2>&1
