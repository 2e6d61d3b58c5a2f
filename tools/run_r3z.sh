mkdir -p gpurun_out/r3z
python -m pytest tests -m gpu -q > gpurun_out/r3z/gputests.log 2>&1
python tools/len_probe.py 6150,1025,25 61500,10250,25 6600,1100,25 7800,1300,25 10200,1700,25 6150,1025,4 2>&1 | cut -c1-330 > gpurun_out/r3z/len.log
