mkdir -p gpurun_out/r3r
python -m pytest tests -m gpu -q > gpurun_out/r3r/gputests.log 2>&1
python tools/len_probe.py 6150,1025,25 61500,10250,25 7800,1300,25 > gpurun_out/r3r/len.log 2>&1
