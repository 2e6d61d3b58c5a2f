mkdir -p gpurun_out/r4m
python -m pytest tests -m gpu -q > gpurun_out/r4m/gputests.log 2>&1
