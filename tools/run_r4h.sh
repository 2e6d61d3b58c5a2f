mkdir -p gpurun_out/r4h
python -m pytest tests -m gpu -q > gpurun_out/r4h/gputests.log 2>&1
