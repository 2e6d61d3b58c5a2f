mkdir -p gpurun_out/r3o
python -m pytest tests -m gpu -q > gpurun_out/r3o/gputests.log 2>&1
