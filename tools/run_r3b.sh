mkdir -p gpurun_out/r3b
python -m pytest tests -m gpu -q > gpurun_out/r3b/gputests.log 2>&1
