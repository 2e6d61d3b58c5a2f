mkdir -p gpurun_out/r4u
python -m pytest tests -m gpu -q > gpurun_out/r4u/gputests.log 2>&1
