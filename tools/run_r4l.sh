mkdir -p gpurun_out/r4l
python -m pytest tests -m gpu -q > gpurun_out/r4l/gputests.log 2>&1
python tools/small_call_probe.py > gpurun_out/r4l/small.log 2>&1
