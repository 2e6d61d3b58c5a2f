mkdir -p gpurun_out/r4e
python -m pytest tests -m gpu -q > gpurun_out/r4e/gputests.log 2>&1
python tools/real_probe.py > gpurun_out/r4e/real.log 2>&1
