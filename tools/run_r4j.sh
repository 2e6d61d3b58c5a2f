mkdir -p gpurun_out/r4j
python -m pytest tests -m gpu -q > gpurun_out/r4j/gputests.log 2>&1
python tools/small_call_probe.py > gpurun_out/r4j/small.log 2>&1
