mkdir -p gpurun_out/r3g
python -m pytest tests -m gpu -q -k "cpp" > gpurun_out/r3g/gputests.log 2>&1
