mkdir -p gpurun_out/r3e
python -m pytest tests -m gpu -q -k "real_valued" > gpurun_out/r3e/gputests.log 2>&1
