mkdir -p gpurun_out/r3l
python -m pytest tests -m gpu -q > gpurun_out/r3l/gputests.log 2>&1
python tools/narrow_probe.py > gpurun_out/r3l/narrow.jsonl 2>&1
