mkdir -p gpurun_out/r3t
python -m pytest tests -m gpu -q > gpurun_out/r3t/gputests.log 2>&1
timeout 600 python tools/blue_probe.py > gpurun_out/r3t/blue.jsonl 2>&1
python tools/bench_lengths.py lattice > gpurun_out/r3t/lat.log 2>&1
