(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_e.log 2>&1
python tools/bench_lengths.py > gpurun_out/bench_lengths3.log 2>&1
