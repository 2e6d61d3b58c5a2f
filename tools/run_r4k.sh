mkdir -p gpurun_out/r4k
python -m pytest tests -m gpu -q > gpurun_out/r4k/gputests.log 2>&1
python bench.py --no-extras > gpurun_out/r4k/bench.json 2> gpurun_out/r4k/bench.err
