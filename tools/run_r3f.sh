mkdir -p gpurun_out/r3f
python -m pytest tests -m gpu -q > gpurun_out/r3f/gputests.log 2>&1
python bench.py > gpurun_out/r3f/bench.json 2> gpurun_out/r3f/bench.err
