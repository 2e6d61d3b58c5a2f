mkdir -p gpurun_out/r4g
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bluestein" > gpurun_out/r4g/t.log 2>&1
