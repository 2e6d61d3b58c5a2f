mkdir -p gpurun_out/r3n
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bluestein or large_prime or legendre" > gpurun_out/r3n/t.log 2>&1
timeout 600 python tools/blue_probe.py > gpurun_out/r3n/blue.jsonl 2>&1
