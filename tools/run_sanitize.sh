mkdir -p gpurun_out/san
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gf_mirror.py -m gpu -q -x -k "bluestein or real_valued or lazy_descriptors or large_prime" > gpurun_out/san/memcheck.log 2>&1
echo "exit $?" >> gpurun_out/san/memcheck.log
