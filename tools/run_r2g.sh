mkdir -p gpurun_out/r2g
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2g/pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g/smoke.log 2>&1
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r2g/bench_ref.json 2> gpurun_out/r2g/bench_ref.err
