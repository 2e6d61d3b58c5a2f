mkdir -p gpurun_out/r3m
python bench.py > gpurun_out/r3m/bench_n1.json 2> gpurun_out/r3m/bench_n1.err
python bench.py --impl reference > gpurun_out/r3m/bench_ref.json 2> gpurun_out/r3m/bench_ref.err
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r3m/bench_plain.json 2>gpurun_out/r3m/bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r3m/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r3m/bench_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3m/smoke.log 2>&1
