mkdir -p gpurun_out/r2h
python tools/ncu_round2.py > gpurun_out/r2h/plain.log 2>&1 && ncu --set full --clock-control none -k regex:'pipe_kernel|interp|invert_warp' -c 60 -o gpurun_out/r2h/round2 python tools/ncu_round2.py > gpurun_out/r2h/ncu.log 2>&1
ncu -i gpurun_out/r2h/round2.ncu-rep --page raw --csv > gpurun_out/r2h/round2_raw.csv 2>/dev/null
ls -la gpurun_out/r2h > gpurun_out/r2h/ls.txt
rm -f gpurun_out/r2h/round2.ncu-rep
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2h/bench_plain.json 2>gpurun_out/r2h/bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2h/bench_ncu.log 2>&1
python tools/bench_lengths.py > gpurun_out/r2h/bench_lengths.log 2>&1
cp gpurun_out/bench_lengths.json gpurun_out/r2h/bench_lengths.json
