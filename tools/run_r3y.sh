mkdir -p gpurun_out/r3y
python tools/ncu_round3.py > gpurun_out/r3y/plain.log 2>&1 && ncu --set full --clock-control none -k regex:'direct_pre|direct_post|lattice_pipe|promote|real_part|pipe_kernel' -c 80 -o gpurun_out/r3y/round3 python tools/ncu_round3.py > gpurun_out/r3y/ncu.log 2>&1
ncu -i gpurun_out/r3y/round3.ncu-rep --page raw --csv > gpurun_out/r3y/round3_raw.csv 2>/dev/null
rm -f gpurun_out/r3y/round3.ncu-rep
python tools/bench_lengths.py > gpurun_out/r3y/bench_lengths.log 2>&1
cp gpurun_out/bench_lengths.json gpurun_out/r3y/bench_lengths.json
