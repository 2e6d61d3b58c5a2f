mkdir -p gpurun_out/r4f
python tools/c2_small.py 16 > gpurun_out/r4f/plain.log 2>&1 && ncu --set full --clock-control none --cache-control none -k regex:matsubara_pipe_kernel -s 4 -c 4 -o gpurun_out/r4f/c2 python tools/c2_small.py 16 > gpurun_out/r4f/ncu.log 2>&1
ncu -i gpurun_out/r4f/c2.ncu-rep --page raw --csv > gpurun_out/r4f/c2_raw.csv 2>/dev/null
rm -f gpurun_out/r4f/c2.ncu-rep
