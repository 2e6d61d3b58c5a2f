mkdir -p gpurun_out/r3d
python tools/c2_small.py 4 > gpurun_out/r3d/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:matsubara_pipe_kernel -s 4 -c 4 -o gpurun_out/r3d/c2 python tools/c2_small.py 4 > gpurun_out/r3d/ncu.log 2>&1
ncu -i gpurun_out/r3d/c2.ncu-rep --page source --csv --print-source sass > gpurun_out/r3d/c2_sass.csv 2>/dev/null
ncu -i gpurun_out/r3d/c2.ncu-rep --page raw --csv > gpurun_out/r3d/c2_raw.csv 2>/dev/null
ls -la gpurun_out/r3d > gpurun_out/r3d/ls.txt
rm -f gpurun_out/r3d/c2.ncu-rep
