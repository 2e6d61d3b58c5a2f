mkdir -p gpurun_out/r3x
out=gpurun_out/r3x/chunk.log
: > $out
for mb in 512 128 64 40 24; do
  echo "chunk $mb" >> $out
  BLUE_CHUNK_MB=$mb python tools/len_probe.py 9999,1000,25 19999,3333,25 2>&1 | cut -c1-120 >> $out
done
python -m pytest tests -m gpu -q > gpurun_out/r3x/gputests.log 2>&1
