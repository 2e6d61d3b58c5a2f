mkdir -p gpurun_out/last
python -m pytest tests -m gpu -q > gpurun_out/last/gputests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last/smoke.log 2>&1
