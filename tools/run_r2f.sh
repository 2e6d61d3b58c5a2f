(timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_f.log 2>&1
python - > gpurun_out/c5.log 2>&1 <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
import triqs_b200, bench
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
npts = 100_000_000
taus = torch.rand(npts, dtype=torch.float64, device=dev) * 10.0
tab = torch.from_numpy(bench.block_gtau(0, 8192, 10.0)).to(dev)
o = torch.empty(npts, 25, dtype=torch.complex128, device=dev)
for _ in range(2): ctx.interp_tau(tab, 10.0, taus, out=o)
ctx.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); ctx.interp_tau(tab, 10.0, taus, out=o); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print(json.dumps({"ms": ms, "GBps": npts * 408.0 / ms / 1e6}))
PY
