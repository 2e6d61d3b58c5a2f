"""Developer probe (needs a -DTQ_DEVELOPER build: TQ_LIB_PATH=tools/ab/libtqgpu_dev.so): per-direction time of every pass-1 tile length of a mesh."""
import json, os, subprocess, sys
L, n_iw, nc = (int(x) for x in sys.argv[1].split(","))
divs = [d for d in range(8, 641) if L % d == 0 and L // d <= 640 and L // d >= 8]
code = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
L, n_iw, nc = %d, %d, %d
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
def timed(fn, reps=4, warm=2):
    for _ in range(warm): fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
n_tau, n_w = L + 1, 2 * n_iw
batch = max(1, int(0.8e9 / (n_tau * nc * 16)))
gt = torch.view_as_complex(torch.randn(batch, n_tau, nc, 2, dtype=torch.float64, device=dev)).mul_(1e-3)
gw = torch.empty(batch, n_w, nc, dtype=torch.complex128, device=dev)
gt2 = torch.empty_like(gt)
mom = torch.zeros(batch, 4, nc, dtype=torch.complex128, device=dev); mom[:, 1] = 1.0
f = timed(lambda: ctx.fourier_tau_to_iw(gt, 20.0, F, n_iw, moments=mom, out=gw))
i = timed(lambda: ctx.fourier_iw_to_tau(gw, 20.0, F, n_tau, moments=mom, out=gt2))
print(json.dumps({"fwd": round(f, 3), "inv": round(i, 3)}))
''' % (L, n_iw, nc)
res = {}
for nb in [0] + divs:
    env = dict(os.environ)
    if nb:
        env["TQ_PLAN_NB_FWD"] = env["TQ_PLAN_NB_INV"] = str(nb)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    out = r.stdout.strip().splitlines()
    res[nb] = out[-1] if out else r.stderr[-200:]
    print(L, "nb", nb, "na", (L // nb if nb else 0), res[nb], flush=True)
