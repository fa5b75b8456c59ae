"""Per-kernel look at the thin-layer path (run under ncu --metrics gpu__time_duration.sum for the launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
rs = np.random.RandomState(0)
rnd = lambda *s: dev.from_numpy(rs.standard_normal(s).astype(np.float32))
N = int(os.environ.get("N", "64"))
dev.lib.eg_debug_set(5, int(os.environ.get("THIN", "38")))     # which passes take the conv_thin.cu route (bit 0 fwd, 1 dgrad, 2 wgrad)
CASES = [("critic l0 128", N, 128, 128, 3, 64, 4, 2, 1), ("gen last 64", N, 64, 64, 3, 64, 5, 2, 1)]
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, N, H, W, Ci, Co, k, s, p in CASES:
    OH, OW = H // s, W // s
    x, w, dy = rnd(N, H, W, Ci), rnd(k, k, Ci, Co), rnd(N, OH, OW, Co)
    y, dx, dw = dev.zeros((N, OH, OW, Co)), dev.zeros((N, H, W, Ci)), dev.zeros((k, k, Ci, Co))
    if os.environ.get("ONLY"):
        f = {"fwd": lambda: dev.conv_fwd(x, w, None, y, s, p, "tc3x"), "dgrad": lambda: dev.conv_bwd_data(dy, w, None, dx, s, p, "tc3x"),
             "wgrad": lambda: dev.conv_bwd_weight(x, dy, dw, s, p, False, "tc3x")}[os.environ["ONLY"]]
        print(name, os.environ["ONLY"], timeit(f) * 1e3, "us")
        break
    for algo in ("simt", "tc3x"):
        t1 = timeit(lambda: dev.conv_fwd(x, w, None, y, s, p, algo))
        t2 = timeit(lambda: dev.conv_bwd_data(dy, w, None, dx, s, p, algo))
        t3 = timeit(lambda: dev.conv_bwd_weight(x, dy, dw, s, p, False, algo))
        print(f"{name:16s} {algo:5s} fwd {t1*1e3:7.1f} us | dgrad {t2*1e3:7.1f} us | wgrad {t3*1e3:7.1f} us", flush=True)
