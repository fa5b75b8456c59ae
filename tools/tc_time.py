"""Timing experiments on the tcgen05 conv kernels (not a test): per-layer time / TFLOP/s for the critic shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
rs = np.random.RandomState(0)
rnd = lambda *s: dev.from_numpy(rs.standard_normal(s).astype(np.float32))
CASES = [("d_conv_1 3B", 192, 32, 64, 64, 128, 4, 2, 1), ("d_conv_3 3B", 192, 16, 32, 128, 256, 4, 2, 1),
         ("d_conv_4 3B", 192, 8, 16, 256, 512, 4, 2, 1), ("patch conv3 3B", 192, 32, 32, 128, 256, 4, 2, 1),
         ("g_dconv_2 (as conv)", 64, 16, 16, 128, 256, 5, 2, 1), ("e res 128@32", 64, 34, 34, 128, 128, 3, 1, 0)]
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
dbgs = [int(a) for a in sys.argv[1:] if not a.startswith("cap=")] or [0]
caps = [int(a[4:]) for a in sys.argv[1:] if a.startswith("cap=")]
if caps:
    for name, N, H, W, Ci, Co, k, s, p in CASES:
        OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        x, dy = rnd(N, H, W, Ci), rnd(N, OH, OW, Co)
        ref, dw = dev.zeros((k, k, Ci, Co)), dev.zeros((k, k, Ci, Co))
        dev.conv_bwd_weight(x, dy, ref, s, p, False, "simt")
        fl = 2.0 * N * OH * OW * k * k * Ci * Co
        for cap in caps:
            dev.lib.eg_debug_set(4, cap)
            t3 = timeit(lambda: dev.conv_bwd_weight(x, dy, dw, s, p, False, "tc3x"))
            err = float((dw - ref).abs().max() / ref.abs().max())
            print(f"{name:22s} cap={cap:4d} wgrad {t3*1e3:7.1f} us {fl/t3/1e9:6.1f} TF/s  relerr vs simt {err:.2e}", flush=True)
    sys.exit(0)
for name, N, H, W, Ci, Co, k, s, p in CASES:
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x, w, dy = rnd(N, H, W, Ci), rnd(k, k, Ci, Co), rnd(N, OH, OW, Co)
    y, dx, dw = dev.zeros((N, OH, OW, Co)), dev.zeros((N, H, W, Ci)), dev.zeros((k, k, Ci, Co))
    fl = 2.0 * N * OH * OW * k * k * Ci * Co
    for algo in ("tc", "tc3x"):
        for d in dbgs:
            dev.lib.eg_debug_set(3, d)
            t1 = timeit(lambda: dev.conv_fwd(x, w, None, y, s, p, algo))
            t2 = timeit(lambda: dev.conv_bwd_data(dy, w, None, dx, s, p, algo))
            t3 = timeit(lambda: dev.conv_bwd_weight(x, dy, dw, s, p, False, algo))
            print(f"{name:22s} {algo:5s} dbg={d}  fwd {t1*1e3:7.1f} us {fl/t1/1e9:6.1f} TF/s | dgrad {t2*1e3:7.1f} us {fl/t2/1e9:6.1f} | wgrad {t3*1e3:7.1f} us {fl/t3/1e9:6.1f}", flush=True)
    dev.lib.eg_debug_set(3, 0)
