"""Routes of the thin (image-side) layers, cold inputs (a 256 MB buffer is rewritten between launches):
forward  = gathered by the conditioning warps (default) | patch matrix + dense product (eg_debug_set(5, |1|8))
dgrad    = dense product + col2im pass (default) | scatter epilogue (|32) | FFMA (algo simt)
wgrad    = patch matrix + tcgen05 filter gradient (default, |4) | FFMA (algo simt)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
dev.set_default_algo("tc3x")
rs = np.random.RandomState(0)
rnd = lambda *s: dev.from_numpy(rs.standard_normal(s).astype(np.float32))
N = int(os.environ.get("N", "128"))
CASES = [("cls u1 Conv_1 8->128 k3 @64", N, 64, 64, 8, 128, 3, 1, 1, 64), ("cls u1 Conv_3 8->128 k1 @64", N, 64, 64, 8, 128, 1, 1, 0, 64),
         ("d_conv_0 3B 3->64 k4 s2 @128", 3 * N, 128, 128, 3, 64, 4, 2, 1, 64), ("d_conv_0 3->64 k4 s2 @128", N, 128, 128, 3, 64, 4, 2, 1, 64),
         ("g_dconv_4 3->64 k5 s2 @64", N, 64, 64, 3, 64, 5, 2, 1, 32), ("cls u2 Conv 3->128 k3 @32", N, 32, 32, 3, 128, 3, 1, 1, 32),
         ("cls u3 Conv 3->256 k3 @16", N, 16, 16, 3, 256, 3, 1, 1, 16)]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
def timeit(f, n=4):
    f(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3
D = 2 | 4 | 64     # + 32 = the scatter epilogue (the default since this measurement); 64 = dedicated 1x1 kernels OFF
for name, N, H, W, Ci, Co, k, s, p, OH in CASES:
    OW = OH
    x, w, dy = rnd(N, H, W, Ci), rnd(k, k, Ci, Co), rnd(N, OH, OW, Co)
    y, dx, dw = dev.zeros((N, OH, OW, Co)), dev.zeros((N, H, W, Ci)), dev.zeros((k, k, Ci, Co))
    r = {}
    for key, mask, algo in (("fwd gather", D, "tc3x"), ("fwd patch", D | 1 | 8, "tc3x")):
        dev.lib.eg_debug_set(5, mask)
        r[key] = timeit(lambda: dev.conv_fwd(x, w, None, y, s, p, algo))
    for key, mask, algo in (("dgrad col2im", D, "tc3x"), ("dgrad scatter", D | 32, "tc3x"), ("dgrad ffma", D, "simt")):
        dev.lib.eg_debug_set(5, mask)
        r[key] = timeit(lambda: dev.conv_bwd_data(dy, w, None, dx, s, p, algo))
    for key, mask, algo in (("wgrad patch", D, "tc3x"), ("wgrad ffma", D, "simt")):
        dev.lib.eg_debug_set(5, mask)
        r[key] = timeit(lambda: dev.conv_bwd_weight(x, dy, dw, s, p, False, algo))
    if k == 1:
        dev.lib.eg_debug_set(5, 2 | 4 | 32)          # default: the streaming 1x1 kernels of conv_small.cu
        r["1x1 fwd"] = timeit(lambda: dev.conv_fwd(x, w, None, y, s, p, "tc3x"))
        r["1x1 dgrad"] = timeit(lambda: dev.conv_bwd_data(dy, w, None, dx, s, p, "tc3x"))
    dev.lib.eg_debug_set(5, 2 | 4 | 32)
    floor = 4.0 * (x.numel() + dy.numel()) / 6.5e12 * 1e6
    print(f"{name:30s} HBM floor {floor:5.1f} us | " + " | ".join(f"{k_} {v:6.1f}" for k_, v in r.items()), flush=True)
