"""Filter gradient of the thin (image-side) layers: FFMA kernels (default) against the patch-matrix route on the tensor
cores (eg_debug_set(5, 2 | 4), the default since this measurement: im2col + tcgen05 filter-gradient kernel), cold inputs (a 256 MB buffer is rewritten between
launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
dev.set_default_algo("tc3x")
rs = np.random.RandomState(0)
rnd = lambda *s: dev.from_numpy(rs.standard_normal(s).astype(np.float32))
N = int(os.environ.get("N", "128"))
CASES = [("cls u1 Conv_1 8->128 k3 @64", N, 64, 64, 8, 128, 3, 1, 1), ("cls u1 Conv_3 8->128 k1 @64", N, 64, 64, 8, 128, 1, 1, 0),
         ("d_conv_0 3B 3->64 k4 s2 @128", 3 * N, 128, 128, 3, 64, 4, 2, 1), ("d_conv_0 3->64 k4 s2 @128", N, 128, 128, 3, 64, 4, 2, 1),
         ("g_dconv_4 3->64 k5 s2 @64", N, 64, 64, 3, 64, 5, 2, 1), ("cls u2 Conv 3->128 k3 @32", N, 32, 32, 3, 128, 3, 1, 1),
         ("cls u3 Conv 3->256 k3 @16", N, 16, 16, 3, 256, 3, 1, 1)]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
def timeit(f, n=4):
    f(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
for name, N, H, W, Ci, Co, k, s, p in CASES:
    OH, OW = (H + 2 * p - k + (1 if (k == 5 and s == 2) else 0)) // s + 1, (W + 2 * p - k + (1 if (k == 5 and s == 2) else 0)) // s + 1
    x, dy = rnd(N, H, W, Ci), rnd(N, OH, OW, Co)
    ref, dw = dev.zeros((k, k, Ci, Co)), dev.zeros((k, k, Ci, Co))
    dev.lib.eg_debug_set(5, 2)
    t0 = timeit(lambda: dev.conv_bwd_weight(x, dy, ref, s, p, False, "tc3x"))
    dev.lib.eg_debug_set(5, 2 | 4)
    t1 = timeit(lambda: dev.conv_bwd_weight(x, dy, dw, s, p, False, "tc3x"))
    err = float((dw - ref).abs().max() / ref.abs().max())
    nbytes = 4.0 * (x.numel() + dy.numel())
    print(f"{name:32s} FFMA {t0*1e3:7.1f} us | patch matrix + tcgen05 {t1*1e3:7.1f} us | HBM floor {nbytes/6.5e12*1e6:6.1f} us | rel diff {err:.1e}", flush=True)
