"""Timing experiments on the tcgen05 filter-gradient kernel (3xTF32): classifier and critic shapes, one CTA per SM with a
4-stage ring against two CTAs per SM with 2 stages each (eg_debug_set(6, 2)); result checked against the 4-stage kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
rs = np.random.RandomState(0)
rnd = lambda *s: dev.from_numpy(rs.standard_normal(s).astype(np.float32))
N = int(os.environ.get("N", "128"))
CASES = [("cls u1 Conv_2 128->128 @64", N, 64, 64, 128, 128, 3, 1, 1), ("cls u2 Conv_2 256->256 @32", N, 32, 32, 256, 256, 3, 1, 1),
         ("cls u3 Conv_2 512->512 @16", N, 16, 16, 512, 512, 3, 1, 1), ("cls u4 Conv_2 768->768 @8", N, 8, 8, 768, 768, 3, 1, 1),
         ("cls u2 Conv_1 128->256 @32", N, 32, 32, 128, 256, 3, 1, 1), ("cls u2 Conv_3 1x1 128->256", N, 32, 32, 128, 256, 1, 1, 0),
         ("d_conv_1 3B 64->128", 3 * N, 64, 64, 64, 128, 4, 2, 1), ("d_conv_3 3B 128->256", 3 * N, 32, 32, 128, 256, 4, 2, 1),
         ("d_conv_4 3B 256->512", 3 * N, 16, 16, 256, 512, 4, 2, 1)]
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, N, H, W, Ci, Co, k, s, p in CASES:
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x, dy = rnd(N, H, W, Ci), rnd(N, OH, OW, Co)
    ref, dw = dev.zeros((k, k, Ci, Co)), dev.zeros((k, k, Ci, Co))
    fl = 2.0 * N * OH * OW * k * k * Ci * Co
    dev.lib.eg_debug_set(6, 2)           # bit 1: force the 4-stage, one-CTA-per-SM kernel
    t4 = timeit(lambda: dev.conv_bwd_weight(x, dy, ref, s, p, False, "tc3x"))
    dev.lib.eg_debug_set(6, 0)           # default: two CTAs per SM, 2 stages each (single-tap layouts)
    t2 = timeit(lambda: dev.conv_bwd_weight(x, dy, dw, s, p, False, "tc3x"))
    err = float((dw - ref).abs().max() / ref.abs().max())
    print(f"{name:30s} 4 stages x 1 CTA {t4*1e3:7.1f} us {fl/t4/1e9:6.1f} TF/s | 2 stages x 2 CTAs {t2*1e3:7.1f} us {fl/t2/1e9:6.1f} TF/s | rel diff {err:.1e}", flush=True)
