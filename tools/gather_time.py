"""Timing experiments on the gather-mode (thin-channel) forward of conv_tc_kmajor: which role bounds a short-K item.
dbg bits (eg_debug_set(3, .)): 1 skip the filter lo load, 2 skip the A writes to TMEM, 4 skip the output stores, 16 L2 prefetch of the items two rounds ahead ON."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
rs = np.random.RandomState(0)
rnd = lambda *s: dev.from_numpy(rs.standard_normal(s).astype(np.float32))
N = int(os.environ.get("N", "128"))
CASES = [("critic l0 128x128x3->64", N, 128, 128, 3, 64, 4, 2, 1), ("cls conv_1 64x64x8->128", N, 64, 64, 8, 128, 3, 1, 1),
         ("cls img 32x32x3->128", N, 32, 32, 3, 128, 3, 1, 1), ("dense 1x1 64x64x64->64 (TMA path)", N, 64, 64, 64, 64, 1, 1, 0),
         ("dense 1x1 64x64x64->128 (TMA path)", N, 64, 64, 64, 128, 1, 1, 0)]
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, N, H, W, Ci, Co, k, s, p in CASES:
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    # COLD inputs: rotate over enough copies of (x, y) to exceed the 126 MB L2 -- in the training step the operands of
    # these layers come from HBM, and an L2-resident micro-benchmark hides the load latency they are exposed to
    w = rnd(k, k, Ci, Co)
    nbytes = (N * H * W * Ci + N * OH * OW * Co) * 4
    ncopy = max(2, int(400e6 // nbytes) + 1)
    xs = [rnd(N, H, W, Ci) for _ in range(ncopy)]
    ys = [dev.zeros((N, OH, OW, Co)) for _ in range(ncopy)]
    x, y = xs[0], ys[0]
    mb = nbytes / 1e6
    cnt = [0]
    def cold(algo="tc3x", act=None):
        i = cnt[0] % ncopy
        cnt[0] += 1
        dev.conv_fwd(xs[i], w, None, ys[i], s, p, algo, act=act)
    row = []
    for d in (0, 16, 2, 4, 1):
        dev.lib.eg_debug_set(3, d)
        t = timeit(lambda: cold(), n=ncopy)
        row.append(f"dbg={d}: {t*1e3:6.1f} us")
    dev.lib.eg_debug_set(3, 0)
    t = timeit(lambda: cold(act="lrelu"), n=ncopy)
    row.append(f"act epi: {t*1e3:6.1f} us")
    t = timeit(lambda: cold("simt"), n=ncopy)
    row.append(f"simt: {t*1e3:6.1f} us")
    print(f"{name:36s} {mb:6.0f} MB (HBM floor {mb/6.5e3*1e3:5.1f} us) | " + " | ".join(row), flush=True)
