"""Diagnostic (not a test): run the tcgen05 conv kernels on a few shapes and print error statistics vs the fp32
SIMT kernels, so descriptor / layout mistakes can be told apart from precision noise."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from edgegan_b200.ops import DeviceOps

dev = DeviceOps()
rs = np.random.RandomState(0)


def rnd(*s, scale=1.0):
    return dev.from_numpy((rs.standard_normal(s) * scale).astype(np.float32))


def stats(tag, got, want):
    g, w = got.cpu().numpy().astype(np.float64), want.cpu().numpy().astype(np.float64)
    err = np.abs(g - w)
    mx = np.abs(w).max()
    bad = err > 1e-2 * mx
    print(f"{tag:34s} relerr {err.max() / mx:9.2e}  bad {bad.mean() * 100:6.2f}%  |got|max {np.abs(g).max():9.3e} "
          f"|want|max {mx:9.3e} nan {np.isnan(g).sum()}", flush=True)
    if bad.mean() > 0.001:
        idx = np.argwhere(bad)
        print("    first bad idx", idx[:4].tolist(), "last", idx[-2:].tolist())
        for ax in range(g.ndim):
            other = tuple(i for i in range(g.ndim) if i != ax)
            prof = bad.mean(axis=other)
            print(f"    bad fraction along axis {ax} (first 16): {np.round(prof[:16], 2).tolist()}")


CASES = [
    (4, 32, 64, 64, 128, 4, 2, 1, 16, 32),
    (8, 8, 16, 256, 512, 4, 2, 1, 4, 8),
    (2, 34, 34, 64, 128, 3, 1, 0, 32, 32),
    (2, 32, 32, 64, 128, 1, 1, 0, 32, 32),
    (4, 8, 8, 256, 512, 5, 2, 1, 4, 4),
]
which = sys.argv[1:] or ["fwd", "dgrad", "wgrad"]
if "sweep" in which:
    N, H, W, Ci, Co, k, s, p, OH, OW = (4, 32, 64, 64, 128, 4, 2, 1, 16, 32)
    x, dy = rnd(N, H, W, Ci), rnd(N, OH, OW, Co)
    a = dev.zeros((k, k, Ci, Co))
    dev.conv_bwd_weight(x, dy, a, s, p, False, "simt")
    for swz in (3, 4, 5):
        for lt in (1, 2):
            for sbo in (512, 1024):
                dev.lib.eg_debug_set(0, swz); dev.lib.eg_debug_set(1, lt); dev.lib.eg_debug_set(2, sbo)
                b = dev.zeros((k, k, Ci, Co))
                dev.conv_bwd_weight(x, dy, b, s, p, False, "tc")
                torch.cuda.synchronize()
                stats(f"wgrad swz={swz} layout={lt} sbo={sbo}", b, a)
    sys.exit(0)
for case in CASES:
    N, H, W, Ci, Co, k, s, p, OH, OW = case
    print("case", case, flush=True)
    x, w, dy = rnd(N, H, W, Ci), rnd(k, k, Ci, Co, scale=0.05), rnd(N, OH, OW, Co)
    if "fwd" in which:
        a, b = dev.zeros((N, OH, OW, Co)), dev.zeros((N, OH, OW, Co))
        dev.conv_fwd(x, w, None, a, s, p, "simt"); dev.conv_fwd(x, w, None, b, s, p, "tc")
        torch.cuda.synchronize(); stats("fwd", b, a)
    if "dgrad" in which:
        a, b = dev.zeros((N, H, W, Ci)), dev.zeros((N, H, W, Ci))
        dev.conv_bwd_data(dy, w, None, a, s, p, "simt"); dev.conv_bwd_data(dy, w, None, b, s, p, "tc")
        torch.cuda.synchronize(); stats("dgrad", b, a)
    if "wgrad" in which:
        a, b = dev.zeros((k, k, Ci, Co)), dev.zeros((k, k, Ci, Co))
        dev.conv_bwd_weight(x, dy, a, s, p, False, "simt"); dev.conv_bwd_weight(x, dy, b, s, p, False, "tc")
        torch.cuda.synchronize(); stats("wgrad", b, a)

# does the tensor core truncate or round fp32 -> tf32?  compare raw inputs with pre-truncated inputs
N, H, W, Ci, Co = 2, 16, 16, 64, 64
x, w = rnd(N, H, W, Ci), rnd(3, 3, Ci, Co, scale=0.05)
xt = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
wt = (w.view(torch.int32) & ~0x1FFF).view(torch.float32)
a, b = dev.zeros((N, H, W, Co)), dev.zeros((N, H, W, Co))
dev.conv_fwd(x, w, None, a, 1, 1, "tc"); dev.conv_fwd(xt, wt, None, b, 1, 1, "tc")
torch.cuda.synchronize()
print("tf32 operand handling: raw vs pre-truncated inputs identical:", bool(torch.equal(a, b)),
      " max diff", float((a - b).abs().max()))

# ---- accumulator precision: operands exactly representable in TF32 (multiples of 2^-8 below 4), so the only
# error source is the fp32 accumulation inside the tensor core vs round-to-nearest FFMA chains
import math
print("accumulation test (tf32-exact operands): relerr vs fp64 for K = taps*Ci")
for (N, H, W, Ci, Co, k) in ((2, 16, 16, 64, 64, 1), (2, 16, 16, 64, 64, 3), (2, 16, 16, 256, 64, 3), (2, 12, 12, 512, 64, 5)):
    xh = (np.round(rs.uniform(-4, 4, (N, H, W, Ci)) * 256) / 256).astype(np.float32)
    wh = (np.round(rs.uniform(-4, 4, (k, k, Ci, Co)) * 256) / 256).astype(np.float32)
    x, w = dev.from_numpy(xh), dev.from_numpy(wh)
    p = k // 2
    xt = torch.from_numpy(xh).double().permute(0, 3, 1, 2)
    wt = torch.from_numpy(wh).double().permute(3, 2, 0, 1)
    want = torch.nn.functional.conv2d(xt, wt, padding=p).permute(0, 2, 3, 1).numpy()
    out = {}
    for algo in ("simt", "tc", "tc3x"):
        y = dev.zeros((N, H, W, Co))
        dev.conv_fwd(x, w, None, y, 1, p, algo)
        torch.cuda.synchronize()
        g = y.cpu().numpy().astype(np.float64)
        out[algo] = (np.abs(g - want).max() / np.abs(want).max(), float(np.mean(np.sign(want) * (g - want))) / np.abs(want).max())
    print(f"  K={k*k*Ci:5d}: " + "  ".join(f"{a}: max {e[0]:.2e} signed-mean {e[1]:+.2e}" for a, e in out.items()), flush=True)
