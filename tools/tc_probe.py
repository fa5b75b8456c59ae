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
