"""Timing of the instance-norm kernels (cold inputs: the buffers rotate over > L2) and of the thin-layer filter gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
rs = np.random.RandomState(0)
def timeit(f, n):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
SHAPES = [(384, 32, 32, 128), (384, 16, 16, 256), (384, 8, 8, 512), (128, 16, 32, 128), (128, 8, 16, 256), (128, 4, 8, 512),
          (128, 32, 32, 128), (128, 16, 16, 256), (128, 32, 32, 64), (128, 4, 4, 512)]
for shape in SHAPES:
    n_el = int(np.prod(shape))
    ncopy = max(2, int(600e6 // (n_el * 4 * 3)) + 1)
    xs = [torch.randn(shape, device=dev.device) for _ in range(ncopy)]
    gs = [torch.randn(shape, device=dev.device) for _ in range(ncopy)]
    ys = [torch.empty(shape, device=dev.device) for _ in range(ncopy)]
    st = dev.empty((shape[0], shape[3], 2))
    k = [0]
    def fwd():
        i = k[0] % ncopy; k[0] += 1
        dev.instnorm_fwd(xs[i], ys[i], st, "lrelu")
    def bwd():
        i = k[0] % ncopy; k[0] += 1
        dev.instnorm_bwd(xs[i], st, gs[i], None, ys[i], "lrelu")
    def bwd2():
        i = k[0] % ncopy; k[0] += 1
        dev.instnorm_bwd2(xs[i], st, gs[i], ys[i], ys[(i + 1) % ncopy], gs[(i + 1) % ncopy], "lrelu")
    mb = n_el * 4 / 1e6
    dev.lib.eg_norm_debug(-1)            # one block per slab (shared-memory resident)
    t1, t2, t3 = timeit(fwd, ncopy), timeit(bwd, ncopy), timeit(bwd2, ncopy)
    dev.lib.eg_norm_debug(1)             # streaming reduce + apply kernels at every size
    t2s = timeit(bwd, ncopy)
    dev.lib.eg_norm_debug(0)
    print(f"IN {str(shape):22s} {mb:6.1f} MB/tensor | fwd {t1*1e3:6.1f} us {2*mb/t1/1e3:5.2f} TB/s | bwd {t2*1e3:6.1f} us {3*mb/t2/1e3:5.2f} TB/s "
          f"(streaming 2-kernel: {t2s*1e3:6.1f} us) | bwd2 {t3*1e3:6.1f} us {5*mb/t3/1e3:5.2f} TB/s", flush=True)
    del xs, gs, ys
# thin-layer filter gradient: dedicated FFMA kernel vs the generic implicit GEMM (eg_debug_set(7, 1))
rnd = lambda *s: dev.from_numpy(rs.standard_normal(s).astype(np.float32))
for name, N, H, W, Ci, Co, kk, s, p in [("critic l0 3B", 384, 128, 128, 3, 64, 4, 2, 1), ("cls conv_1 8->128", 128, 64, 64, 8, 128, 3, 1, 1),
                                          ("cls img 3->128 @32", 128, 32, 32, 3, 128, 3, 1, 1), ("cls img 3->256 @16", 128, 16, 16, 3, 256, 3, 1, 1)]:
    OH, OW = (H + 2 * p - kk) // s + 1, (W + 2 * p - kk) // s + 1
    x, dy = rnd(N, H, W, Ci), rnd(N, OH, OW, Co)
    a, b = dev.zeros((kk, kk, Ci, Co)), dev.zeros((kk, kk, Ci, Co))
    fl = 2.0 * N * OH * OW * kk * kk * Ci * Co
    dev.lib.eg_debug_set(7, 1)
    t0 = timeit(lambda: dev.conv_bwd_weight(x, dy, a, s, p, False, "simt"), 5)
    dev.lib.eg_debug_set(7, 0)
    t1 = timeit(lambda: dev.conv_bwd_weight(x, dy, b, s, p, False, "simt"), 5)
    err = float((a - b).abs().max() / a.abs().max())
    print(f"thin wgrad {name:22s} generic {t0*1e3:7.1f} us {fl/t0/1e9:5.1f} TF/s | dedicated {t1*1e3:7.1f} us {fl/t1/1e9:5.1f} TF/s | rel diff {err:.1e}", flush=True)
