"""Timing of the instance-norm kernels (cold inputs: the buffers rotate over > L2): one block per slab against slabs split
over thread-block clusters (the default)."""
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
SHAPES = [(128, 64, 64, 64), (384, 32, 32, 128), (384, 16, 16, 256), (384, 8, 8, 512), (128, 16, 32, 128), (128, 8, 16, 256), (128, 4, 8, 512),
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
    dev.lib.eg_norm_debug(-1)            # never the streaming pair
    dev.lib.eg_norm_debug(-2)            # one block per slab (shared-memory resident)
    o1, o2, o3 = timeit(fwd, ncopy), timeit(bwd, ncopy), timeit(bwd2, ncopy)
    dev.lib.eg_norm_debug(-3)            # default: slabs split over thread-block clusters
    t1, t2, t3 = timeit(fwd, ncopy), timeit(bwd, ncopy), timeit(bwd2, ncopy)
    if os.environ.get("ROWS"):
        res = []
        for cap in (256, 128, 512):
            dev.lib.eg_norm_debug(-cap)
            res.append((cap, timeit(fwd, ncopy) * 1e3, timeit(bwd, ncopy) * 1e3, timeit(bwd2, ncopy) * 1e3))
        dev.lib.eg_norm_debug(-256)
        print(f"IN {str(shape):22s} threads per block -> fwd / bwd / bwd2 us: " + " | ".join(f"{c}: {a:6.1f} {b:6.1f} {d:6.1f}" for c, a, b, d in res), flush=True)
    print(f"IN {str(shape):22s} {mb:6.1f} MB/tensor | fwd {o1*1e3:6.1f} -> {t1*1e3:6.1f} us {2*mb/t1/1e3:5.2f} TB/s | bwd {o2*1e3:6.1f} -> {t2*1e3:6.1f} us "
          f"{3*mb/t2/1e3:5.2f} TB/s | bwd2 {o3*1e3:6.1f} -> {t3*1e3:6.1f} us {5*mb/t3/1e3:5.2f} TB/s   (one block per slab -> clusters)", flush=True)
    del xs, gs, ys
