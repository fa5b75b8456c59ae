"""Per-shape timing of every conv launch of one update_model (CUDA events around each call)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.config import Flags
from edgegan_b200.models.edgegan import EdgeGAN
from edgegan_b200.ops import DeviceOps

B = int(os.environ.get("B", "64"))
ops = DeviceOps()
MULTI = os.environ.get("MULTI", "0") == "1"
flags = Flags(batch_size=B, multiclasses=MULTI)
if not MULTI:
    flags.num_classes = None
m = EdgeGAN(None, flags, None, ops=ops, seed=1)
m.build_train_model()
rs = np.random.RandomState(0)
img = ops.from_numpy(rs.uniform(-1, 1, (B, 64, 128, 3)))
zn = rs.normal(size=(B, 100))
if MULTI:
    zn = np.concatenate([zn, rs.randint(0, 14, (B, 1))], 1)
z = ops.from_numpy(zn)
al = ops.from_numpy(rs.uniform(0, 1, (3, B)))
for _ in range(2):
    m.update_model(img, z, al, 0.3)
torch.cuda.synchronize()
recs = []
orig = {n: getattr(ops, n) for n in ("conv_fwd", "conv_bwd_data", "conv_bwd_weight")}
def wrap(name):
    f = orig[name]
    def g(*a, **kw):
        if name == "conv_fwd": x, w, y = a[0], a[1], a[3]; macs = y.numel() * w.shape[0] * w.shape[1] * w.shape[2]; key = (name, tuple(x.shape), tuple(w.shape))
        elif name == "conv_bwd_data": dy, w, dx = a[0], a[1], a[3]; macs = dy.numel() * w.shape[0] * w.shape[1] * w.shape[2]; key = (name, tuple(dx.shape), tuple(w.shape))
        else: x, dy, dw = a[0], a[1], a[2]; macs = dy.numel() * dw.shape[0] * dw.shape[1] * dw.shape[2]; key = (name, tuple(x.shape), tuple(dw.shape))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(*a, **kw); e1.record()
        recs.append((key, 2.0 * macs, e0, e1))
    return g
for n in orig: setattr(ops, n, wrap(n))
m.update_model(img, z, al, 0.3)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n, f in orig.items(): setattr(ops, n, f)
ev0.record(); m.update_model(img, z, al, 0.3); ev1.record(); torch.cuda.synchronize()
print(f"step ms {ev0.elapsed_time(ev1):.2f}")
agg = collections.OrderedDict()
for key, fl, e0, e1 in recs:
    a = agg.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl
tot = sum(a[1] for a in agg.values())
print(f"total conv ms {tot:.2f}")
for key, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{key[0]:16s} x/dx{str(key[1]):22s} w{str(key[2]):20s} n={n:2d} {ms*1e3/n:8.1f} us/launch {ms:6.2f} ms {fl/ms/1e9:6.1f} TF/s")
