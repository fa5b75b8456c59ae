"""Diagnostic (GPU): run the classifier run d_optim2 on the reference's example images with the device operator set and
with the fp64 CPU operator set (same host code, same weights), then compare EVERY named buffer in creation order --
the first buffer that diverges names the kernel at fault.  usage: python tools/classifier_diff.py [B] [runs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import edgegan_oracle as O
from ref_ops import RefOps
from edgegan_b200.config import Flags
from edgegan_b200.models.edgegan import EdgeGAN
from edgegan_b200.ops import DeviceOps

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
runs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["d_optim2"]
ocfg = O.Config(batch_size=B, multiclasses=True)
v, u = O.init_variables(ocfg, seed=5)
d = np.load(os.path.join(ROOT, "tests", "golden", "example_images.npz"))
imgs = np.concatenate([d["train"], d["test"]], 0).astype(np.float32) / 127.5 - 1.0
inp = O.make_inputs(ocfg, seed=41)
if os.environ.get("RANDOM_IMAGES") != "1":
    inp = O.StepInputs(np.ascontiguousarray(imgs[:B]), inp.z, inp.alpha, inp.eps)
models = []
for ops in (DeviceOps(), RefOps(torch.float64)):
    if hasattr(ops, "lib"):
        ops.set_default_algo(os.environ.get("ALGO", "tc3x"))
    flags = Flags(batch_size=B, multiclasses=True)
    m = EdgeGAN(None, flags, None, ops=ops)
    m.build_train_model()
    allv = dict(v); allv.update(u)
    m.load_variables(allv)
    m.update_model(ops.from_numpy(inp.images), ops.from_numpy(inp.z), ops.from_numpy(inp.alpha), inp.eps, runs=runs)
    models.append((m, ops))
(md, od), (mr, orf) = models
torch.cuda.synchronize()
DETAIL = set(os.environ.get("DETAIL", "g_ht,g_img,g_cg,mm").split(","))
print(f"{'buffer':60s} {'max|ref|':>10s} {'rel err':>10s}")
for key, tr in orf._bufs.items():
    td = od._bufs.get(key)
    if td is None or tuple(td.shape) != tuple(tr.shape):
        continue
    a, b = od.to_numpy(td).astype(np.float64), orf.to_numpy(tr)
    if not np.isfinite(b).all():
        continue
    sc = np.abs(b).max() + 1e-30
    e = np.abs(a - b).max() / sc
    flag = " <<<" if e > 1e-4 else ""
    print(f"{key:60s} {sc:10.3e} {e:10.2e}{flag}")
    if flag and key.split("/")[-1] in DETAIL:
        idx = np.argsort(-np.abs(a - b).ravel())[:6]
        for i in idx:
            print("      ", np.unravel_index(i, a.shape), f"dev {a.ravel()[i]:+.6e} ref {b.ravel()[i]:+.6e}")
        nbad = int((np.abs(a - b) > 1e-3 * sc).sum())
        print(f"       elements off by > 1e-3 of max: {nbad} of {a.size}")
gd, gr = md.export_variables("grad"), mr.export_variables("grad")
print("---- gradients")
for n in gr:
    if n.startswith("D2/"):
        sc = np.abs(gr[n]).max() + 1e-30
        e = np.abs(gd[n].astype(np.float64) - gr[n]).max() / sc
        if e > 1e-4:
            print(f"{n:70s} {sc:10.3e} {e:10.2e}")
