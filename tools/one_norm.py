"""One instance-norm forward / backward / second-order launch on a 64x64-pixel map ([128, 64, 64, 64]: slabs split over
thread-block clusters) -- the target of an `ncu --set full` capture (profiles/r02_instnorm_cluster_full.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
shape = (128, 64, 64, 64)
x, g, t = (torch.randn(shape, device=dev.device) for _ in range(3))
y, y2 = dev.empty(shape), dev.empty(shape)
st = dev.empty((shape[0], shape[3], 2))
for _ in range(2):
    dev.instnorm_fwd(x, y, st, "lrelu")
    dev.instnorm_bwd(x, st, g, None, y, "lrelu")
    dev.instnorm_bwd2(x, st, g, t, y, y2, "lrelu")
torch.cuda.synchronize()
print("ok")
