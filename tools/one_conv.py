"""One layer of the dominant class (classifier Conv_2 of unit 1: 3x3, 128 -> 128 at 64x64, batch 128; 0.604 GMAC per image)
launched a few times through the C ABI -- the target of the `ncu --set full` capture behind `roofline.traffic`.
usage: python tools/one_conv.py [fwd|dgrad|wgrad]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgegan_b200.ops import DeviceOps
dev = DeviceOps()
rs = np.random.RandomState(0)
N, H, C = 128, 64, 128
x = torch.randn((N, H, H, C), device=dev.device)
dy = torch.randn((N, H, H, C), device=dev.device)
w = dev.from_numpy((rs.standard_normal((3, 3, C, C)) * 0.05).astype(np.float32))
y, dw = dev.empty((N, H, H, C)), dev.empty((3, 3, C, C))
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
fs = dev.filter_set([w])
fs.prepare()
for _ in range(4):
    if what == "fwd":
        dev.conv_fwd(x, w, None, y, 1, 1, "tc3x")
    elif what == "dgrad":
        dev.conv_bwd_data(dy, w, None, y, 1, 1, "tc3x")
    else:
        dev.conv_bwd_weight(x, dy, dw, 1, 1, False, "tc3x")
torch.cuda.synchronize()
alg = (x.numel() + y.numel() + 2 * w.numel()) * 4
print(f"{what}: algorithmic bytes per launch {alg} ({alg / 1e6:.1f} MB), flop {2.0 * N * H * H * 9 * C * C:.3e}")
