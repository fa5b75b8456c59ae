"""Device side of the loader / checkpoint / CLI rows (-m gpu): byte upload + table lookup kernel, the prefetching
feeder on a CUDA operator set, and the train -> checkpoint -> resume -> test command lines on the real kernels."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from edgegan_b200.ops import DeviceOps
    return DeviceOps()


def test_u8_lut_kernel_is_the_host_table(dev):
    from edgegan_b200.utils import _TO_UNIT
    rs = np.random.RandomState(0)
    lut = torch.from_numpy(_TO_UNIT.copy()).to(dev.device)
    for n, off in ((1, 0), (15, 0), (16, 0), (4099, 0), (4099, 3), (64 * 64 * 128 * 3, 0)):
        host = rs.randint(0, 256, n + off).astype(np.uint8)
        src = torch.from_numpy(host).to(dev.device)[off:]
        dst = torch.full((n + 5,), 7.0, device=dev.device)
        dev.u8_lut(src, lut, dst[:n])
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
        assert np.array_equal(got[:n], _TO_UNIT[host[off:]]), (n, off)
        assert np.all(got[n:] == 7.0)                                    # nothing written past the end


def _tree(root, classes=(0, 1), n=6, h=64, w=128):
    rs = np.random.RandomState(0)
    for c in classes:
        for i in range(n):
            for phase in ("train", "test"):
                p = os.path.join(root, "data", "toy", phase, str(c), f"{c}{i}.png")
                os.makedirs(os.path.dirname(p), exist_ok=True)
                Image.fromarray(rs.randint(20, 230, (h, w, 3)).astype(np.uint8)).save(p)


def test_prefetcher_device_batches_equal_the_sequential_loader(dev, tmp_path):
    from edgegan_b200.utils.data import Dataset, DevicePrefetcher
    root = str(tmp_path)
    _tree(root)
    cfg = dict(input_height=64, input_width=128, output_height=64, output_width=128, crop=False, grayscale=False, z_dim=100)
    ds = Dataset(os.path.join(root, "data"), "toy", 1000, 4, cfg, num_classes=2, phase="train")
    np.random.seed(3)
    seq = [ds[i] for i in range(len(ds))]
    np.random.seed(3)
    got = []
    for im, z, names in DevicePrefetcher(ds, dev, workers=2, depth=2):
        got.append((im.clone(), z.clone(), names))                       # slots are recycled after `depth` batches
    assert len(got) == len(seq) == 3
    for (im, z, names), (gi, gz, gn) in zip(seq, got):
        assert gn == names and gi.is_cuda and gi.dtype == torch.float32
        assert np.array_equal(gi.cpu().numpy(), im)
        assert np.array_equal(gz.cpu().numpy(), z.astype(np.float32))


def test_train_checkpoint_resume_test_cli_on_the_device(dev, tmp_path):
    from edgegan_b200 import checkpoint as ck
    from edgegan_b200 import test as test_cli
    from edgegan_b200 import train as train_cli
    from edgegan_b200.config import parse_flags, update_flags
    from edgegan_b200.models.edgegan import EdgeGAN
    root = str(tmp_path)
    _tree(root)
    common = ["--dataroot", os.path.join(root, "data"), "--dataset", "toy", "--outputsroot", os.path.join(root, "outputs"),
              "--name", "t", "--num_classes", "2"]
    train_args = common + ["--batch_size", "4", "--epoch", "2", "--save_checkpoint_frequency", "3"]
    np.random.seed(1)
    assert train_cli.main(train_args, ops=dev, max_steps=2) == 3
    ckdir = os.path.join(root, "outputs", "t", "checkpoints")
    assert ck.get_checkpoint_state(ckdir)["model_checkpoint_path"] == "EdgeGAN-Model-2"
    # the bundle holds exactly what a fresh model restores (bit-exact), including the RMSProp slots
    rd = ck.BundleReader(os.path.join(ckdir, "EdgeGAN-Model-2"))
    m = EdgeGAN(None, update_flags(parse_flags(train_args)), None, ops=dev, seed=9)
    m.build_train_model()
    assert m.load(None, ckdir) == (True, 2)
    v, ms = m.export_variables("var"), m.export_variables("ms")
    for k in ("G1/g_dconv_2/deconv2d/w", "D/d_conv_3/conv2d/w", "E/FC8_mu/w", "D2/fully_connected/weights"):
        assert np.array_equal(v[k], rd.tensor(k)), k
        assert np.array_equal(ms[k], rd.tensor(k + "/RMSProp")), k
        assert not np.all(ms[k] == 1.0), k                               # the slot moved during the two steps
    rd.close()
    assert train_cli.main(train_args, ops=dev, max_steps=1) == 3        # resume continues from step 2
    written = test_cli.main(common, ops=dev)
    assert written == 12
    im = np.array(Image.open(os.path.join(root, "outputs", "t", "test_output", "toy", "1", "13.png")))
    assert im.shape == (64, 128 + 64 + 64, 3) and im.min() == 0 and im.max() == 255


def test_train_loop_graph_step_equals_eager_step(dev):
    """The training loop's CUDA-graph iteration (EdgeGAN.train(use_graph=True) -> _GraphStep: iterations 1-2 eager,
    iteration 3 captured, then replays over static buffers) is the same function as the eager iteration: from one
    snapshot of weights + RMSProp slots, with the same batch and the same host draws of alpha / eps, the replayed step
    and update_model leave the same weights and losses (up to the reordering of fp32 atomics; the GAN step is chaotic,
    so whole trajectories of two runs cannot be compared)."""
    from edgegan_b200.config import Flags
    from edgegan_b200.models.edgegan import EdgeGAN, _GraphStep
    B = 2
    flags = Flags(batch_size=B, multiclasses=True, num_classes=2)
    m = EdgeGAN(None, flags, None, ops=dev, seed=3)
    m.build_train_model()
    rs = np.random.RandomState(5)
    batches = []
    for _ in range(4):
        z = np.concatenate([rs.normal(size=(B, 100)), rs.randint(0, 2, (B, 1))], 1).astype(np.float32)
        batches.append((dev.from_numpy(rs.uniform(-1, 1, (B, 64, 128, 3)).astype(np.float32)), dev.from_numpy(z)))
    step = _GraphStep(m)
    m._rs = np.random.RandomState(11)
    step(*batches[0])
    step(*batches[1])
    assert step.graph is None                                   # two eager iterations first
    for k in (2, 3):                                            # the capturing iteration, then a pure replay
        torch.cuda.synchronize()
        var, ms, state = m.export_variables("var"), m.export_variables("ms"), m._rs.get_state()
        step(*batches[k])
        assert step.graph is not None
        torch.cuda.synchronize()
        w_graph, l_graph = m.export_variables("var"), m.read_losses()
        m.load_variables(var)                                   # back to the snapshot ...
        for st in m.stores.values():
            st.load({n: ms[n] for n in st.offsets}, strict=True, what="ms")
        m._rs.set_state(state)
        m.update_model(*batches[k])                             # ... and the same iteration launched eagerly
        torch.cuda.synchronize()
        w_eager, l_eager = m.export_variables("var"), m.read_losses()
        # What separates a functional difference (stale alpha, a missed buffer refill: RMSProp then moves MOST weights of
        # the networks behind it by a sizeable part of a learning rate, 2e-4) from the reordering of fp32 atomics (filter
        # gradients, the thin layers' scattered input gradient), which the penalty and the later runs amplify in a FEW
        # weights (measured over 20 repetitions: worst single weight up to 0.14 lr in the critics, 0.54 lr in G / E, heavy-
        # tailed; 99th percentile of the element-wise deviation 3e-7 ... 3e-4 lr in the critics, 2e-6 ... 1.7e-2 lr in
        # G / E; the negative control below gives 1.4 lr): the 99th percentile must stay below 0.15 lr, the worst single
        # weight below 5 lr.
        lr = 2e-4
        groups = {"critics": [], "G/E": []}
        for n in w_eager:
            assert np.isfinite(w_graph[n]).all(), n
            crit = n.startswith(("D/", "D_patch2/", "D_patch3/", "D2/"))
            groups["critics" if crit else "G/E"].append(np.abs(w_graph[n] - w_eager[n]).ravel())
        for g, parts in groups.items():
            d = np.concatenate(parts)
            p99, worst = float(np.quantile(d, 0.99)), float(d.max())
            print(f"graph vs eager, iteration {k}, {g}: 99th percentile of |dw| {p99 / lr:.2e} lr (bar 0.15), worst {worst / lr:.2e} lr (bar 5)")
            assert p99 <= 0.15 * lr and worst <= 5 * lr, (k, g, p99 / lr, worst / lr)
        for n in l_eager:
            assert abs(l_graph[n] - l_eager[n]) <= 5e-3 * max(1.0, abs(l_eager[n])), (k, n, l_graph[n], l_eager[n])
    # negative control: the check has power -- the same iteration with OTHER host draws of alpha / eps (what a stale
    # buffer would amount to) is far outside the bar
    m.load_variables(var)
    for st in m.stores.values():
        st.load({n: ms[n] for n in st.offsets}, strict=True, what="ms")
    m._rs = np.random.RandomState(12345)
    m.update_model(*batches[3])
    torch.cuda.synchronize()
    w_other = m.export_variables("var")
    d = np.concatenate([np.abs(w_other[n] - w_eager[n]).ravel() for n in w_eager if n.startswith(("D/", "D_patch2/", "D_patch3/"))])
    p99 = float(np.quantile(d, 0.99))
    print(f"negative control (other alpha / eps draws), critics: 99th percentile of |dw| {p99 / lr:.2e} lr")
    assert p99 > 0.5 * lr, p99 / lr
