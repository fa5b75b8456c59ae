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


def test_train_loop_graph_replay_equals_eager(dev, tmp_path):
    """EdgeGAN.train(use_graph=True): iterations 1-2 launch eagerly, iteration 3 is captured into a CUDA graph and
    iterations 3-5 are replays over static input buffers.  Same data order, same host random draws (z, alpha, eps) ->
    the weights after 5 iterations must equal the all-eager loop's up to the reordering of fp32 atomics."""
    from edgegan_b200.config import parse_flags, update_flags
    from edgegan_b200.models.edgegan import EdgeGAN
    from edgegan_b200.utils.data import Dataset
    root = str(tmp_path)
    _tree(root)
    args = ["--dataroot", os.path.join(root, "data"), "--dataset", "toy", "--outputsroot", os.path.join(root, "outputs"),
            "--name", "g", "--num_classes", "2", "--batch_size", "2", "--epoch", "3"]
    out = []
    for use_graph in (False, True):
        flags = update_flags(parse_flags(args))
        flags.logdir = None
        flags.checkpoint_dir = os.path.join(root, "ck_%d" % use_graph)
        cfg = {"input_height": flags.input_height, "input_width": flags.input_width, "output_height": flags.output_height,
               "output_width": flags.output_width, "crop": flags.crop, "grayscale": False, "z_dim": flags.z_dim}
        np.random.seed(7)
        ds = Dataset(flags.dataroot, flags.dataset, flags.train_size, flags.batch_size, cfg, flags.num_classes, "train")
        m = EdgeGAN(None, flags, ds, ops=dev, seed=3)
        m.build_train_model()
        logs = []
        m.train(max_steps=5, prefetch_workers=0, log=logs.append, use_graph=use_graph)
        torch.cuda.synchronize()
        out.append((m.export_variables("var"), m.read_losses(), [l for l in logs if l.startswith("Epoch")]))
    (v0, l0, log0), (v1, l1, log1) = out
    assert len(log0) == len(log1) == 5
    for k in v0:
        assert np.isfinite(v1[k]).all(), k
        d = np.abs(v0[k] - v1[k]).max()
        assert d <= 1e-4 * 5, (k, d)           # 5 RMSProp steps of lr = 2e-4 move a weight by at most ~3e-3
    for k in l0:
        assert abs(l0[k] - l1[k]) <= 2e-3 * max(1.0, abs(l0[k])), (k, l0[k], l1[k])
