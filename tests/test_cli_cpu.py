"""Train / test command lines (edgegan_b200/train.py, test.py; reference edgegan/train.py, test.py and
EdgeGAN.train / .test at edgegan/models/edgegan.py:425-489, 551-633) on the CPU reference operator set: a tiny
dataset on disk -> two training steps -> checkpoint -> resume -> inference sheets."""
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from ref_ops import RefOps

from edgegan_b200 import checkpoint as ck
from edgegan_b200 import test as test_cli
from edgegan_b200 import train as train_cli
from edgegan_b200.config import parse_flags


def _tree(root, classes=(0, 1), n=3, h=32, w=64):
    rs = np.random.RandomState(0)
    for c in classes:
        for i in range(n):
            for phase in ("train", "test"):
                p = os.path.join(root, "data", "toy", phase, str(c), f"{c}{i}.png")
                os.makedirs(os.path.dirname(p), exist_ok=True)
                Image.fromarray(rs.randint(0, 256, (h, w, 3)).astype(np.uint8)).save(p)
    os.makedirs(os.path.join(root, "data", "toy", "test", "notaclass"), exist_ok=True)
    Image.fromarray(rs.randint(0, 256, (h, w, 3)).astype(np.uint8)).save(os.path.join(root, "data", "toy", "test", "notaclass", "x.png"))


def test_flag_parsing_follows_tf_app_flags():
    f = parse_flags(["--batch_size", "8", "--nomulticlasses", "--name=run1", "--crop", "--learning_rate=0.001"])
    assert f.batch_size == 8 and f.multiclasses is False and f.name == "run1" and f.crop is True and f.learning_rate == 0.001
    assert parse_flags(["--multiclasses=false"]).multiclasses is False
    assert parse_flags([]).num_classes == 14 and parse_flags([]).save_checkpoint_frequency == 500


@pytest.mark.parametrize("multiclass", [True, False])
def test_train_resume_and_test_cli(tmp_path, multiclass):
    root = str(tmp_path)
    _tree(root)
    common = ["--dataroot", os.path.join(root, "data"), "--dataset", "toy", "--outputsroot", os.path.join(root, "outputs"),
              "--name", "t", "--input_height", "32", "--input_width", "64", "--output_height", "32", "--output_width", "64",
              "--image_dis_size", "64", "--edge_dis_size", "64", "--num_classes", "2"]
    if not multiclass:
        common += ["--nomulticlasses"]
        # single class: flat train directory
        for i, f in enumerate(sorted(os.listdir(os.path.join(root, "data", "toy", "train", "0")))):
            os.replace(os.path.join(root, "data", "toy", "train", "0", f), os.path.join(root, "data", "toy", "train", f))
    train_args = common + ["--batch_size", "2", "--epoch", "3", "--save_checkpoint_frequency", "3"]
    ops = RefOps(torch.float32)
    np.random.seed(1)
    counter = train_cli.main(train_args, ops=ops, max_steps=2)
    assert counter == 3
    out = os.path.join(root, "outputs", "t")
    flags = json.load(open(os.path.join(out, "flags.json")))
    assert flags["batch_size"] == 2 and flags["checkpoint_dir"].endswith("checkpoints")
    # the reference saves when counter % frequency == 2 (edgegan.py:487): with frequency 3 that is counter 2
    st = ck.get_checkpoint_state(os.path.join(out, "checkpoints"))
    assert st is not None and st["model_checkpoint_path"] == "EdgeGAN-Model-2"
    # scalar summaries of the two steps under the reference's tags
    from edgegan_b200 import summary as sm
    logs = os.path.join(out, "logs")
    ev = sm.read_events(os.path.join(logs, sorted(os.listdir(logs))[0]))
    assert [e["step"] for e in ev] == [0, 1, 2] and "joint_dis_dloss" in ev[1]["scalars"] and "zl_loss" in ev[2]["scalars"]
    # resume: the counter continues from the checkpoint's step
    counter = train_cli.main(train_args, ops=RefOps(torch.float32), max_steps=1)
    assert counter == 3
    # inference sheets: [input | G1 | G2] = 3 x the pair width for 'full'
    written = test_cli.main(common + ["--output_combination", "full"], ops=RefOps(torch.float32))
    tdir = os.path.join(out, "test_output", "toy")
    if multiclass:
        assert written == 6                                  # the picture outside a class directory is skipped
        im = np.array(Image.open(os.path.join(tdir, "1", "10.png")))
    else:
        assert written == 7
        im = np.array(Image.open(os.path.join(tdir, "notaclass", "x.png")))
    assert im.shape == (32, 64 + 32 + 32, 3) and im.min() == 0 and im.max() == 255
    written = test_cli.main(common + ["--output_combination", "outputR"], ops=RefOps(torch.float32))
    assert np.array(Image.open(os.path.join(tdir, "0", "00.png"))).shape == (32, 32, 3)


def test_reference_module_paths_resolve(tmp_path):
    """`python -m edgegan.train --nomulticlasses ...` / `from edgegan import nn` (reference README.md:80,89,
    train.py:138, test.py:130): the `edgegan` alias package serves the reference's module paths with the SAME module
    objects as edgegan_b200 (one variable store, one CUDA library)."""
    import subprocess
    import sys
    import edgegan
    import edgegan.models.edgegan as a
    import edgegan.test as cli_test
    import edgegan.train as cli_train
    import edgegan_b200.models.edgegan as b
    from edgegan import models, nn, utils  # noqa: F401
    from edgegan.models import Discriminator, Generator  # noqa: F401
    assert a is b and nn.conv2d is __import__("edgegan_b200.nn", fromlist=["conv2d"]).conv2d
    assert cli_train.main is train_cli.main and cli_test.main is test_cli.main
    root = str(tmp_path)
    _tree(root, classes=(0,), n=2)
    for f in sorted(os.listdir(os.path.join(root, "data", "toy", "train", "0"))):
        os.replace(os.path.join(root, "data", "toy", "train", "0", f), os.path.join(root, "data", "toy", "train", f))
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "edgegan.train", "--nomulticlasses", "--dataroot", os.path.join(root, "data"), "--dataset", "toy",
           "--outputsroot", os.path.join(root, "outputs"), "--name", "t", "--input_height", "32", "--input_width", "64",
           "--output_height", "32", "--output_width", "64", "--image_dis_size", "64", "--edge_dis_size", "64", "--batch_size", "2",
           "--epoch", "1"]
    out = subprocess.run(cmd, cwd=repo, capture_output=True, text=True, timeout=600)
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stderr[-2000:]
    else:
        # flags parsed, outputs dir + flags.json written, dataset built -- then the product refuses to run without a GPU
        assert out.returncode != 0 and "CUDA device" in out.stderr
    assert json.load(open(os.path.join(root, "outputs", "t", "flags.json")))["multiclasses"] is False
