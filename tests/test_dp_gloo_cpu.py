"""Data-parallel host logic (world_size 2, gloo, CPU): batch sharding, global-batch loss scaling, the gradient
all-reduce before each RMSProp apply and the sync-BN sum exchange must reproduce the single-process global-batch
step (which tests/test_host_step_cpu.py ties to the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup_path():
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)


def _build(comm, B):
    _setup_path()
    from ref_ops import RefOps
    from edgegan_b200.config import Flags
    from edgegan_b200.models.edgegan import EdgeGAN
    flags = Flags(batch_size=B, input_height=32, input_width=64, output_height=32, output_width=64,
                  multiclasses=False, image_dis_size=64, edge_dis_size=64)
    flags.num_classes = None
    m = EdgeGAN(None, flags, None, ops=RefOps(torch.float64), comm=comm, seed=5)
    m.build_train_model()
    return m


def _inputs(Bglobal):
    rs = np.random.RandomState(9)
    images = rs.uniform(-1, 1, (Bglobal, 32, 64, 3))
    z = rs.normal(size=(Bglobal, 100))
    alpha = rs.uniform(0, 1, (3, Bglobal))
    return images, z, alpha, 0.37


def _worker(rank, world, port, out_path):
    _setup_path()
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(world)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from edgegan_b200.comm import TorchDistComm
    comm = TorchDistComm("gloo")
    Bg = 4
    per = Bg // world
    m = _build(comm, per)
    images, z, alpha, eps = _inputs(Bg)
    sl = slice(rank * per, (rank + 1) * per)
    ops = m.ops
    m.update_model(ops.from_numpy(images[sl]), ops.from_numpy(z[sl]), ops.from_numpy(alpha[:, sl]), eps)
    losses = m.read_losses()          # collective: every rank takes part
    if rank == 0:
        np.savez(out_path, **{k.replace("/", "|"): a for k, a in m.export_variables("var").items()},
                 **{"loss|" + k: np.array(val) for k, val in losses.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_step_equals_global_batch_step(tmp_path):
    out = str(tmp_path / "dp.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    from edgegan_b200.models.edgegan import LocalComm
    torch.set_num_threads(4)
    m = _build(LocalComm(), 4)
    images, z, alpha, eps = _inputs(4)
    ops = m.ops
    m.update_model(ops.from_numpy(images), ops.from_numpy(z), ops.from_numpy(alpha), eps)
    want = m.export_variables("var")
    for k, a in want.items():
        g = got[k.replace("/", "|")]
        assert np.abs(g - a).max() <= 1e-9 * max(1.0, np.abs(a).max()), k
    # rank-local loss scalars are partial sums over the rank's shard (each already divided by the global batch)
    wl = m.read_losses()
    for k, val in wl.items():
        assert abs(float(got["loss|" + k]) - val) <= 1e-9 * max(1.0, abs(val)), k


def _train_worker(rank, world, port, root, out_dir):
    _setup_path()
    torch.set_num_threads(2)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ref_ops import RefOps
    from edgegan_b200.comm import TorchDistComm
    from edgegan_b200.config import Flags, update_flags
    from edgegan_b200.models.edgegan import EdgeGAN
    from edgegan_b200.utils.data import Dataset
    flags = update_flags(Flags(batch_size=2, input_height=32, input_width=64, output_height=32, output_width=64,
                               multiclasses=False, image_dis_size=64, edge_dis_size=64, epoch=1,
                               outputsroot=os.path.join(root, "outputs"), name="dp", dataroot=os.path.join(root, "data"),
                               dataset="toy", save_checkpoint_frequency=3))
    cfg = dict(input_height=32, input_width=64, output_height=32, output_width=64, crop=False, grayscale=False, z_dim=100)
    ds = Dataset(flags.dataroot, flags.dataset, flags.train_size, flags.batch_size, cfg, None, "train")
    m = EdgeGAN(None, flags, ds, ops=RefOps(torch.float64), comm=TorchDistComm("gloo"), seed=5)
    np.random.seed(100 + rank)                                   # ranks do NOT share numpy's generator state
    counter = m.train(max_steps=None, prefetch_workers=0, log=lambda *a: None)      # the WHOLE epoch: no rank may run an extra batch
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), files=np.array(ds.data), counter=counter,
             w=m.export_variables("var")["D/d_conv_3/conv2d/w"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_train_loop_shards_the_epoch_and_keeps_weights_in_sync(tmp_path):
    from PIL import Image
    root = str(tmp_path)
    rs = np.random.RandomState(0)
    # 7 files, 2 ranks, batch 2: ceil(7 / 2) = 4 = two batches for rank 0 but 3 files = one batch for rank 1 if the shards
    # were files[rank::world] of the full list -- rank 0 would then wait forever in a second gradient all-reduce
    for i in range(7):
        p = os.path.join(root, "data", "toy", "train", f"{i}.png")
        os.makedirs(os.path.dirname(p), exist_ok=True)
        Image.fromarray(rs.randint(0, 256, (32, 64, 3)).astype(np.uint8)).save(p)
    port = _free_port()
    mp.spawn(_train_worker, args=(2, port, root, root), nprocs=2, join=True)
    r0, r1 = np.load(os.path.join(root, "rank0.npz")), np.load(os.path.join(root, "rank1.npz"))
    assert len(r0["files"]) == len(r1["files"]) == 3
    assert int(r0["counter"]) == int(r1["counter"]) == 2         # one batch each (counter starts at 1)
    assert not set(r0["files"]) & set(r1["files"])               # disjoint shards of one global permutation
    assert np.array_equal(r0["w"], r1["w"])                      # same all-reduced gradients -> identical weights
    # rank 0 wrote the checkpoint (counter 2 with frequency 3)
    assert os.path.exists(os.path.join(root, "outputs", "dp", "checkpoints", "EdgeGAN-Model-2.index"))
