"""Image I/O and dataset (edgegan_b200/utils; reference: edgegan/utils/utils.py:41-164, edgegan/utils/data/dataset.py).
scipy 1.2.2's scipy.misc cannot be installed here, so its semantics (documented in edgegan_b200/utils/utils.py) are
checked against hand-computed values and an independent PIL restatement inside this file."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from edgegan_b200 import utils as U
from edgegan_b200.utils.data import Dataset, DevicePrefetcher

CFG = dict(input_height=64, input_width=128, output_height=64, output_width=128, crop=False, grayscale=False, z_dim=100)


def test_bytescale_matches_scipy_misc_definition():
    a = np.array([[0, 100, 255]], np.uint8)
    assert U.bytescale(a) is a                                   # uint8 passes through untouched
    f = np.array([0.0, 0.5, 1.0, 0.25])
    assert U.bytescale(f).tolist() == [0, 128, 255, 64]          # (x - min) * 255 / (max - min), + 0.5, truncate
    g = np.array([30.0, 115.0, 200.0])
    assert U.bytescale(g).tolist() == [0, 128, 255]              # the data's own min / max are the range
    assert U.bytescale(np.full((2, 2), 7.0)).tolist() == [[0, 0], [0, 0]]      # constant image: scale 1, all zero
    assert U.bytescale(g, cmin=0, cmax=255).tolist() == [30, 115, 200]
    with pytest.raises(ValueError):
        U.bytescale(g, high=300)


def _write_png(path, arr):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    Image.fromarray(arr).save(path)


def _sample(rs, h=64, w=128, lo=30, hi=200):
    a = rs.randint(lo, hi + 1, (h, w, 3)).astype(np.uint8)
    a[0, 0], a[0, 1] = lo, hi
    return a


def test_transform_includes_the_contrast_stretch(tmp_path):
    rs = np.random.RandomState(0)
    a = _sample(rs)
    p = str(tmp_path / "x.png")
    _write_png(p, a)
    img = U.imread(p)
    assert img.dtype == np.float64 and np.array_equal(img, a.astype(np.float64))
    got = U.get_image(p, 64, 128, 64, 128, crop=False)
    # same size: PIL resize is the identity, what remains is bytescale of the FLOAT image: 30..200 -> 0..255 -> -1..1
    want = ((a.astype(np.float64) - 30.0) * (255.0 / 170.0)).clip(0, 255) + 0.5
    want = want.astype(np.uint8) / 127.5 - 1.0
    assert np.array_equal(got, want)
    assert got.min() == -1.0 and got.max() == 1.0
    # resize to half size: independent restatement (stretch -> PIL bilinear with (width, height) order)
    got = U.get_image(p, 64, 128, 32, 64, crop=False)
    stretched = (((a.astype(np.float64) - 30.0) * (255.0 / 170.0)).clip(0, 255) + 0.5).astype(np.uint8)
    want = np.array(Image.fromarray(stretched).resize((64, 32), Image.BILINEAR)) / 127.5 - 1.0
    assert got.shape == (32, 64, 3) and np.array_equal(got, want)
    # centre crop: rows / columns taken with round((h - crop) / 2)
    got = U.get_image(p, 32, 100, 32, 100, crop=True)
    sub = a[16:48, 14:114].astype(np.float64)
    want = (((sub - sub.min()) * (255.0 / (sub.max() - sub.min()))).clip(0, 255) + 0.5).astype(np.uint8) / 127.5 - 1.0
    assert np.array_equal(got, want)


def test_fast_byte_path_is_bit_identical(tmp_path):
    """Dataset / DevicePrefetcher decode through get_image_bytes + the 256-entry table: same float32 bits as the
    reference-shaped composition imread -> transform -> astype(float32)."""
    rs = np.random.RandomState(5)
    for k, (lo, hi) in enumerate([(30, 200), (0, 255), (7, 8), (99, 99)]):
        a = _sample(rs, 64, 128, lo, hi)
        p = str(tmp_path / f"f{k}.png")
        _write_png(p, a)
        for args in [(64, 128, 64, 128, False), (64, 128, 32, 64, False), (64, 128, 128, 256, False), (48, 100, 64, 128, True)]:
            slow = U.get_image(p, *args).astype(np.float32)
            fast = U.get_image_fast(p, *args)
            assert fast.dtype == np.float32 and np.array_equal(fast, slow), (k, args)
            b = U.get_image_bytes(p, *args)
            assert b.dtype == np.uint8 and np.array_equal(U._TO_UNIT[b], slow)
    g = U.get_image_fast(str(tmp_path / "f0.png"), 64, 128, 64, 128, False, True)       # grayscale: slow path
    assert g.shape == (64, 128) and g.dtype == np.float32


def test_palette_and_grayscale_reads(tmp_path):
    rs = np.random.RandomState(1)
    a = _sample(rs, 8, 8)
    p = str(tmp_path / "p.png")
    Image.fromarray(a).convert("P", palette=Image.ADAPTIVE, colors=16).save(p)
    img = U.imread(p)
    assert img.shape == (8, 8, 3)                                                # palette expanded to RGB
    g = U.imread(p, grayscale=True)
    assert g.shape == (8, 8) and g.dtype == np.float64


def test_merge_and_imsave_stretch(tmp_path):
    rs = np.random.RandomState(2)
    imgs = rs.uniform(-0.5, 0.8, (6, 4, 5, 3))
    sheet = U.merge(imgs, (2, 3))
    assert sheet.shape == (8, 15, 3)
    assert np.array_equal(sheet[4:8, 5:10], imgs[4])                              # idx 4 -> row 1, column 1
    assert U.merge(imgs[..., :1], (2, 3)).shape == (8, 15)
    with pytest.raises(ValueError):
        U.merge(np.zeros((2, 4, 4, 2)), (1, 2))
    p = str(tmp_path / "out" / "sheet.png")
    os.makedirs(os.path.dirname(p))
    U.save_images(imgs, (2, 3), p)
    back = np.array(Image.open(p))
    lin = U.merge(U.inverse_transform(imgs), (2, 3))
    want = (((lin - lin.min()) * (255.0 / (lin.max() - lin.min()))).clip(0, 255) + 0.5).astype(np.uint8)
    assert np.array_equal(back, want)                                             # min-max stretched on save
    assert back.min() == 0 and back.max() == 255
    assert U.image_manifold_size(6) == (2, 3)
    assert U.pathsplit("/a/b/test/3/x.png")[-3:] == ["test", "3", "x.png"]


def _make_tree(root, rs, classes=(0, 1, 2), per_class=5):
    for c in classes:
        for i in range(per_class):
            ext = "jpg" if i == 0 else "png"
            path = os.path.join(root, "data", "train", str(c), f"{c}_{i}.{ext}")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            Image.fromarray(_sample(rs)).save(path)
    for i in range(4):
        _write_png(os.path.join(root, "data", "test", str(i % 2), f"t{i}.png"), _sample(rs))
    for i in range(3):
        _write_png(os.path.join(root, "flat", "train", f"f{i}.png"), _sample(rs))


def test_dataset_layout_classes_and_z(tmp_path):
    rs = np.random.RandomState(3)
    root = str(tmp_path)
    _make_tree(root, rs)
    ds = Dataset(root, "data", 1000, 4, CFG, num_classes=3, phase="train")
    assert len(ds.data) == 15 and len(ds) == 3
    np.random.seed(7)
    ds.shuffle()
    images, z, names = ds[1]
    assert images.dtype == np.float32 and images.shape == (4, 64, 128, 3)
    assert z.shape == (4, 101) and z.dtype == np.float64
    assert z[:, -1].tolist() == [float(os.path.basename(os.path.dirname(n))) for n in names]
    np.random.seed(7)
    ds2 = Dataset(root, "data", 1000, 4, CFG, num_classes=3, phase="train")
    ds2.shuffle()
    assert ds2.data == ds.data                                                    # numpy's global generator, as in the reference
    # classes beyond num_classes are not listed; `size` caps the epoch
    assert len(Dataset(root, "data", 1000, 2, CFG, num_classes=2).data) == 10
    assert len(Dataset(root, "data", 9, 4, CFG, num_classes=3)) == 2
    # single class: flat directory, png only, z without the class column
    flat = Dataset(root, "flat", 1000, 3, CFG, num_classes=None)
    _, z, _ = flat[0]
    assert z.shape == (3, 100)
    # test phase: recursive, sorted, (images, filenames)
    te = Dataset(root, "data", 1000, 2, CFG, num_classes=3, phase="test")
    assert te.data == sorted(te.data) and len(te.data) == 4
    images, names = te[0]
    assert images.shape == (2, 64, 128, 3) and len(names) == 2
    with pytest.raises(Exception, match="No data found"):
        Dataset(root, "nothing", 10, 2, CFG, num_classes=None)
    with pytest.raises(Exception, match="less than the configured batch_size"):
        Dataset(root, "flat", 10, 8, CFG, num_classes=None)


def test_rank_shards_are_disjoint_and_cover_the_epoch(tmp_path):
    rs = np.random.RandomState(6)
    root = str(tmp_path)
    _make_tree(root, rs, per_class=6)
    shards = []
    for rank in range(3):
        ds = Dataset(root, "data", 1000, 2, CFG, num_classes=3, phase="train")
        ds.shuffle(seed=1234, rank=rank, world=3)
        assert len(ds.data) == 6 and len(ds) == 3
        shards.append(list(ds.data))
    flat = [f for s in shards for f in s]
    assert len(set(flat)) == 18                                                   # disjoint, complete
    ds.shuffle(seed=99, rank=2, world=3)
    assert set(ds.data) != set(shards[2]) and len(ds.data) == 6                   # a new epoch reshuffles the FULL list


def test_rank_shards_have_equal_batch_counts_for_awkward_file_counts():
    """ceil(n / world) a multiple of the batch while n % world != 0 (e.g. 255 files, 2 ranks, batch 64) used to give
    rank 0 one more batch than the last rank -> mismatched collectives.  All ranks must agree on len(dataset)."""
    for n, world, batch, cap in ((255, 2, 64, float("inf")), (17, 3, 2, float("inf")), (31, 4, 4, 30), (9, 8, 1, 1000)):
        lens, sizes = [], []
        for rank in range(world):
            ds = Dataset.__new__(Dataset)
            ds.data = [f"f{i:04d}.png" for i in range(n)]
            ds.batchsize, ds._cap = batch, cap
            ds.shuffle(seed=7, rank=rank, world=world)
            lens.append(len(ds)); sizes.append(len(ds.data))
        assert len(set(lens)) == 1 and len(set(sizes)) == 1, (n, world, batch, lens, sizes)
        assert sizes[0] == n // world and lens[0] == int(min(n // world, cap // world if cap != float("inf") else n)) // batch


def test_prefetcher_matches_sequential_loader(tmp_path):
    from ref_ops import RefOps
    rs = np.random.RandomState(4)
    root = str(tmp_path)
    _make_tree(root, rs, per_class=6)
    ds = Dataset(root, "data", 1000, 4, CFG, num_classes=3, phase="train")
    np.random.seed(11)
    seq = [ds[i] for i in range(len(ds))]
    np.random.seed(11)
    ops = RefOps(torch.float32)
    got = list(DevicePrefetcher(ds, ops, workers=3, depth=2))
    assert len(got) == len(seq) == 4
    for (im, z, names), (gi, gz, gn) in zip(seq, got):
        assert gn == names
        assert np.array_equal(ops.to_numpy(gi), im)
        assert np.array_equal(ops.to_numpy(gz), z.astype(np.float32))
    # a decode error surfaces in the consumer
    open(ds.data[5], "wb").write(b"not an image")
    with pytest.raises(Exception):
        list(DevicePrefetcher(ds, ops, workers=2))
    # early exit stops the producer
    it = iter(DevicePrefetcher(Dataset(root, "flat", 1000, 1, CFG), ops, workers=1))
    next(it)
    it.close()
