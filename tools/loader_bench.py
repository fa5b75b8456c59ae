"""Throughput of the dataset loader (decode + scipy.misc-style resize + normalise) on synthetic 64x128 PNGs:
images/s for the sequential reference-style loop and for DevicePrefetcher's decode pool.  Host-only (CPU)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image
from concurrent.futures import ThreadPoolExecutor
from edgegan_b200.utils.data import Dataset

n = int(os.environ.get("N", "2048"))
cfg = dict(input_height=64, input_width=128, output_height=64, output_width=128, crop=False, grayscale=False, z_dim=100)
with tempfile.TemporaryDirectory() as root:
    rs = np.random.RandomState(0)
    d = os.path.join(root, "data", "train")
    os.makedirs(d)
    for i in range(n):
        Image.fromarray(rs.randint(0, 256, (64, 128, 3)).astype(np.uint8)).save(os.path.join(d, f"{i}.png"))
    ds = Dataset(root, "data", n, 64, cfg, num_classes=None)
    t = time.perf_counter()
    for i in range(len(ds)):
        ds[i]
    dt = time.perf_counter() - t
    print(f"sequential          {len(ds) * 64 / dt:9.0f} images/s")
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    for w in (2, 4, 8, 16, 32):
        if w > 2 * os.cpu_count():
            break
        with ProcessPoolExecutor(w, mp_context=mp.get_context("fork")) as pool:
            ds.finish_batch(0, ds.submit_batch(0, pool, w))                 # start the workers
            t = time.perf_counter()
            pending = {}
            for i in range(len(ds)):                                         # DevicePrefetcher's producer loop
                for k in range(i, min(len(ds), i + 3)):
                    if k not in pending:
                        pending[k] = ds.submit_batch(k, pool, w)
                images, z, names = ds.finish_batch(i, pending.pop(i))
            dt = time.perf_counter() - t
        print(f"decode pool {w:2d} proc {len(ds) * 64 / dt:9.0f} images/s   ({images.dtype} batches, host cores: {os.cpu_count()})")
