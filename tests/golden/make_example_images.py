"""Packs the reference's 9 example pictures (images/dataset_example/{train,test}/*.png, 64x128 RGB edge|image pairs --
the only real data the reference repository holds) into tests/golden/example_images.npz as raw bytes, so parity tests
can feed the step and the inference graph with realistic sketch|photo inputs on the GPU box, where /root/reference does
not exist.  Run in the build container from the repo root:  python tests/golden/make_example_images.py

Pixels are stored exactly as decoded (uint8, no resize: the files already are 128x64); the loaders' float conversion
x / 127.5 - 1 (reference utils/utils.py:133-135,156-160) is applied by the tests."""
import glob
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/images/dataset_example"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    out = {}
    for split in ("train", "test"):
        files = sorted(glob.glob(os.path.join(SRC, split, "*.png")))
        arr = np.stack([np.asarray(Image.open(f).convert("RGB"), np.uint8) for f in files])
        assert arr.shape[1:] == (64, 128, 3), arr.shape
        out[split] = arr
        out[split + "_names"] = np.array([os.path.basename(f) for f in files])
    np.savez_compressed(os.path.join(HERE, "example_images.npz"), **out)
    print({k: v.shape for k, v in out.items()})
