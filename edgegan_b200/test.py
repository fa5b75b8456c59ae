"""`python -m edgegan_b200.test` -- the reference's inference CLI (edgegan/test.py:14-135): restores
<outputsroot>/<name>/checkpoints and writes [input | G1 | G2] sheets for every picture below
<dataroot>/<dataset>/test into <outputsroot>/<name>/test_output/<dataset>/.  numpy is seeded with 2333 like test.py:14."""
from __future__ import annotations

import os

import numpy as np

from .config import parse_flags, update_test_flags
from .utils import makedirs

phase = "test"


def subdirs(root):
    return [name for name in os.listdir(root) if os.path.isdir(os.path.join(root, name))]


def make_outputs_dir(flags):
    """test.py:77-80."""
    makedirs(os.path.join(flags.test_output_dir, flags.dataset))
    for path in subdirs(os.path.join(flags.dataroot, flags.dataset, phase)):
        makedirs(os.path.join(flags.test_output_dir, flags.dataset, path))


def create_dataset(flags):
    """test.py:99-112."""
    from .utils.data import Dataset
    dataset_config = {
        "input_height": flags.input_height, "input_width": flags.input_width,
        "output_height": flags.output_height, "output_width": flags.output_width,
        "crop": flags.crop, "grayscale": False,
    }
    return Dataset(flags.dataroot, flags.dataset, flags.train_size, 1, dataset_config, None, phase)


def main(argv=None, *, ops=None):
    from .models.edgegan import EdgeGAN
    np.random.seed(2333)
    flags = update_test_flags(parse_flags(argv, __doc__))
    make_outputs_dir(flags)
    model = EdgeGAN(None, flags, None, ops=ops)
    model.dataset = create_dataset(flags)
    return model.test()


if __name__ == "__main__":
    main()
