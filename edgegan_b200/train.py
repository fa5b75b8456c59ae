"""`python -m edgegan_b200.train` -- the reference's training CLI (edgegan/train.py:14-138): same flag names and
defaults, outputs below <outputsroot>/<name>/{checkpoints,logs,flags.json}, dataset from <dataroot>/<dataset>/train.

One process per GPU: launched under torchrun (WORLD_SIZE > 1) every rank trains on its own shuffled view of the data
with NCCL gradient all-reduce; rank 0 writes the checkpoints."""
from __future__ import annotations

import json
import os

from .config import parse_flags, update_flags
from .utils import makedirs


def make_outputs_dir(flags):
    """train.py:78-81."""
    makedirs(flags.outputsroot)
    makedirs(flags.checkpoint_dir)
    makedirs(flags.logdir)


def save_flags(flags):
    """train.py:100-107."""
    path = os.path.join(flags.outputsroot, flags.name)
    d = flags.flag_values_dict()
    d["train_size"] = d["train_size"] if d["train_size"] != float("inf") else "inf"
    with open(os.path.join(path, "flags.json"), "w") as f:
        json.dump(d, f, indent=4)
    return flags


def main(argv=None, *, ops=None, max_steps=None):
    from .models.edgegan import EdgeGAN
    from .utils.data import Dataset
    flags = update_flags(parse_flags(argv, __doc__))
    make_outputs_dir(flags)
    save_flags(flags)
    dataset_config = {
        "input_height": flags.input_height, "input_width": flags.input_width,
        "output_height": flags.output_height, "output_width": flags.output_width,
        "crop": flags.crop, "grayscale": False, "z_dim": flags.z_dim,
    }
    comm = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from .comm import TorchDistComm
        from .ops import DeviceOps
        ops = ops or DeviceOps("cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")))
        comm = TorchDistComm()
    dataset = Dataset(flags.dataroot, flags.dataset, flags.train_size, flags.batch_size, dataset_config,
                      flags.num_classes, "train")
    model = EdgeGAN(None, flags, dataset, z_dim=flags.z_dim, ops=ops, comm=comm)
    return model.train(max_steps=max_steps)


if __name__ == "__main__":
    main()
