"""Flags of the reference CLIs (edgegan/train.py:14-74, edgegan/test.py:18-65) as a plain dataclass.

Same names and defaults; `update_flags` follows train.py:85-98 / test.py:69-95.
"""
from __future__ import annotations

import os
from dataclasses import asdict, dataclass


@dataclass
class Flags:
    gpu: str = "0"
    name: str = "edgegan"
    outputsroot: str = "outputs"
    epoch: int = 100
    learning_rate: float = 0.0002
    train_size: float = float("inf")
    batch_size: int = 64
    input_height: int = 64
    input_width: int = 128
    output_height: int = 64
    output_width: int = 128
    dataset: str = "class14"
    input_fname_pattern: str = "*png"
    checkpoint_dir: str = None
    logdir: str = None
    dataroot: str = "./data"
    save_checkpoint_frequency: int = 500
    crop: bool = False
    stage1_zl_loss: float = 10.0
    multiclasses: bool = True
    num_classes: int = 14
    SPECTRAL_NORM_UPDATE_OPS: str = "spectral_norm_update_ops"
    if_resnet_e: bool = True
    if_resnet_g: bool = False
    if_resnet_d: bool = False
    lambda_gp: float = 10.0
    E_norm: str = "instance"
    G_norm: str = "instance"
    D_norm: str = "instance"
    use_image_discriminator: bool = True
    image_dis_size: int = 128
    use_edge_discriminator: bool = True
    edge_dis_size: int = 128
    joint_dweight: float = 1.0
    image_dweight: float = 1.0
    edge_dweight: float = 1.0
    z_dim: int = 100
    # test.py only
    test_output_dir: str = "test_output"
    output_combination: str = "full"

    def flag_values_dict(self):
        return asdict(self)

    def validate(self):
        """The hot path implements the default model variants only (SURVEY.md 2.1 'flag-off variants')."""
        if self.if_resnet_g or self.if_resnet_d or not self.if_resnet_e:
            raise NotImplementedError("only if_resnet_e=True, if_resnet_g=False, if_resnet_d=False are implemented")
        if (self.E_norm, self.G_norm, self.D_norm) != ("instance",) * 3:
            raise NotImplementedError("only instance norm is implemented for E/G/D")
        if self.output_height % 16 or self.output_width % 32:
            raise ValueError("output_height must be a multiple of 16 and output_width of 32")
        return self


def update_flags(flags: Flags) -> Flags:
    """train.py:85-98."""
    if flags.input_width is None:
        flags.input_width = flags.input_height
    if flags.output_width is None:
        flags.output_width = flags.output_height
    if not flags.multiclasses:
        flags.num_classes = None
    path = os.path.join(flags.outputsroot, flags.name)
    flags.checkpoint_dir = os.path.join(path, "checkpoints")
    flags.logdir = os.path.join(path, "logs")
    return flags


def update_test_flags(flags: Flags) -> Flags:
    """test.py:83-96 (note: batch size forced to 1; num_classes is left alone)."""
    if flags.input_width is None:
        flags.input_width = flags.input_height
    if flags.output_width is None:
        flags.output_width = flags.output_height
    flags.batch_size = 1
    path = os.path.join(flags.outputsroot, flags.name)
    flags.checkpoint_dir = os.path.join(path, "checkpoints")
    flags.logdir = os.path.join(path, "logs")
    flags.test_output_dir = os.path.join(path, "test_output")
    return flags


def parse_flags(argv=None, description=""):
    """tf.app.flags-style command line: --name value / --name=value, booleans as --flag / --noflag / --flag=false."""
    import argparse
    import dataclasses
    ap = argparse.ArgumentParser(description=description)
    for f in dataclasses.fields(Flags):
        if f.type in ("bool", bool):
            ap.add_argument("--" + f.name, dest=f.name, nargs="?", const=True, default=f.default,
                            type=lambda s: str(s).lower() in ("1", "true", "yes"))
            ap.add_argument("--no" + f.name, dest=f.name, action="store_false")
        else:
            typ = {"int": int, "float": float, "str": str}.get(f.type if isinstance(f.type, str) else f.type.__name__, str)
            ap.add_argument("--" + f.name, type=typ, default=f.default)
    return Flags(**vars(ap.parse_args(argv)))
