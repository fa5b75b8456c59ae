"""`edgegan` -- the reference's module path, served by the B200-native package.

The reference is used as `python -m edgegan.train ...` / `python -m edgegan.test ...` and `from edgegan import nn,
models, utils` (README.md:80,89; edgegan/train.py:138, test.py:130).  This alias package makes those paths resolve to
`edgegan_b200`: every submodule below is the SAME module object as its `edgegan_b200.*` counterpart (registered in
sys.modules), not a copy, so there is one variable store and one loaded CUDA library whichever name is imported.
"""
import importlib
import sys

import edgegan_b200 as _impl
from edgegan_b200 import *  # noqa: F401,F403

__version__ = _impl.__version__

for _name in ("nn", "nn.modules", "models", "models.edgegan", "models.generator", "models.discriminator",
              "models.encoder", "models.classifier", "utils", "utils.utils", "utils.data", "utils.data.dataset",
              "config", "checkpoint", "summary", "variables"):
    _mod = importlib.import_module("edgegan_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        setattr(sys.modules[__name__], _name, _mod)
