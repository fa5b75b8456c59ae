"""Eager mirror of the reference op library `edgegan.nn` (edgegan/nn/__init__.py:1, nn/modules/*.py).

Same function names, positional order and defaults as the reference; what TensorFlow supplied
implicitly is an explicit context here:

    with nn.variable_context(ops, variables=None, seed=0) as ctx:      # replaces tf.variable_scope / get_variable
        y = nn.conv_block(x, 64, 'd_conv_0', 4, 2, True, False, None, 'lrelu')

The classifier-side ops (`conv2d2`, `mru_conv`, `fully_connected`, `prelu`, `spectral_normed_weight`) are here too;
`variable_scope(None, default_name='Conv')` numbers repeated default scopes like TensorFlow (Conv, Conv_1, ...).
Variables are created on first use under '/'-joined scope names with the reference's initialisers (or taken from
`variables`, a name -> numpy dict, e.g. the oracle's or a checkpoint's) and kept in `ctx.variables`.  Every
function runs the CUDA kernels behind include/edgegan_b200.h immediately and returns a new device tensor; this
layer is forward-only -- the training step uses the layer objects in edgegan_b200.models, which carry the explicit
backward passes.
"""
from .modules import (activation_fn, conv2d, conv2d2, conv_block, deconv2d, deconv_block, fully_connected, linear,  # noqa: F401
                      lrelu, mean_pool, mlp, mru_conv, mru_conv_block_v3, norm, prelu, relu, residual,
                      spectral_normed_weight, variable_context, variable_scope)
