"""Functional ops with the reference's signatures (edgegan/nn/modules/{conv,linear,normalization,activation,
pooling}.py), executed eagerly on the device.  See edgegan_b200/nn/__init__.py."""
from __future__ import annotations

import contextlib

import numpy as np

from ..variables import VarSpec

_CTX = []


class _Context:
    def __init__(self, ops, variables=None, seed=0):
        self.ops = ops
        self.given = dict(variables or {})
        self.variables = {}
        self.rs = np.random.RandomState(seed)
        self.scope = []
        self._uid = 0

    def name(self, leaf):
        return "/".join(self.scope + [leaf])

    def get_variable(self, leaf, shape, init, std=0.02, value=0.0):
        """tf.get_variable under the current scope: reuse if it exists, else take it from `given`, else initialise."""
        name = self.name(leaf)
        t = self.variables.get(name)
        if t is None:
            if name in self.given:
                a = np.asarray(self.given[name], np.float32).reshape(shape)
            else:
                a = VarSpec(name, tuple(shape), init, std, value).sample(self.rs)
            t = self.ops.from_numpy(a)
            self.variables[name] = t
        return t

    def new(self, shape):
        self._uid += 1
        return self.ops.empty(shape)


@contextlib.contextmanager
def variable_context(ops, variables=None, seed=0):
    ctx = _Context(ops, variables, seed)
    _CTX.append(ctx)
    try:
        yield ctx
    finally:
        _CTX.pop()


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    """tf.variable_scope(name): pushes a scope component (reuse is implied by the variable table)."""
    ctx = _cur()
    ctx.scope.append(name)
    try:
        yield ctx
    finally:
        ctx.scope.pop()


def _cur():
    if not _CTX:
        raise RuntimeError("edgegan_b200.nn functions must run inside `with nn.variable_context(ops): ...`")
    return _CTX[-1]


def _same_out(size, stride):
    return -(-size // stride)


def conv2d(input, output_dim, filter_size=5, stride=2, reuse=False, pad="SAME", bias=True, name=None):
    """conv.py:13-36.  SAME / VALID / REFLECT padding; w [k,k,Cin,Cout] truncated-normal(0.02); optional bias."""
    ctx = _cur()
    ops = ctx.ops
    n, H, W, ci = input.shape
    with variable_scope(name or "conv2d", reuse):
        w = ctx.get_variable("w", (filter_size, filter_size, ci, output_dim), "trunc_normal")
        b = ctx.get_variable("b", (output_dim,), "zeros") if bias else None
    x, p = input, 0
    if pad == "REFLECT":
        pr = (filter_size - 1) // 2
        if pr:
            x = ctx.new((n, H + 2 * pr, W + 2 * pr, ci))
            ops.reflect_pad_fwd(input, x, pr)
        oh, ow = (x.shape[1] - filter_size) // stride + 1, (x.shape[2] - filter_size) // stride + 1
    elif pad == "SAME":
        oh, ow = _same_out(H, stride), _same_out(W, stride)
        p = (max((oh - 1) * stride + filter_size - H, 0) // 2, max((ow - 1) * stride + filter_size - W, 0) // 2)
    else:
        assert pad == "VALID"
        oh, ow = (H - filter_size) // stride + 1, (W - filter_size) // stride + 1
    y = ctx.new((n, oh, ow, output_dim))
    ops.conv_fwd(x, w, b, y, stride, p)
    return y


def deconv2d(input_, output_shape, with_w=False, filter_size=5, stride=2, reuse=False, name=None):
    """conv.py:39-58: tf.nn.conv2d_transpose SAME + bias; w [k,k,Cout,Cin] normal(0.02)  (SURVEY A2)."""
    ctx = _cur()
    n, h, wd, ci = input_.shape
    co = output_shape[-1]
    with variable_scope(name or "deconv2d", reuse):
        w = ctx.get_variable("w", (filter_size, filter_size, co, ci), "normal")
        b = ctx.get_variable("b", (co,), "zeros")
    y = ctx.new((n, output_shape[1], output_shape[2], co))
    # conv2d_transpose == input gradient of the SAME conv whose padding-before is (k - stride) // 2
    ctx.ops.conv_bwd_data(input_, w, b, y, stride, (filter_size - stride) // 2)
    return (y, w, b) if with_w else y


def norm(input, is_train, norm="batch", epsilon=1e-5, momentum=0.9, name=None):
    """normalization.py:10-29.  'instance': (x-mean)/(sqrt(var)+1e-5); 'batch': batch statistics with gamma/beta
    (always is_training=True in the reference); None: identity."""
    assert norm in ["instance", "batch", None]
    ctx = _cur()
    ops = ctx.ops
    if norm is None:
        return input
    y = ctx.new(input.shape)
    if norm == "instance":
        st = ctx.new((input.shape[0], input.shape[-1], 2))
        ops.instnorm_fwd(input, y, st, "none")
        return y
    C = input.shape[-1]
    with variable_scope(name or "batch_norm"):
        with variable_scope("BatchNorm"):
            beta = ctx.get_variable("beta", (C,), "zeros")
            gamma = ctx.get_variable("gamma", (C,), "ones")
    sums = ctx.new((2 * C,))
    ops.bn_stats(input.view(-1, C), sums)
    ops.bn_apply(input.view(-1, C), sums, input.numel() // C, gamma, beta, y.view(-1, C), "none")
    return y


def activation_fn(input, name="lrelu"):
    """activation.py:4-15."""
    assert name in ["relu", "lrelu", "tanh", "sigmoid", None]
    if name is None:
        return input
    ctx = _cur()
    y = ctx.new(input.shape)
    ctx.ops.act_fwd(input, y, name)
    return y


def lrelu(x, leak=0.2, name="lrelu"):
    """activation.py:30-32: tf.maximum(leak*x, x)."""
    assert abs(leak - 0.2) < 1e-12, "the device kernel implements the reference's leak = 0.2"
    ctx = _cur()
    y = ctx.new(x.shape)
    ctx.ops.act_fwd(x, y, "lrelu2")
    return y


def conv_block(input, num_filters, name, k_size, stride, is_train, reuse, norm, activation, pad="SAME", bias=False):
    """conv.py:61-67 (norm() here is the module-level function; the argument shadows it like in the reference)."""
    with variable_scope(name, reuse):
        out = conv2d(input, num_filters, k_size, stride, reuse, pad, bias)
        out = globals()["norm"](out, is_train, norm)
        return activation_fn(out, activation)


def deconv_block(input, output_shape, name, k_size, stride, is_train, reuse, norm, activation, with_w=False):
    """conv.py:124-130."""
    with variable_scope(name, reuse):
        out = deconv2d(input, output_shape, with_w, k_size, stride, reuse)
        out = globals()["norm"](out, is_train, norm)
        return activation_fn(out, activation)


def residual(input, num_filters, name, is_train, reuse, norm, pad="REFLECT", bias=False):
    """conv.py:70-85."""
    ctx = _cur()
    _norm = globals()["norm"]
    with variable_scope(name, reuse):
        with variable_scope("res1", reuse):
            out = conv2d(input, num_filters, 3, 1, reuse, pad, bias)
            out = activation_fn(_norm(out, is_train, norm), "relu")
        with variable_scope("res2", reuse):
            out = conv2d(out, num_filters, 3, 1, reuse, pad, bias)
            out = _norm(out, is_train, norm)
        with variable_scope("shortcut", reuse):
            shortcut = conv2d(input, num_filters, 1, 1, reuse, pad, bias)
        y = ctx.new(out.shape)
        ctx.ops.copy(shortcut, y)
        ctx.ops.axpby(out, y, 1.0, 1.0)
        return activation_fn(y, "relu")


def linear(input_, output_size, with_w=False, reuse=False, name=None):
    """linear.py:10-31: x @ Matrix + bias, Matrix normal(0.02)."""
    ctx = _cur()
    n, k = input_.shape
    with variable_scope(name or "Linear", reuse):
        m = ctx.get_variable("Matrix", (k, output_size), "normal")
        b = ctx.get_variable("bias", (output_size,), "zeros")
    y = ctx.new((n, output_size))
    ctx.ops.conv_fwd(input_.view(n, 1, 1, k), m.view(1, 1, k, output_size), b, y.view(n, 1, 1, output_size), 1, 0)
    return (y, m, b) if with_w else y


def mlp(input, out_dim, name, is_train, reuse, norm=None, activation=None, dtype=None, bias=True):
    """linear.py:79-92."""
    ctx = _cur()
    n, k = input.shape
    with variable_scope(name, reuse):
        w = ctx.get_variable("w", (k, out_dim), "normal")
        b = ctx.get_variable("b", (out_dim,), "zeros") if bias else None
    y = ctx.new((n, out_dim))
    ctx.ops.conv_fwd(input.view(n, 1, 1, k), w.view(1, 1, k, out_dim), b, y.view(n, 1, 1, out_dim), 1, 0)
    y = activation_fn(y, activation)
    return globals()["norm"](y, is_train, norm)


def mean_pool(input, data_format="NHWC"):
    """pooling.py:4-8 (2x2 mean).  Device tensors are NHWC; pass the NHWC buffer."""
    assert data_format == "NHWC", "device tensors are NHWC (the reference's NCHW is an API-only layout)"
    ctx = _cur()
    n, H, W, C = input.shape
    y = ctx.new((n, H // 2, W // 2, C))
    ctx.ops.add_pool2_fwd(input, None, y)
    return y
