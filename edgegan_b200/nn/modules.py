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
        self._used = {}                     # scope path -> names already opened there (tf default_name uniquification)

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
def variable_scope(name, reuse=None, default_name=None):
    """tf.variable_scope(name_or_None, default_name): pushes a scope component (reuse is implied by the variable
    table).  With name None the component is `default_name`, made unique inside the current scope the way TensorFlow
    does it: Conv, Conv_1, Conv_2, ... -- the classifier's layer names come from this rule."""
    ctx = _cur()
    if name is None:
        used = ctx._used.setdefault("/".join(ctx.scope), set())
        name, k = default_name, 0
        while name in used:
            k += 1
            name = f"{default_name}_{k}"
        used.add(name)
    else:
        ctx._used.setdefault("/".join(ctx.scope), set()).add(name)
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
    """pooling.py:4-8 (2x2 mean).  Device tensors are always NHWC: `data_format` is accepted for signature parity (the
    reference's classifier passes 'NCHW') and does not change the buffer layout."""
    assert data_format in ("NHWC", "NCHW")
    ctx = _cur()
    n, H, W, C = input.shape
    y = ctx.new((n, H // 2, W // 2, C))
    ctx.ops.add_pool2_fwd(input, None, y)
    return y


# ---- classifier ops (activation.py:23-27, conv.py:133-357, linear.py:34-76, normalization.py:38-76) -------------------
def relu(x):
    """tf.nn.relu, the default `activation_fn` of conv2d2 / mru_conv."""
    return activation_fn(x, "relu")


def prelu(x, name="prelu"):
    """activation.py:23-27: tf.maximum(leak * x, x) with a trainable scalar `param` (initial value 0.2)."""
    ctx = _cur()
    with variable_scope(name):
        leak = ctx.get_variable("param", (), "const", value=0.2)
    y = ctx.new(x.shape)
    ctx.ops.prelu_fwd(x, leak, y)
    return y


def spectral_normed_weight(W, scope_name):
    """normalization.py:38-76 with num_iters=1 and an update collection (u is read, never assigned inside the step):
    W / sigma(W).  `u` [1, Cout] ~ truncated normal(0, 1) lives next to the weights (the reference's doubled scope
    path is a checkpoint-name detail handled by EdgeGAN._tf_name)."""
    ctx = _cur()
    ops = ctx.ops
    cn = W.shape[-1]
    k = W.numel() // cn
    name = "/".join(ctx.scope + ["u"])
    u = ctx.variables.get(name)
    if u is None:
        a = ctx.given[name] if name in ctx.given else VarSpec(name, (1, cn), "trunc_normal", 1.0).sample(ctx.rs)
        u = ctx.variables[name] = ops.from_numpy(np.asarray(a, np.float32).reshape(1, cn))
    wbar = ctx.new(W.shape)
    ws = ctx.new((ops.sn_ws_floats(k, cn),))
    ops.spectral_norm_fwd(W, u, wbar, ws)
    return wbar


def conv2d2(inputs, num_outputs, kernel_size, sn, stride=1, rate=1, data_format='NCHW', activation_fn=relu,
            normalizer_fn=None, normalizer_params=None, weights_regularizer=None, weights_initializer=None,
            biases_initializer=0.0, biases_regularizer=None, reuse=None, scope=None,
            SPECTRAL_NORM_UPDATE_OPS='spectral_norm_update_ops'):
    """conv.py:246-295: SAME conv with optionally spectrally normalised weights [k,k,Cin,Cout] + bias [1,Cout,1,1],
    then `activation_fn` (a function of this module or None).  `weights_initializer`: None = xavier uniform, a float =
    normal(0, that std); `biases_initializer`: a constant or None (no bias).  Device tensors are NHWC whatever
    `data_format` says (see mean_pool)."""
    assert data_format == 'NCHW' and rate == 1 and normalizer_fn is None
    ctx = _cur()
    ops = ctx.ops
    n, H, W_, ci = inputs.shape
    with variable_scope(scope, reuse, default_name='Conv'):
        shape = (kernel_size, kernel_size, ci, num_outputs)
        if weights_initializer is None:         # ly.xavier_initializer(): uniform(+-sqrt(6 / (fan_in + fan_out)))
            w = ctx.get_variable("weights", shape, "uniform", std=float(np.sqrt(6.0 / (kernel_size * kernel_size * (ci + num_outputs)))))
        else:
            w = ctx.get_variable("weights", shape, "normal", std=float(weights_initializer))
        if sn:
            w = spectral_normed_weight(w, None)
        b = None
        if biases_initializer is not None:
            b = ctx.get_variable("biases", (1, num_outputs, 1, 1), "const", value=float(biases_initializer)).view(-1)
        oh, ow = _same_out(H, stride), _same_out(W_, stride)
        pt = max((oh - 1) * stride + kernel_size - H, 0) // 2
        pl = max((ow - 1) * stride + kernel_size - W_, 0) // 2
        y = ctx.new((n, oh, ow, num_outputs))
        ops.conv_fwd(inputs, w, b, y, stride, (pt, pl))
        if activation_fn is not None:
            y = activation_fn(y)
    return y


def fully_connected(inputs, num_outputs, sn, activation_fn=None, normalizer_fn=None, normalizer_params=None,
                    weights_initializer=None, weight_decay_rate=1e-6, biases_initializer=0.0, biases_regularizer=None,
                    reuse=None, scope=None, SPECTRAL_NORM_UPDATE_OPS='spectral_norm_update_ops'):
    """linear.py:34-76: x @ W (spectrally normalised when sn) + bias."""
    assert normalizer_fn is None
    ctx = _cur()
    n, k = inputs.shape
    with variable_scope(scope, reuse, default_name='fully_connected'):
        w = ctx.get_variable("weights", (k, num_outputs), "uniform", std=float(np.sqrt(6.0 / (k + num_outputs))))
        if sn:
            w = spectral_normed_weight(w, None)
        b = ctx.get_variable("biases", (num_outputs,), "const", value=float(biases_initializer or 0.0))
        y = ctx.new((n, num_outputs))
        ctx.ops.conv_fwd(inputs.view(n, 1, 1, k), w.view(1, 1, k, num_outputs), b, y.view(n, 1, 1, num_outputs), 1, 0)
        if activation_fn is not None:
            y = activation_fn(y)
    return y


def _concat_channels(a, b):
    ctx = _cur()
    n, H, W_, ca = a.shape
    cb = b.shape[3]
    out = ctx.new((n, H, W_, ca + cb))
    ctx.ops.copy_cslice(a, 0, out, 0, ca)
    ctx.ops.copy_cslice(b, 0, out, ca, cb)
    return out


def mru_conv_block_v3(inp, ht, filter_depth, sn, stride, dilate=1, activation_fn=relu, normalizer_fn=None,
                      normalizer_params=None, weights_initializer=None, biases_initializer_mask=0.5,
                      biases_initializer_h=-1, data_format='NCHW', weight_decay_rate=1e-8, norm_mask=False,
                      norm_input=True, deconv=False):
    """conv.py:133-243 (deconv=False): gate = minmax(lrelu(conv(concat(act(ht), inp)))) ; ht' = act(ht + gate *
    conv(inp)) ; out = [1x1 conv](ht) + conv(act(conv(ht'))) ; 2x2 mean pool when stride == 2."""
    assert not deconv and dilate == 1 and normalizer_fn is None and stride in (1, 2)
    ctx = _cur()
    ops = ctx.ops
    hidden_depth = ht.shape[3]
    act = activation_fn if activation_fn is not None else (lambda t: t)
    with variable_scope('norm_activation_in'):
        full_inp = _concat_channels(act(ht) if norm_input else ht, inp)
    rg = conv2d2(full_inp, hidden_depth, 3, sn=sn, stride=1, data_format=data_format, activation_fn=lrelu,
                 weights_initializer=weights_initializer, biases_initializer=biases_initializer_mask, scope='update_gate')
    n, H, W_, _ = rg.shape
    gate, mm = ctx.new(rg.shape), ctx.new((n, hidden_depth, 2))
    ops.minmax_fwd(rg, gate, mm)
    img_new = conv2d2(inp, hidden_depth, 3, sn=sn, stride=1, data_format=data_format, activation_fn=None,
                      weights_initializer=weights_initializer)
    ht_plus = ctx.new(ht.shape)
    ops.fma3(ht, gate, img_new, ht_plus)
    with variable_scope('norm_activation_merge_1'):
        ht_new_in = act(ht_plus)
    h_new = conv2d2(ht_new_in, filter_depth, 3, sn=sn, stride=1, data_format=data_format, activation_fn=activation_fn,
                    weights_initializer=weights_initializer)
    h_new = conv2d2(h_new, filter_depth, 3, sn=sn, stride=1, data_format=data_format, activation_fn=None,
                    weights_initializer=weights_initializer)
    ht_orig = ht
    if hidden_depth != filter_depth:
        ht_orig = conv2d2(ht, filter_depth, 1, sn=sn, stride=1, data_format=data_format, activation_fn=None,
                          weights_initializer=weights_initializer)
    if stride == 2:
        out = ctx.new((n, H // 2, W_ // 2, filter_depth))
        ops.add_pool2_fwd(ht_orig, h_new, out)
    else:
        out = ctx.new(h_new.shape)
        ops.copy(ht_orig, out)
        ops.axpby(h_new, out, 1.0, 1.0)
    return out


def mru_conv(x, ht, filter_depth, sn, stride=2, dilate_rate=1, num_blocks=5, last_unit=False, activation_fn=relu,
             normalizer_fn=None, normalizer_params=None, weights_initializer=None, weight_decay_rate=1e-5, unit_num=0,
             data_format='NCHW'):
    """conv.py:298-357: `num_blocks` chained MRU cells (the classifier uses 1); `last_unit` appends act() under the
    scope 'mru_conv_unit_last_norm', which sits outside the unit's scope."""
    assert len(ht) == num_blocks and dilate_rate == 1
    hts_new, inp = [], x
    for i in range(num_blocks):
        h = ht[i]
        if i > 0 and stride == 2:
            h = mean_pool(h, data_format=data_format)
        with variable_scope('mru_conv_unit_t_%d_layer_%d' % (unit_num, i)):
            inp = mru_conv_block_v3(inp, h, filter_depth, sn=sn, stride=stride if i == 0 else 1,
                                    activation_fn=activation_fn, weights_initializer=weights_initializer,
                                    data_format=data_format, weight_decay_rate=weight_decay_rate)
        hts_new.append(inp)
    if last_unit:
        with variable_scope('mru_conv_unit_last_norm'):
            hts_new[-1] = activation_fn(hts_new[-1]) if activation_fn is not None else hts_new[-1]
    return hts_new
