"""Explicit variable store replacing tf.get_variable / tf.variable_scope.

Each network owns ONE flat fp32 device buffer holding all of its trainables (each variable starts on a
256-byte boundary so float4 / TMA loads are legal), a same-shaped gradient buffer and the RMSProp
`rms` slot (initialised to ones).  That makes the optimizer a single elementwise kernel
and the data-parallel gradient exchange a single NCCL all-reduce per optimizer run.

Variable names, shapes and initialisers follow the reference:
  generator      models/generator.py:35-74, nn/modules/linear.py:13-27, conv.py:41-52, normalization.py:20-25
  discriminator  models/discriminator.py:58-81, conv.py:19-22 (layers d_conv_0,1,3,4 -- no d_conv_2)
  encoder        models/encoder.py:54-84, conv.py:70-85, linear.py:79-92
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import numpy as np

ALIGN = 64   # floats (256 B)


@dataclass(frozen=True)
class VarSpec:
    name: str
    shape: tuple
    init: str = "zeros"      # zeros | ones | normal | trunc_normal | const
    std: float = 0.02
    value: float = 0.0

    def sample(self, rs: np.random.RandomState) -> np.ndarray:
        if self.init == "zeros":
            return np.zeros(self.shape, np.float32)
        if self.init == "ones":
            return np.ones(self.shape, np.float32)
        if self.init == "const":
            return np.full(self.shape, self.value, np.float32)
        if self.init == "normal":              # tf.random_normal_initializer(stddev)
            return rs.normal(0.0, self.std, size=self.shape).astype(np.float32)
        if self.init == "uniform":             # xavier-style uniform(-std, std) (linear.py:36)
            return rs.uniform(-self.std, self.std, size=self.shape).astype(np.float32)
        if self.init == "trunc_normal":        # tf.truncated_normal_initializer: redraw beyond 2 sigma
            x = rs.normal(0.0, self.std, size=self.shape)
            bad = np.abs(x) > 2 * self.std
            while bad.any():
                x[bad] = rs.normal(0.0, self.std, size=int(bad.sum()))
                bad = np.abs(x) > 2 * self.std
            return x.astype(np.float32)
        raise ValueError(self.init)


def generator_specs(name, in_dim, out_h, out_w, gf_dim=64, c_dim=3):
    sh, sw = out_h // 16, out_w // 16
    chans = [gf_dim * 8, gf_dim * 4, gf_dim * 2, gf_dim, c_dim]
    v = [VarSpec(f"{name}/g_lin_0/Matrix", (in_dim, gf_dim * 8 * sh * sw), "normal"),
         VarSpec(f"{name}/g_lin_0/bias", (gf_dim * 8 * sh * sw,), "zeros"),
         VarSpec(f"{name}/batch_norm/BatchNorm/beta", (gf_dim * 8,), "zeros"),
         VarSpec(f"{name}/batch_norm/BatchNorm/gamma", (gf_dim * 8,), "ones")]
    for i in range(1, 5):
        v.append(VarSpec(f"{name}/g_dconv_{i}/deconv2d/w", (5, 5, chans[i], chans[i - 1]), "normal"))
        v.append(VarSpec(f"{name}/g_dconv_{i}/deconv2d/b", (chans[i],), "zeros"))
    return v


D_LAYERS = ("d_conv_0", "d_conv_1", "d_conv_3", "d_conv_4")


def discriminator_specs(name, in_h, in_w, df_dim=64, c_dim=3):
    chans = [c_dim, df_dim, df_dim * 2, df_dim * 4, df_dim * 8]
    v = [VarSpec(f"{name}/{l}/conv2d/w", (4, 4, chans[i], chans[i + 1]), "trunc_normal")
         for i, l in enumerate(D_LAYERS)]
    feat = (in_h // 16) * (in_w // 16) * df_dim * 8
    v.append(VarSpec(f"{name}/d_linear_5/Matrix", (feat, 1), "normal"))
    v.append(VarSpec(f"{name}/d_linear_5/bias", (1,), "zeros"))
    return v


def encoder_blocks(image_size):
    """(name suffix, filters) of the residual blocks (encoder.py:63-67)."""
    nf = [128, 256, 512, 512] + ([512] if image_size == 256 else [])
    return [(f"e_resnet_{n}_{i + 1}", n) for i, n in enumerate(nf)]


def encoder_specs(name, image_size, z_dim=100, c_dim=3):
    v = [VarSpec(f"{name}/e_resnet_64_0/conv2d/w", (4, 4, c_dim, 64), "trunc_normal"),
         VarSpec(f"{name}/e_resnet_64_0/conv2d/b", (64,), "zeros")]
    cin = 64
    for blk, n in encoder_blocks(image_size):
        for sub, k, ci in (("res1", 3, cin), ("res2", 3, n), ("shortcut", 1, cin)):
            v.append(VarSpec(f"{name}/{blk}/{sub}/conv2d/w", (k, k, ci, n), "trunc_normal"))
            v.append(VarSpec(f"{name}/{blk}/{sub}/conv2d/b", (n,), "zeros"))
        cin = n
    for fc in ("FC8_mu", "FC8_sigma"):
        v.append(VarSpec(f"{name}/{fc}/w", (cin, z_dim), "normal"))
        v.append(VarSpec(f"{name}/{fc}/b", (z_dim,), "zeros"))
    return v


class ParamStore:
    """Flat parameter / gradient / RMSProp-slot buffers of one network with named views."""

    def __init__(self, ops, specs, rs=None, conv_filter_set=True):
        self.ops = ops
        self.conv_filter_set = conv_filter_set     # False: the convs never read these weights directly (classifier: Wbar)
        self.specs = list(specs)
        self.offsets = OrderedDict()
        off = 0
        for s in self.specs:
            self.offsets[s.name] = off
            n = int(np.prod(s.shape)) if len(s.shape) else 1
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.size = off
        self.num_params = sum(int(np.prod(s.shape)) if len(s.shape) else 1 for s in self.specs)
        self.flat = ops.zeros((self.size,))
        self.grad = ops.zeros((self.size,))
        self.ms = ops.zeros((self.size,))
        ops.fill(self.ms, 1.0)
        self.var = OrderedDict()
        self.g = OrderedDict()
        for s in self.specs:
            o = self.offsets[s.name]
            n = int(np.prod(s.shape)) if len(s.shape) else 1
            self.var[s.name] = self.flat[o:o + n].view(s.shape)
            self.g[s.name] = self.grad[o:o + n].view(s.shape)
        self._fset, self._fset_algo = None, None
        if rs is not None:
            self.load({s.name: s.sample(rs) for s in self.specs})

    # ---- prepared copies of the conv filters (ops.filter_set) --------------------------------------------------
    def conv_filters(self):
        """the 4-D filters a tensor-core conv kernel reads through a prepared copy (ops.in_filter_set)"""
        return [self.var[s.name] for s in self.specs if self.ops.in_filter_set(s.shape)]

    def prepare_filters(self):
        """Refresh the library-side prepared copies of this network's conv filters: ONE kernel, enqueued after every
        write to the flat buffer (load, RMSProp).  The set is (re)built lazily for the current default conv algorithm."""
        if not self.conv_filter_set:
            return
        algo = getattr(self.ops, "default_algo", None)
        if self._fset is None or self._fset_algo != algo:
            if self._fset is not None:
                self._fset.close()
            self._fset, self._fset_algo = self.ops.filter_set(self.conv_filters()), algo
        if self._fset is not None:
            self._fset.prepare()

    def names(self):
        return list(self.offsets)

    def load(self, values, strict=True, what="var"):
        """values: name -> numpy array (e.g. oracle / checkpoint weights); what = 'var' or 'ms' (the RMSProp rms slot)."""
        dst = {"var": self.flat, "ms": self.ms}[what]
        host = np.zeros(self.size, np.float32)
        cur = None
        for s in self.specs:
            o = self.offsets[s.name]
            n = int(np.prod(s.shape)) if len(s.shape) else 1
            if s.name in values:
                a = np.asarray(values[s.name], np.float32)
                if a.size != n:
                    raise ValueError(f"{s.name}: expected shape {s.shape}, got {a.shape}")
                host[o:o + n] = a.reshape(-1)
            elif strict:
                raise KeyError(s.name)
            else:
                if cur is None:
                    cur = self.ops.to_numpy(dst)
                host[o:o + n] = cur[o:o + n]
        self.ops.upload(dst, host)
        if what == "var":
            self.prepare_filters()

    def export(self, what="var"):
        src = {"var": self.flat, "grad": self.grad, "ms": self.ms}[what]
        host = self.ops.to_numpy(src)
        out = OrderedDict()
        for s in self.specs:
            o = self.offsets[s.name]
            n = int(np.prod(s.shape)) if len(s.shape) else 1
            out[s.name] = host[o:o + n].reshape(s.shape).copy()
        return out

    def zero_grad(self):
        self.ops.fill(self.grad, 0.0)

    def rmsprop(self, lr):
        self.ops.rmsprop(self.flat, self.grad, self.ms, lr)
        self.prepare_filters()
