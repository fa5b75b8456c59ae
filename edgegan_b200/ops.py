"""Device operator set of the EdgeGAN hot path: thin wrappers that hand torch CUDA buffers to the
C ABI (include/edgegan_b200.h).  torch is used for device memory and streams only -- no torch op
computes anything here.  Every method writes into caller-provided output buffers and is
asynchronous on the current CUDA stream.

The same method surface is re-implemented on the CPU in ``tests/ref_ops.py`` (test infrastructure)
so that the host-side derivations in ``edgegan_b200.models`` can be checked against the oracle's
autograd without a GPU; the product never imports that file.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ConvShape

ACT = {None: 0, "none": 0, "relu": 1, "lrelu": 2, "tanh": 3, "sigmoid": 4, "lrelu2": 5}
ALGO = {None: 0, "auto": 0, "simt": 1, "tc": 2, "tc3x": 3}
IN_EPS = 1e-5      # normalization.py:15
BN_EPS = 1e-5      # normalization.py:11 (epsilon default)


def _p(t):
    return None if t is None else t.data_ptr()


class FilterSet:
    """Library-side prepared copies of all tensor-core conv filters of one network (include/edgegan_b200.h,
    eg_filter_set_*): `prepare()` = one kernel on the current stream, to be enqueued after every write to the
    filters.  The set is destroyed when any of its filter tensors is garbage collected (before its memory can be
    reused), so a stale pointer can never be served."""

    def __init__(self, ops, filters, algo):
        import weakref
        self.ops, self.algo = ops, algo
        descs = (_lib.FilterDesc * len(filters))()
        for d, w in zip(descs, filters):
            kh, kw, ci, co = w.shape
            if not w.is_contiguous():
                raise ValueError("filters of a prepared-filter set must be contiguous")
            d.w, d.taps, d.Ci, d.Co = w.data_ptr(), kh * kw, ci, co
        h = C.c_longlong(-1)
        _lib.check(ops.lib.eg_filter_set_create(descs, len(filters), ALGO[algo], C.byref(h)), "eg_filter_set_create")
        self.handle = h.value
        lib, handle = ops.lib, self.handle
        self._fin = [weakref.finalize(w, lib.eg_filter_set_destroy, handle) for w in filters]
        for f in self._fin:
            f.atexit = False         # at interpreter exit the CUDA context may already be gone; the driver frees the memory

    def prepare(self):
        _lib.check(self.ops.lib.eg_filter_set_prepare(self.handle, self.ops._st), "eg_filter_set_prepare")

    def close(self):
        for f in self._fin:
            f.detach()
        self.ops.lib.eg_filter_set_destroy(self.handle)


class SpectralNormSet:
    """Device-side descriptor table of the spectrally normalised weights of one network; `fwd()` / `bwd()` enqueue 3 / 4
    graph nodes for the whole network on the current stream.  Destroyed with the first of its tensors (FilterSet's rule)."""

    def __init__(self, ops, items):
        import weakref
        self.ops = ops
        descs = (_lib.SnDesc * len(items))()
        keep = []
        for d, it in zip(descs, items):
            W = it["W"]
            Cn = W.shape[-1]
            d.K, d.C = W.numel() // Cn, Cn
            for name in ("W", "u", "Wbar", "ws", "G", "gW", "Wa", "Wi", "Ga", "Gi"):
                t = it.get(name)
                if t is not None:
                    if not t.is_contiguous():
                        raise ValueError(f"spectral_norm_set: {name} must be contiguous")
                    setattr(d, name, t.data_ptr())
                    keep.append(t)
            if it.get("Wa") is not None:
                d.cin, d.hd = W.shape[-2], it["hd"]
        h = C.c_longlong(-1)
        _lib.check(ops.lib.eg_spectral_norm_set_create(descs, len(items), C.byref(h)), "eg_spectral_norm_set_create")
        self.handle = h.value
        lib, handle = ops.lib, self.handle
        self._fin = [weakref.finalize(t, lib.eg_spectral_norm_set_destroy, handle) for t in keep]
        for f in self._fin:
            f.atexit = False

    def fwd(self):
        _lib.check(self.ops.lib.eg_spectral_norm_set_fwd(self.handle, self.ops._st), "eg_spectral_norm_set_fwd")

    def bwd(self):
        _lib.check(self.ops.lib.eg_spectral_norm_set_bwd(self.handle, self.ops._st), "eg_spectral_norm_set_bwd")

    def close(self):
        for f in self._fin:
            f.detach()
        _lib.check(self.ops.lib.eg_spectral_norm_set_destroy(self.handle), "eg_spectral_norm_set_destroy")


class DeviceOps:
    """CUDA implementation (the only one the product has)."""

    def __init__(self, device="cuda:0"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("edgegan_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self.lib.eg_device_info(C.byref(sm), C.byref(maj), C.byref(mnr)), "eg_device_info")
        self.sm_count, self.cc = sm.value, (maj.value, mnr.value)
        self._bufs = {}
        self._shape_cache = {}
        self.pass_algo = {}          # debugging: per-pass algorithm override {'fwd'|'dgrad'|'wgrad': name}

    @property
    def launches(self):
        """kernels launched by the library so far (counted inside the library, eg_kernel_launches)"""
        return int(self.lib.eg_kernel_launches())

    # ---- memory -----------------------------------------------------------------------------
    def empty(self, shape):
        return torch.empty(tuple(shape), dtype=torch.float32, device=self.device)

    def zeros(self, shape):
        t = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
        self.fill(t, 0.0)
        return t

    def buf(self, key, shape):
        """Persistent named scratch buffer (allocated once -> the step is CUDA-graph capturable)."""
        shape = tuple(int(s) for s in shape)
        t = self._bufs.get(key)
        if t is None or t.shape != shape:
            t = torch.empty(shape, dtype=torch.float32, device=self.device)
            self._bufs[key] = t
        return t

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

    def to_numpy(self, t):
        return t.detach().cpu().numpy()

    def upload(self, dst, src_host):
        """host (numpy / pinned torch) -> existing device buffer"""
        if isinstance(src_host, np.ndarray):
            src_host = torch.from_numpy(np.ascontiguousarray(src_host, dtype=np.float32))
        dst.copy_(src_host.reshape(dst.shape), non_blocking=True)

    def bytes_allocated(self):
        return sum(t.numel() * 4 for t in self._bufs.values())

    @property
    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    def set_default_algo(self, algo):
        _lib.check(self.lib.eg_set_default_algo(ALGO[algo]), "eg_set_default_algo")

    @property
    def default_algo(self):
        """what the library's EG_ALGO_AUTO currently resolves to for tensor-core-capable layers ('tc3x' unless
        set_default_algo chose otherwise)"""
        code = int(self.lib.eg_get_default_algo())
        name = {v: k for k, v in ALGO.items() if k}[code]
        return "tc3x" if name == "auto" else name

    # ---- convolution trio -------------------------------------------------------------------
    def _cs(self, xs, ws, ys, stride, pad):
        key = (tuple(xs), tuple(ws), tuple(ys), stride, pad)
        s = self._shape_cache.get(key)
        if s is None:
            N, H, W, Ci = xs
            KH, KW, wci, Co = ws
            N2, OH, OW, Co2 = ys
            if not (N == N2 and wci == Ci and Co == Co2):
                raise ValueError(f"conv shape mismatch x{tuple(xs)} w{tuple(ws)} y{tuple(ys)}")
            pt, pl = (pad, pad) if isinstance(pad, int) else pad
            s = ConvShape(N, H, W, Ci, OH, OW, Co, KH, KW, stride, pt, pl)
            self._shape_cache[key] = s
        return s

    @staticmethod
    def _epi(act, mask):
        """fused epilogue selector: act alone -> out = act(v); act + mask -> out = v * act'(mask)"""
        if mask is not None:
            return 2, ACT[act], _p(mask)
        if act not in (None, "none"):
            return 1, ACT[act], None
        return 0, 0, None

    def conv_fwd(self, x, w, bias, y, stride, pad, algo=None, act=None, mask=None):
        """y = conv(x, w) (+bias): x[N,H,W,Ci] w[KH,KW,Ci,Co] y[N,OH,OW,Co]; zero pad `pad` before.
        act: fused activation; act + mask: y = conv * act'(mask) (mask shaped like y)."""
        s = self._cs(x.shape, w.shape, y.shape, stride, pad)
        epi, a, m = self._epi(act, mask)
        _lib.check(self.lib.eg_conv2d_fwd_ex(C.byref(s), _p(x), _p(w), _p(bias), _p(y), epi, a, m,
                                             ALGO[algo or self.pass_algo.get("fwd")], self._st), "conv2d_fwd")

    def conv_bwd_data(self, dy, w, bias, dx, stride, pad, algo=None, act=None, mask=None):
        """dx = conv input-gradient (== conv2d_transpose forward) (+bias over dx channels); act / mask as in conv_fwd."""
        s = self._cs(dx.shape, w.shape, dy.shape, stride, pad)
        epi, a, m = self._epi(act, mask)
        _lib.check(self.lib.eg_conv2d_bwd_data_ex(C.byref(s), _p(dy), _p(w), _p(bias), _p(dx), epi, a, m,
                                                  ALGO[algo or self.pass_algo.get("dgrad")], self._st), "conv2d_bwd_data")

    def conv_bwd_weight(self, x, dy, dw, stride, pad, accumulate=False, algo=None):
        """dw (+)= filter gradient."""
        s = self._cs(x.shape, dw.shape, dy.shape, stride, pad)
        _lib.check(self.lib.eg_conv2d_bwd_weight(C.byref(s), _p(x), _p(dy), _p(dw), int(accumulate), ALGO[algo or self.pass_algo.get("wgrad")], self._st), "conv2d_bwd_weight")

    def bias_grad(self, dy, db, accumulate=False):
        Cn = dy.shape[-1]
        _lib.check(self.lib.eg_bias_grad(_p(dy), dy.numel() // Cn, Cn, _p(db), int(accumulate), self._st), "bias_grad")

    # ---- norms / activations ----------------------------------------------------------------
    @staticmethod
    def _npc(x):
        N, Cn = x.shape[0], x.shape[-1]
        return N, x.numel() // (N * Cn), Cn

    def instnorm_fwd(self, x, y, stats, act):
        N, P, Cn = self._npc(x)
        _lib.check(self.lib.eg_instnorm_fwd(_p(x), _p(y), _p(stats), N, P, Cn, IN_EPS, ACT[act], self._st), "instnorm_fwd")

    def instnorm_bwd(self, x, stats, gy, addend, gx, act):
        N, P, Cn = self._npc(x)
        _lib.check(self.lib.eg_instnorm_bwd(_p(x), _p(stats), _p(gy), _p(addend), _p(gx), N, P, Cn, IN_EPS, ACT[act], self._st), "instnorm_bwd")

    def instnorm_bwd2(self, x, stats, gy, t, out_gy, out_x, act):
        N, P, Cn = self._npc(x)
        _lib.check(self.lib.eg_instnorm_bwd2(_p(x), _p(stats), _p(gy), _p(t), _p(out_gy), _p(out_x), N, P, Cn, IN_EPS, ACT[act], self._st), "instnorm_bwd2")

    def act_fwd(self, x, y, act):
        _lib.check(self.lib.eg_act_fwd(_p(x), _p(y), x.numel(), ACT[act], self._st), "act_fwd")

    def act_bwd(self, x_pre, gy, gx, act):
        _lib.check(self.lib.eg_act_bwd(_p(x_pre), _p(gy), _p(gx), x_pre.numel(), ACT[act], self._st), "act_bwd")

    def bn_stats(self, x, sums):
        Cn = x.shape[-1]
        _lib.check(self.lib.eg_bn_stats(_p(x), _p(sums), x.numel() // Cn, Cn, self._st), "bn_stats")

    def bn_apply(self, x, sums, count, gamma, beta, y, act):
        Cn = x.shape[-1]
        _lib.check(self.lib.eg_bn_apply(_p(x), _p(sums), float(count), _p(gamma), _p(beta), _p(y), x.numel() // Cn, Cn, BN_EPS, ACT[act], self._st), "bn_apply")

    def bn_bwd_reduce(self, x, sums, count, gamma, beta, gy, red, act):
        Cn = x.shape[-1]
        _lib.check(self.lib.eg_bn_bwd_reduce(_p(x), _p(sums), float(count), _p(gamma), _p(beta), _p(gy), _p(red), x.numel() // Cn, Cn, BN_EPS, ACT[act], self._st), "bn_bwd_reduce")

    def bn_bwd_apply(self, x, sums, count, gamma, beta, gy, red, gx, act):
        Cn = x.shape[-1]
        _lib.check(self.lib.eg_bn_bwd_apply(_p(x), _p(sums), float(count), _p(gamma), _p(beta), _p(gy), _p(red), _p(gx), x.numel() // Cn, Cn, BN_EPS, ACT[act], self._st), "bn_bwd_apply")

    # ---- discriminator head -------------------------------------------------------------------
    def rowdot_fwd(self, h, w, bias, d):
        B = h.shape[0]
        _lib.check(self.lib.eg_rowdot_fwd(_p(h), _p(w), _p(bias), _p(d), B, h.numel() // B, self._st), "rowdot_fwd")

    def rowdot_bwd_input(self, gd, w, gh):
        B = gh.shape[0]
        _lib.check(self.lib.eg_rowdot_bwd_input(_p(gd), _p(w), _p(gh), B, gh.numel() // B, self._st), "rowdot_bwd_input")

    def rowdot_bwd_weight(self, gd, h, gw, gb, accumulate=False):
        B = h.shape[0]
        _lib.check(self.lib.eg_rowdot_bwd_weight(_p(gd), _p(h), _p(gw), _p(gb), B, h.numel() // B, int(accumulate), self._st), "rowdot_bwd_weight")

    # ---- resize / slices ----------------------------------------------------------------------
    def bicubic_up2_fwd(self, x, y):
        N, H, W, Cn = x.shape
        _lib.check(self.lib.eg_bicubic_up2_fwd(_p(x), _p(y), N, H, W, Cn, self._st), "bicubic_up2_fwd")

    def bicubic_up2_bwd(self, gy, gx):
        N, H, W, Cn = gx.shape
        _lib.check(self.lib.eg_bicubic_up2_bwd(_p(gy), _p(gx), N, H, W, Cn, self._st), "bicubic_up2_bwd")

    def copy_wslice(self, src, src_w0, dst, dst_w0, width):
        """dst[:, :, dst_w0:dst_w0+width, :] = src[:, :, src_w0:src_w0+width, :]  (NHWC width slices)"""
        N, H, Ws, Cn = src.shape
        Wd = dst.shape[2]
        _lib.check(self.lib.eg_copy2d(src.data_ptr() + 4 * src_w0 * Cn, Ws * Cn, dst.data_ptr() + 4 * dst_w0 * Cn,
                                      Wd * Cn, N * H, width * Cn, self._st), "copy2d")

    def copy2d(self, src, src_off, src_stride, dst, dst_off, dst_stride, rows, cols):
        """dst.flat[dst_off + r*dst_stride + c] = src.flat[src_off + r*src_stride + c]  (element units)"""
        _lib.check(self.lib.eg_copy2d(src.data_ptr() + 4 * src_off, src_stride, dst.data_ptr() + 4 * dst_off, dst_stride,
                                      rows, cols, self._st), "copy2d")

    def copy(self, src, dst):
        _lib.check(self.lib.eg_copy2d(_p(src), src.numel(), _p(dst), dst.numel(), 1, src.numel(), self._st), "copy2d")

    # ---- prepared-filter sets (eg_filter_set_*) -------------------------------------------------------
    @staticmethod
    def in_filter_set(shape):
        """filters [kh, kw, ci, co] the library keeps prepared copies of: both channel counts multiples of 32 (the operand
        layouts of the tcgen05 kernels), or a thin image-side filter (ci <= 8: the gathered forward's K-major copy and the
        patch-matrix input gradient's padded copy)"""
        if len(shape) != 4:
            return False
        kh, kw, ci, co = shape
        return (ci % 32 == 0 and co % 32 == 0 and kh * kw <= 25) or (ci <= 8 and co % 32 == 0 and kh * kw * ci <= 128)

    def filter_set(self, filters):
        """-> FilterSet over the given 4-D HWIO filter tensors (or None when the default conv algorithm is the SIMT
        path, which reads the filters as they are).  Call `.prepare()` after every write to those tensors."""
        algo = self.default_algo
        if algo not in ("tc", "tc3x") or not filters:
            return None
        return FilterSet(self, filters, algo)

    def filter_set_hits(self):
        return int(self.lib.eg_filter_set_hits())

    def u8_lut(self, src_u8, lut, dst, stream=None):
        """dst[i] = lut[src[i]] (image bytes -> float with a 256-entry device table); `stream`: torch stream or None"""
        st = self._st if stream is None else C.c_void_p(stream.cuda_stream)
        _lib.check(self.lib.eg_u8_lut_f32(_p(src_u8), _p(lut), _p(dst), src_u8.numel(), st), "u8_lut")

    def fill(self, dst, value):
        _lib.check(self.lib.eg_fill(_p(dst), dst.numel(), float(value), self._st), "fill")

    def axpby(self, x, y, a, b):
        """y = a*x + b*y"""
        _lib.check(self.lib.eg_axpby(_p(x), _p(y), x.numel(), float(a), float(b), self._st), "axpby")

    # ---- WGAN-GP --------------------------------------------------------------------------------
    def gp_interpolate(self, real, fake, alpha, xhat):
        B = real.shape[0]
        _lib.check(self.lib.eg_gp_interpolate(_p(real), _p(fake), _p(alpha), _p(xhat), B, real.numel() // B, self._st), "gp_interpolate")

    def gp_seed(self, d, dd):
        _lib.check(self.lib.eg_gp_seed(_p(d), _p(dd), d.numel(), self._st), "gp_seed")

    def gp_penalty(self, g, gbar, norms, loss, weight, inv_global_batch):
        B = g.shape[0]
        _lib.check(self.lib.eg_gp_penalty(_p(g), _p(gbar), _p(norms), _p(loss), B, g.numel() // B, float(weight), float(inv_global_batch), self._st), "gp_penalty")

    def gp_seed_bwd(self, d, ddbar, dbar):
        _lib.check(self.lib.eg_gp_seed_bwd(_p(d), _p(ddbar), _p(dbar), d.numel(), self._st), "gp_seed_bwd")

    def sum_scaled(self, x, scale, out, accumulate=False):
        _lib.check(self.lib.eg_sum_scaled(_p(x), x.numel(), float(scale), _p(out), int(accumulate), self._st), "sum_scaled")

    # ---- encoder pieces ---------------------------------------------------------------------------
    def reflect_pad_fwd(self, x, y, p):
        N, H, W, Cn = x.shape
        _lib.check(self.lib.eg_reflect_pad_fwd(_p(x), _p(y), N, H, W, Cn, p, self._st), "reflect_pad_fwd")

    def reflect_pad_bwd(self, gy, gx, p):
        N, H, W, Cn = gx.shape
        _lib.check(self.lib.eg_reflect_pad_bwd(_p(gy), _p(gx), N, H, W, Cn, p, self._st), "reflect_pad_bwd")

    def addrelu_pool2_fwd(self, a, b, y):
        N, H, W, Cn = a.shape
        _lib.check(self.lib.eg_addrelu_pool2_fwd(_p(a), _p(b), _p(y), N, H, W, Cn, self._st), "addrelu_pool2_fwd")

    def addrelu_pool2_bwd(self, a, b, gy, g):
        N, H, W, Cn = a.shape
        _lib.check(self.lib.eg_addrelu_pool2_bwd(_p(a), _p(b), _p(gy), _p(g), N, H, W, Cn, self._st), "addrelu_pool2_bwd")

    def relu_globalmean_fwd(self, x, y):
        N, P, Cn = self._npc(x)
        _lib.check(self.lib.eg_relu_globalmean_fwd(_p(x), _p(y), N, P, Cn, self._st), "relu_globalmean_fwd")

    def relu_globalmean_bwd(self, x, gy, gx):
        N, P, Cn = self._npc(x)
        _lib.check(self.lib.eg_relu_globalmean_bwd(_p(x), _p(gy), _p(gx), N, P, Cn, self._st), "relu_globalmean_bwd")

    def reparam_fwd(self, mu, ls, eps, z):
        ed = eps if isinstance(eps, torch.Tensor) else None       # device scalar (graph replay) or python float
        _lib.check(self.lib.eg_reparam_fwd(_p(mu), _p(ls), 0.0 if ed is not None else float(eps), _p(ed), _p(z), mu.numel(), self._st), "reparam_fwd")

    def zl1_loss_bwd(self, mu, ls, eps, target, weight, inv_global_count, gmu, gls, loss):
        B, Z = mu.shape
        ed = eps if isinstance(eps, torch.Tensor) else None
        _lib.check(self.lib.eg_zl1_loss_bwd(_p(mu), _p(ls), 0.0 if ed is not None else float(eps), _p(ed), _p(target), target.shape[1], B, Z, float(weight),
                                            float(inv_global_count), _p(gmu), _p(gls), _p(loss), self._st), "zl1_loss_bwd")

    def onehot_concat(self, z, zdim, classes, out):
        _lib.check(self.lib.eg_onehot_concat(_p(z), z.shape[0], zdim, classes, _p(out), self._st), "onehot_concat")

    # ---- classifier pieces ------------------------------------------------------------------------
    def prelu_fwd(self, x, leak, y):
        _lib.check(self.lib.eg_prelu_fwd(_p(x), _p(leak), _p(y), x.numel(), self._st), "prelu_fwd")

    def prelu_bwd(self, x, leak, gy, gx, gleak, accumulate_leak=False, accumulate_gx=False):
        """gx (= or +=) gy * prelu'(x); gleak (= or +=) the leak gradient (either output may be None)"""
        _lib.check(self.lib.eg_prelu_bwd_ex(_p(x), _p(leak), _p(gy), _p(gx), _p(gleak), x.numel(), int(accumulate_leak),
                                            int(accumulate_gx), self._st), "prelu_bwd")

    def prelu_fwd2(self, x, leak, y, leak2, y2):
        """y = prelu(x; leak), y2 = prelu(y; leak2) in one pass"""
        _lib.check(self.lib.eg_prelu_fwd2(_p(x), _p(leak), _p(y), _p(leak2), _p(y2), x.numel(), self._st), "prelu_fwd2")

    def mru_gate_fwd(self, cg, cg_i, ht, img, leak, stats, plus, hin):
        """fused middle of an MRU unit (eg_mru_gate_fwd): cg is overwritten with rgl = lrelu(cg + cg_i)"""
        N, P, Cn = self._npc(cg)
        _lib.check(self.lib.eg_mru_gate_fwd(_p(cg), _p(cg_i), _p(ht), _p(img), _p(leak), _p(stats), _p(plus), _p(hin),
                                            N, P, Cn, self._st), "mru_gate_fwd")

    def mru_gate_bwd(self, plus, g_hin, img, rgl, stats, leak, g_ht, g_img, g_cg, gleak, accumulate_leak=False):
        N, P, Cn = self._npc(plus)
        _lib.check(self.lib.eg_mru_gate_bwd(_p(plus), _p(g_hin), _p(img), _p(rgl), _p(stats), _p(leak), _p(g_ht), _p(g_img),
                                            _p(g_cg), _p(gleak), int(accumulate_leak), N, P, Cn, self._st), "mru_gate_bwd")

    def minmax_fwd(self, x, y, stats):
        N, P, Cn = self._npc(x)
        _lib.check(self.lib.eg_minmax_fwd(_p(x), _p(y), _p(stats), N, P, Cn, self._st), "minmax_fwd")

    def minmax_bwd(self, x, stats, gy, gx):
        N, P, Cn = self._npc(x)
        _lib.check(self.lib.eg_minmax_bwd(_p(x), _p(stats), _p(gy), _p(gx), N, P, Cn, self._st), "minmax_bwd")

    def fma3(self, a, b, c, out):
        _lib.check(self.lib.eg_fma3(_p(a), _p(b), _p(c), _p(out), a.numel(), self._st), "fma3")

    def mul(self, a, b, out):
        _lib.check(self.lib.eg_mul(_p(a), _p(b), _p(out), a.numel(), self._st), "mul")

    def add_pool2_fwd(self, a, b, y, leak=None, y_act=None):
        """y = mean_pool2x2(a + b); with leak / y_act also y_act = prelu(y; leak)"""
        N, H, W, Cn = a.shape
        _lib.check(self.lib.eg_add_pool2_prelu_fwd(_p(a), _p(b), _p(y), _p(leak), _p(y_act), N, H, W, Cn, self._st), "add_pool2_fwd")

    def pool2_bwd(self, gy, gx, accumulate=False):
        N, H, W, Cn = gx.shape
        _lib.check(self.lib.eg_pool2_bwd(_p(gy), _p(gx), N, H, W, Cn, int(accumulate), self._st), "pool2_bwd")

    def globalmean_fwd(self, x, y):
        N, P, Cn = self._npc(x)
        _lib.check(self.lib.eg_globalmean_fwd(_p(x), _p(y), N, P, Cn, self._st), "globalmean_fwd")

    def globalmean_bwd(self, gy, gx):
        N, P, Cn = self._npc(gx)
        _lib.check(self.lib.eg_globalmean_bwd(_p(gy), _p(gx), N, P, Cn, self._st), "globalmean_bwd")

    def sn_ws_floats(self, K, Cn):
        return int(self.lib.eg_spectral_norm_ws_floats(K, Cn))

    def spectral_norm_fwd(self, W, u, Wbar, ws):
        Cn = W.shape[-1]
        _lib.check(self.lib.eg_spectral_norm_fwd(_p(W), _p(u), _p(Wbar), _p(ws), W.numel() // Cn, Cn, self._st), "spectral_norm_fwd")

    def spectral_norm_bwd(self, W, u, ws, Gbar, gW):
        Cn = W.shape[-1]
        _lib.check(self.lib.eg_spectral_norm_bwd(_p(W), _p(u), _p(ws), _p(Gbar), _p(gW), W.numel() // Cn, Cn, self._st), "spectral_norm_bwd")

    def spectral_norm_set(self, items):
        """One table for all spectrally normalised weights of a network (eg_spectral_norm_set_*): `items` = dicts with W,
        u, Wbar, ws and optionally G, gW (backward) and Wa, Wi, Ga, Gi, hd (tensor split along its input-channel axis)."""
        return SpectralNormSet(self, items)

    def softmax_ce_bwd(self, logits, z, label_col, focal, weight, inv_global_batch, glogits, loss):
        B, Cn = logits.shape
        _lib.check(self.lib.eg_softmax_ce_bwd(_p(logits), _p(z), z.shape[1], label_col, B, Cn, int(focal), float(weight),
                                              float(inv_global_batch), _p(glogits), _p(loss), self._st), "softmax_ce_bwd")

    def copy_cslice(self, src, src_c0, dst, dst_c0, width):
        """dst[..., dst_c0:dst_c0+width] = src[..., src_c0:src_c0+width]  (channel slices / tf.concat on channels)"""
        Cs, Cd = src.shape[-1], dst.shape[-1]
        rows = src.numel() // Cs
        _lib.check(self.lib.eg_copy2d(src.data_ptr() + 4 * src_c0, Cs, dst.data_ptr() + 4 * dst_c0, Cd, rows, width, self._st), "copy2d")

    # ---- optimizer ----------------------------------------------------------------------------------
    def rmsprop(self, var, grad, ms, lr, decay=0.9, eps=1e-10):
        _lib.check(self.lib.eg_rmsprop(_p(var), _p(grad), _p(ms), var.numel(), float(lr), float(decay), float(eps), self._st), "rmsprop")
