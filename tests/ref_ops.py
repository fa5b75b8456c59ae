"""TEST INFRASTRUCTURE: a torch-CPU re-implementation of the `edgegan_b200.ops.DeviceOps` method surface.

It lets the `-m "not gpu"` tests run the product's host-side logic (the hand-derived backward and
WGAN-GP double-backward orchestration in edgegan_b200/models/*.py) on the CPU and compare it with
the oracle's autograd, and it is the per-operator reference the `-m gpu` tests check each CUDA kernel
against.  It is never imported by the product package.

Each method follows the contract written in include/edgegan_b200.h, in plain torch (fp64 by default).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as TF

IN_EPS = 1e-5
BN_EPS = 1e-5


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def act_fwd(act, x):
    if act in (None, "none"):
        return x
    if act == "relu":
        return torch.where(x > 0, x, torch.zeros_like(x))
    if act == "lrelu":
        return torch.where(x >= 0, x, 0.2 * x)
    if act == "tanh":
        return torch.tanh(x)
    if act == "sigmoid":
        return torch.sigmoid(x)
    if act == "lrelu2":
        return torch.where(x > 0, x, 0.2 * x)
    raise ValueError(act)


def act_grad(act, x):
    if act in (None, "none"):
        return torch.ones_like(x)
    if act == "relu":
        return (x > 0).to(x.dtype)
    if act == "lrelu":
        return torch.where(x >= 0, torch.ones_like(x), torch.full_like(x, 0.2))
    if act == "tanh":
        return 1 - torch.tanh(x) ** 2
    if act == "sigmoid":
        s = torch.sigmoid(x)
        return s * (1 - s)
    if act == "lrelu2":
        return torch.where(x > 0, torch.ones_like(x), torch.full_like(x, 0.2))
    raise ValueError(act)


def _pads(H, OH, k, stride, pt):
    """(before, after) zero padding implied by the conv descriptor (after may be negative = crop)."""
    return pt, (OH - 1) * stride + k - H - pt


def conv_fwd_ref(x, w, bias, out_hw, stride, pad):
    pt, pl = (pad, pad) if isinstance(pad, int) else pad
    N, H, W, Ci = x.shape
    KH, KW = w.shape[0], w.shape[1]
    OH, OW = out_hw
    pb = _pads(H, OH, KH, stride, pt)[1]
    pr = _pads(W, OW, KW, stride, pl)[1]
    xi = TF.pad(_nchw(x), (pl, max(pr, 0), pt, max(pb, 0)))
    y = TF.conv2d(xi, w.permute(3, 2, 0, 1), stride=stride)[:, :, :OH, :OW]
    y = _nhwc(y)
    if bias is not None:
        y = y + bias
    return y


class RefOps:
    def __init__(self, dtype=torch.float64):
        self.dtype = dtype
        self.device = torch.device("cpu")
        self._bufs = {}
        self.launches = 0
        self.sm_count = 148

    # ---- memory ---------------------------------------------------------------------------------
    def empty(self, shape):
        return torch.full(tuple(shape), float("nan"), dtype=self.dtype)

    def zeros(self, shape):
        return torch.zeros(tuple(shape), dtype=self.dtype)

    def buf(self, key, shape):
        shape = tuple(int(s) for s in shape)
        t = self._bufs.get(key)
        if t is None or t.shape != shape:
            t = torch.full(shape, float("nan"), dtype=self.dtype)
            self._bufs[key] = t
        return t

    def from_numpy(self, a):
        return torch.tensor(np.asarray(a), dtype=self.dtype)

    def to_numpy(self, t):
        return t.detach().numpy().copy()

    def upload(self, dst, src_host):
        dst.copy_(torch.as_tensor(np.asarray(src_host), dtype=self.dtype).reshape(dst.shape))

    def set_default_algo(self, algo):
        pass

    # ---- conv -------------------------------------------------------------------------------------
    @staticmethod
    def _epilogue(v, act, mask):
        if mask is not None:
            return v * act_grad(act, mask)
        return act_fwd(act, v)

    @staticmethod
    def in_filter_set(shape):
        return len(shape) == 4

    def filter_set(self, filters):
        return None          # the CPU operator set reads the filters as they are

    def conv_fwd(self, x, w, bias, y, stride, pad, algo=None, act=None, mask=None):
        y.copy_(self._epilogue(conv_fwd_ref(x, w, bias, (y.shape[1], y.shape[2]), stride, pad), act, mask))

    def conv_bwd_data(self, dy, w, bias, dx, stride, pad, algo=None, act=None, mask=None):
        xz = torch.zeros(dx.shape, dtype=self.dtype, requires_grad=True)
        y = conv_fwd_ref(xz, w, None, (dy.shape[1], dy.shape[2]), stride, pad)
        (g,) = torch.autograd.grad(y, xz, dy)
        if bias is not None:
            g = g + bias
        dx.copy_(self._epilogue(g, act, mask))

    def conv_bwd_weight(self, x, dy, dw, stride, pad, accumulate=False, algo=None):
        wz = torch.zeros(dw.shape, dtype=self.dtype, requires_grad=True)
        y = conv_fwd_ref(x, wz, None, (dy.shape[1], dy.shape[2]), stride, pad)
        (g,) = torch.autograd.grad(y, wz, dy)
        if accumulate:
            dw.add_(g)
        else:
            dw.copy_(g)

    def bias_grad(self, dy, db, accumulate=False):
        g = dy.reshape(-1, dy.shape[-1]).sum(0).reshape(db.shape)
        db.add_(g) if accumulate else db.copy_(g)

    # ---- instance norm ------------------------------------------------------------------------------
    @staticmethod
    def _in(x, mean=None, sd=None):
        N, C = x.shape[0], x.shape[-1]
        x3 = x.reshape(N, -1, C)
        if mean is None:
            mean = x3.mean(1, keepdim=True)
            sd = torch.sqrt(((x3 - mean) ** 2).mean(1, keepdim=True))
        return x3, mean, sd

    def instnorm_fwd(self, x, y, stats, act):
        x3, mean, sd = self._in(x)
        y.copy_(act_fwd(act, (x3 - mean) / (sd + IN_EPS)).reshape(x.shape))
        stats.copy_(torch.stack([mean[:, 0, :], sd[:, 0, :]], dim=-1))

    def _in_autograd(self, x, act, gy):
        """first-order backward as an autograd graph (used for bwd and, differentiated again, bwd2)."""
        x3 = x.reshape(x.shape[0], -1, x.shape[-1])
        mean = x3.mean(1, keepdim=True)
        c = x3 - mean
        sd = torch.sqrt((c ** 2).mean(1, keepdim=True))
        r = 1 / (sd + IN_EPS)
        gn = gy.reshape(x3.shape) * act_grad(act, (c * r).detach())
        gx = r * (gn - gn.mean(1, keepdim=True)) - c * r * r / sd * (gn * c).mean(1, keepdim=True)
        return gx.reshape(x.shape)

    def instnorm_bwd(self, x, stats, gy, addend, gx, act):
        g = self._in_autograd(x, act, gy)
        if addend is not None:
            g = g + addend
        gx.copy_(g)

    def instnorm_bwd2(self, x, stats, gy, t, out_gy, out_x, act):
        xr = x.detach().clone().requires_grad_(True)
        gr = gy.detach().clone().requires_grad_(True)
        gx = self._in_autograd(xr, act, gr)
        a, b = torch.autograd.grad(gx, [gr, xr], t)
        out_gy.copy_(a)
        out_x.copy_(b)

    def act_fwd(self, x, y, act):
        y.copy_(act_fwd(act, x).reshape(y.shape))

    def act_bwd(self, x_pre, gy, gx, act):
        gx.copy_((gy.reshape(x_pre.shape) * act_grad(act, x_pre)).reshape(gx.shape))

    # ---- batch norm -----------------------------------------------------------------------------------
    def bn_stats(self, x, sums):
        C = x.shape[-1]
        x2 = x.reshape(-1, C)
        sums.copy_(torch.cat([x2.sum(0), (x2 * x2).sum(0)]))

    @staticmethod
    def _bn(x, sums, count):
        C = x.shape[-1]
        mu = sums[:C] / count
        var = (sums[C:] / count - mu * mu).clamp_min(0)
        rstd = 1 / torch.sqrt(var + BN_EPS)
        return (x.reshape(-1, C) - mu) * rstd, rstd

    def bn_apply(self, x, sums, count, gamma, beta, y, act):
        xh, _ = self._bn(x, sums, count)
        y.copy_(act_fwd(act, gamma * xh + beta).reshape(y.shape))

    def bn_bwd_reduce(self, x, sums, count, gamma, beta, gy, red, act):
        xh, _ = self._bn(x, sums, count)
        gp = gy.reshape(xh.shape) * act_grad(act, gamma * xh + beta)
        red.copy_(torch.cat([gp.sum(0), (gp * xh).sum(0)]))

    def bn_bwd_apply(self, x, sums, count, gamma, beta, gy, red, gx, act):
        C = x.shape[-1]
        xh, rstd = self._bn(x, sums, count)
        gp = gy.reshape(xh.shape) * act_grad(act, gamma * xh + beta)
        gx.copy_((gamma * rstd * (gp - red[:C] / count - xh * red[C:] / count)).reshape(gx.shape))

    # ---- head -------------------------------------------------------------------------------------------
    def rowdot_fwd(self, h, w, bias, d):
        v = h.reshape(h.shape[0], -1) @ w.reshape(-1)
        d.copy_(v + (bias.reshape(()) if bias is not None else 0))

    def rowdot_bwd_input(self, gd, w, gh):
        gh.copy_((gd.reshape(-1, 1) * w.reshape(1, -1)).reshape(gh.shape))

    def rowdot_bwd_weight(self, gd, h, gw, gb, accumulate=False):
        g = (gd.reshape(1, -1) @ h.reshape(h.shape[0], -1)).reshape(gw.shape)
        gw.add_(g) if accumulate else gw.copy_(g)
        if gb is not None:
            s = gd.sum().reshape(gb.shape)
            gb.add_(s) if accumulate else gb.copy_(s)

    # ---- resize / slices ----------------------------------------------------------------------------------
    @staticmethod
    def _up2(t, axis):
        n = t.shape[axis]
        idx = torch.arange(n)

        def tk(o):
            return t.index_select(axis, (idx + o).clamp(0, n - 1))
        odd = -0.09375 * tk(-1) + 0.59375 * tk(0) + 0.59375 * tk(1) - 0.09375 * tk(2)
        st = torch.stack([t, odd], dim=axis + 1)
        shp = list(t.shape)
        shp[axis] = 2 * n
        return st.reshape(shp)

    def bicubic_up2_fwd(self, x, y):
        y.copy_(self._up2(self._up2(x, 1), 2))

    def bicubic_up2_bwd(self, gy, gx):
        xz = torch.zeros(gx.shape, dtype=self.dtype, requires_grad=True)
        (g,) = torch.autograd.grad(self._up2(self._up2(xz, 1), 2), xz, gy)
        gx.copy_(g)

    def copy_wslice(self, src, src_w0, dst, dst_w0, width):
        dst[:, :, dst_w0:dst_w0 + width, :] = src[:, :, src_w0:src_w0 + width, :]

    def copy2d(self, src, src_off, src_stride, dst, dst_off, dst_stride, rows, cols):
        sf, df = src.reshape(-1), dst.reshape(-1)
        for r in range(rows):
            df[dst_off + r * dst_stride: dst_off + r * dst_stride + cols] = sf[src_off + r * src_stride: src_off + r * src_stride + cols]

    def copy(self, src, dst):
        dst.copy_(src.reshape(dst.shape))

    def fill(self, dst, value):
        dst.fill_(value)

    def axpby(self, x, y, a, b):
        y.copy_(a * x.reshape(y.shape) + (b * y if b != 0 else 0))

    # ---- WGAN-GP ----------------------------------------------------------------------------------------------
    def gp_interpolate(self, real, fake, alpha, xhat):
        a = alpha.reshape(-1, *([1] * (real.dim() - 1)))
        xhat.copy_(real + a * (fake - real))

    def gp_seed(self, d, dd):
        s = torch.sigmoid(d)
        dd.copy_(1 + s * (1 - s))

    def gp_penalty(self, g, gbar, norms, loss, weight, inv_global_batch):
        B = g.shape[0]
        n = torch.sqrt((g.reshape(B, -1) ** 2).sum(1))
        norms.copy_(n)
        coef = weight * 2 * (n - 1) * inv_global_batch / n
        gbar.copy_(coef.reshape(-1, *([1] * (g.dim() - 1))) * g)
        loss.add_(weight * ((n - 1) ** 2).sum() * inv_global_batch)

    def gp_seed_bwd(self, d, ddbar, dbar):
        s = torch.sigmoid(d)
        dbar.copy_(ddbar * s * (1 - s) * (1 - 2 * s))

    def sum_scaled(self, x, scale, out, accumulate=False):
        v = scale * x.sum()
        out.add_(v) if accumulate else out.fill_(float(v))

    # ---- encoder pieces ------------------------------------------------------------------------------------------
    def reflect_pad_fwd(self, x, y, p):
        y.copy_(_nhwc(TF.pad(_nchw(x), (p, p, p, p), mode="reflect")))

    def reflect_pad_bwd(self, gy, gx, p):
        xz = torch.zeros(gx.shape, dtype=self.dtype, requires_grad=True)
        (g,) = torch.autograd.grad(_nhwc(TF.pad(_nchw(xz), (p, p, p, p), mode="reflect")), xz, gy)
        gx.copy_(g)

    def addrelu_pool2_fwd(self, a, b, y):
        v = torch.relu(a + b) if b is not None else torch.relu(a)
        y.copy_(_nhwc(TF.avg_pool2d(_nchw(v), 2, 2)))

    def addrelu_pool2_bwd(self, a, b, gy, g):
        v = a + b if b is not None else a
        up = gy.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        g.copy_((v > 0).to(self.dtype) * up * 0.25)

    def relu_globalmean_fwd(self, x, y):
        N, C = x.shape[0], x.shape[-1]
        y.copy_(torch.relu(x).reshape(N, -1, C).mean(1))

    def relu_globalmean_bwd(self, x, gy, gx):
        N, C = x.shape[0], x.shape[-1]
        P = x.numel() // (N * C)
        g = (x.reshape(N, P, C) > 0).to(self.dtype) * gy.reshape(N, 1, C) / P
        gx.copy_(g.reshape(gx.shape))

    def reparam_fwd(self, mu, ls, eps, z):
        eps = eps.reshape(()) if isinstance(eps, torch.Tensor) else eps
        z.copy_(mu + eps * torch.exp(ls))

    def zl1_loss_bwd(self, mu, ls, eps, target, weight, inv_global_count, gmu, gls, loss):
        Z = mu.shape[1]
        e = torch.exp(ls)
        eps = eps.reshape(()) if isinstance(eps, torch.Tensor) else eps
        diff = target[:, :Z] - (mu + eps * e)
        loss.add_(weight * inv_global_count * diff.abs().sum())
        gz = -weight * inv_global_count * torch.sign(diff)
        gmu.copy_(gz)
        gls.copy_(gz * eps * e)

    def onehot_concat(self, z, zdim, classes, out):
        lab = z[:, zdim].to(torch.int64)
        out.copy_(torch.cat([z[:, :zdim], TF.one_hot(lab, classes).to(self.dtype)], dim=1))

    # ---- classifier pieces -------------------------------------------------------------------------------------
    def prelu_fwd(self, x, leak, y):
        a = leak.reshape(()) * x
        y.copy_(torch.where(a >= x, a, x))

    def prelu_fwd2(self, x, leak, y, leak2, y2):
        self.prelu_fwd(x, leak, y)
        self.prelu_fwd(y, leak2, y2)

    def mru_gate_fwd(self, cg, cg_i, ht, img, leak, stats, plus, hin):
        """composition of the separate reference ops (axpby, lrelu2, minmax, fma3, prelu)"""
        rgl = act_fwd("lrelu2", cg + cg_i)
        cg.copy_(rgl)
        rg = torch.empty_like(rgl)
        mm = torch.empty(stats.shape[:-1] + (2,), dtype=self.dtype)
        self.minmax_fwd(rgl, rg, mm)
        x3 = rgl.reshape(rgl.shape[0], -1, rgl.shape[-1])
        cnt_min = (x3 == mm[:, None, :, 0]).to(self.dtype).sum(1)
        cnt_max = (x3 == mm[:, None, :, 1]).to(self.dtype).sum(1)
        stats.copy_(torch.stack([mm[..., 0], mm[..., 1], cnt_min, cnt_max], dim=-1))
        self.fma3(ht, rg, img, plus)
        self.prelu_fwd(plus, leak, hin)

    def mru_gate_bwd(self, plus, g_hin, img, rgl, stats, leak, g_ht, g_img, g_cg, gleak, accumulate_leak=False):
        g_plus = torch.empty_like(plus)
        self.prelu_bwd(plus, leak, g_hin, g_plus, gleak, accumulate_leak)
        g_ht.add_(g_plus)
        rg = torch.empty_like(rgl)
        self.minmax_fwd(rgl, rg, torch.empty(stats.shape[:-1] + (2,), dtype=self.dtype))
        g_img.copy_(g_plus * rg)
        g_rgl = torch.empty_like(rgl)
        self.minmax_bwd(rgl, stats[..., :2], g_plus * img, g_rgl)
        g_cg.copy_(g_rgl * act_grad("lrelu2", rgl))

    def prelu_bwd(self, x, leak, gy, gx, gleak, accumulate_leak=False, accumulate_gx=False):
        first = (leak.reshape(()) * x >= x)
        if gx is not None:
            o = torch.where(first, gy * leak.reshape(()), gy)
            gx.add_(o) if accumulate_gx else gx.copy_(o)
        if gleak is not None:
            v = (gy * x * first.to(self.dtype)).sum().reshape(gleak.shape)
            gleak.add_(v) if accumulate_leak else gleak.copy_(v)

    @staticmethod
    def _mm(x):
        N, C = x.shape[0], x.shape[-1]
        x3 = x.reshape(N, -1, C)
        return x3, x3.amin(1, keepdim=True), x3.amax(1, keepdim=True)

    def minmax_fwd(self, x, y, stats):
        x3, mn, mx = self._mm(x)
        y.copy_(((x3 - mn) / (mx - mn)).reshape(x.shape))
        stats.copy_(torch.stack([mn[:, 0, :], mx[:, 0, :]], dim=-1))

    def minmax_bwd(self, x, stats, gy, gx):
        xr = x.detach().clone().requires_grad_(True)
        x3 = xr.reshape(x.shape[0], -1, x.shape[-1])
        mn, mx = x3.amin(1, keepdim=True), x3.amax(1, keepdim=True)      # amin/amax split the gradient between ties
        yv = ((x3 - mn) / (mx - mn)).reshape(x.shape)
        (g,) = torch.autograd.grad(yv, xr, gy)
        gx.copy_(g)

    def fma3(self, a, b, c, out):
        out.copy_(a + b * c)

    def mul(self, a, b, out):
        out.copy_(a * b)

    def add_pool2_fwd(self, a, b, y, leak=None, y_act=None):
        v = a + b if b is not None else a
        y.copy_(_nhwc(TF.avg_pool2d(_nchw(v), 2, 2)))
        if y_act is not None:
            self.prelu_fwd(y, leak, y_act)

    def pool2_bwd(self, gy, gx, accumulate=False):
        up = gy.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2) * 0.25
        gx.add_(up) if accumulate else gx.copy_(up)

    def globalmean_fwd(self, x, y):
        N, C = x.shape[0], x.shape[-1]
        y.copy_(x.reshape(N, -1, C).mean(1))

    def globalmean_bwd(self, gy, gx):
        N, C = gx.shape[0], gx.shape[-1]
        P = gx.numel() // (N * C)
        gx.copy_((gy.reshape(N, 1, C) / P).expand(N, P, C).reshape(gx.shape))

    def sn_ws_floats(self, K, C):
        return 2 * K + C + 8

    @staticmethod
    def _sn(W, u):
        Wm = W.reshape(-1, W.shape[-1])

        def l2n(t):
            return t / (torch.sum(t ** 2) ** 0.5 + 1e-12)
        v1 = l2n(u.reshape(1, -1) @ Wm.t())
        u1 = l2n(v1 @ Wm)
        sigma = (v1 @ Wm @ u1.t())[0, 0]
        return (Wm / sigma).reshape(W.shape)

    def spectral_norm_fwd(self, W, u, Wbar, ws):
        Wbar.copy_(self._sn(W, u))

    def spectral_norm_bwd(self, W, u, ws, Gbar, gW):
        Wr = W.detach().clone().requires_grad_(True)
        (g,) = torch.autograd.grad(self._sn(Wr, u), Wr, Gbar)
        gW.copy_(g)

    def spectral_norm_set(self, items):
        ops = self

        class Set:
            def fwd(self):
                for it in items:
                    ops.spectral_norm_fwd(it["W"], it["u"], it["Wbar"], it["ws"])
                    if it.get("Wa") is not None:
                        hd = it["hd"]
                        it["Wa"].copy_(it["Wbar"][:, :, :hd]), it["Wi"].copy_(it["Wbar"][:, :, hd:])

            def bwd(self):
                for it in items:
                    G = it["G"] if it.get("Ga") is None else torch.cat([it["Ga"], it["Gi"]], dim=2)
                    ops.spectral_norm_bwd(it["W"], it["u"], it["ws"], G, it["gW"])

            def close(self):
                pass
        return Set()

    def softmax_ce_bwd(self, logits, z, label_col, focal, weight, inv_global_batch, glogits, loss):
        lr_ = logits.detach().clone().requires_grad_(True)
        lab = z[:, label_col].to(torch.int64)
        ce = TF.cross_entropy(lr_, lab, reduction="none")
        if focal:
            py = torch.softmax(lr_, dim=1).gather(1, lab.view(-1, 1)).squeeze(1)
            per = (1 - py) ** 2 * ce
        else:
            per = ce
        total = weight * inv_global_batch * per.sum()
        (g,) = torch.autograd.grad(total, lr_)
        glogits.copy_(g)
        loss.add_(total.detach())

    def copy_cslice(self, src, src_c0, dst, dst_c0, width):
        dst[..., dst_c0:dst_c0 + width] = src[..., src_c0:src_c0 + width]

    def rmsprop(self, var, grad, ms, lr, decay=0.9, eps=1e-10):
        ms.copy_(decay * ms + (1 - decay) * grad * grad)
        var.sub_(lr * grad / torch.sqrt(ms + eps))
