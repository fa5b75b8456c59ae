"""CPU oracle for the EdgeGAN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``edgegan_b200``) never does.

PARITY UNPINNED: the reference (sysu-imsl/EdgeGAN) is a TensorFlow-1.14 graph program whose
arithmetic lives in ``tensorflow_gpu==1.14.0`` (reference ``requirements.txt:3``), which cannot be
installed here (python 3.12, no network), and the reference ships no tests, golden vectors or
fixtures for this path.  This file is therefore a *restatement* in torch-CPU (fp32, or fp64 for
finite-difference checks) of exactly the graph the reference builds, op by op, with the TF-1.14
op semantics written down in SURVEY.md appendix A.  Its self-checks (tests/test_oracle_*.py) are:
independent numpy scatter/loop definitions of conv-transpose, bicubic, instance-norm and RMSProp,
fp64 finite differences for every hand-derived backward used by the CUDA path, and invariants.

All tensors are NHWC float arrays at the API (the classifier is NCHW at *its* API, as in the
reference); all randomness (weights, z, alpha, eps) is an explicit input.

Reference citations are relative to /root/reference/edgegan/.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as TF

# --------------------------------------------------------------------------------------
# configuration (reference train.py:14-74 defaults; models/edgegan.py:46-48 constants)
# --------------------------------------------------------------------------------------


@dataclass
class Config:
    batch_size: int = 64
    output_height: int = 64          # train.py:24
    output_width: int = 128          # train.py:26 (pair width; each generator makes width/2)
    multiclasses: bool = True        # train.py:45
    num_classes: int = 14            # train.py:46
    z_dim: int = 100                 # train.py:74
    gf_dim: int = 64                 # edgegan.py:47
    df_dim: int = 64
    c_dim: int = 3
    learning_rate: float = 2e-4      # train.py:18
    lambda_gp: float = 10.0          # train.py:56
    stage1_zl_loss: float = 10.0     # train.py:44
    image_dis_size: int = 128        # train.py:64
    edge_dis_size: int = 128         # train.py:67
    joint_dweight: float = 1.0
    image_dweight: float = 1.0
    edge_dweight: float = 1.0
    use_image_discriminator: bool = True
    use_edge_discriminator: bool = True

    @property
    def g_in_dim(self):
        return self.z_dim + (self.num_classes if self.multiclasses else 0)


# --------------------------------------------------------------------------------------
# variable store (names/shapes: SURVEY.md appendix B; initialisers: appendix A14)
# --------------------------------------------------------------------------------------

def _trunc_normal(rs, shape, std):
    """tf.truncated_normal_initializer: re-draw samples beyond 2 sigma (conv.py:21)."""
    x = rs.normal(0.0, std, size=shape)
    bad = np.abs(x) > 2 * std
    while bad.any():
        x[bad] = rs.normal(0.0, std, size=int(bad.sum()))
        bad = np.abs(x) > 2 * std
    return x.astype(np.float32)


def _normal(rs, shape, std=0.02):
    return rs.normal(0.0, std, size=shape).astype(np.float32)


def generator_variables(cfg: Config, name: str, rs) -> "OrderedDict[str, np.ndarray]":
    """generator.py:35-74 + linear.py:13-27 + normalization.py:20-25 + conv.py:41-52."""
    v = OrderedDict()
    gf = cfg.gf_dim
    sh, sw = cfg.output_height // 16, cfg.output_width // 2 // 16
    v[f"{name}/g_lin_0/Matrix"] = _normal(rs, (cfg.g_in_dim, gf * 8 * sh * sw))
    v[f"{name}/g_lin_0/bias"] = np.zeros(gf * 8 * sh * sw, np.float32)
    v[f"{name}/batch_norm/BatchNorm/beta"] = np.zeros(gf * 8, np.float32)
    v[f"{name}/batch_norm/BatchNorm/gamma"] = np.ones(gf * 8, np.float32)
    chans = [gf * 8, gf * 4, gf * 2, gf, cfg.c_dim]
    for i in range(1, 5):
        v[f"{name}/g_dconv_{i}/deconv2d/w"] = _normal(rs, (5, 5, chans[i], chans[i - 1]))
        v[f"{name}/g_dconv_{i}/deconv2d/b"] = np.zeros(chans[i], np.float32)
    return v


def discriminator_variables(cfg: Config, name: str, in_h: int, in_w: int, rs):
    """discriminator.py:58-81 (layers d_conv_0,1,3,4 -- there is no d_conv_2)."""
    v = OrderedDict()
    df = cfg.df_dim
    chans = [cfg.c_dim, df, df * 2, df * 4, df * 8]
    for i, lname in enumerate(["d_conv_0", "d_conv_1", "d_conv_3", "d_conv_4"]):
        v[f"{name}/{lname}/conv2d/w"] = _trunc_normal(rs, (4, 4, chans[i], chans[i + 1]), 0.02)
    feat = (in_h // 16) * (in_w // 16) * df * 8
    v[f"{name}/d_linear_5/Matrix"] = _normal(rs, (feat, 1))
    v[f"{name}/d_linear_5/bias"] = np.zeros(1, np.float32)
    return v


def encoder_variables(cfg: Config, name: str, rs):
    """encoder.py:54-84, conv.py:70-85, linear.py:79-92."""
    v = OrderedDict()
    v[f"{name}/e_resnet_64_0/conv2d/w"] = _trunc_normal(rs, (4, 4, cfg.c_dim, 64), 0.02)
    v[f"{name}/e_resnet_64_0/conv2d/b"] = np.zeros(64, np.float32)
    cin = 64
    nf = [128, 256, 512, 512] + ([512] if cfg.output_height == 256 else [])
    for i, n in enumerate(nf):
        p = f"{name}/e_resnet_{n}_{i + 1}"
        for sub, k, ci in (("res1", 3, cin), ("res2", 3, n), ("shortcut", 1, cin)):
            v[f"{p}/{sub}/conv2d/w"] = _trunc_normal(rs, (k, k, ci, n), 0.02)
            v[f"{p}/{sub}/conv2d/b"] = np.zeros(n, np.float32)
        cin = n
    for fc in ("FC8_mu", "FC8_sigma"):
        v[f"{name}/{fc}/w"] = _normal(rs, (cin, cfg.z_dim))
        v[f"{name}/{fc}/b"] = np.zeros(cfg.z_dim, np.float32)
    return v


CLASSIFIER_UNITS = ((8, 128), (128, 256), (256, 512), (512, 768))   # (hidden_depth, filter_depth)


def classifier_variables(cfg: Config, name: str, rs):
    """classifier.py:12-119 + conv.py:133-357 + linear.py:34-76 + normalization.py:38-76.

    Returns (trainables, sn_u) -- `u` vectors are non-trainable and never updated."""
    v, u = OrderedDict(), OrderedDict()

    def conv(scope, k, ci, co, bias_init=0.0, prelu=False):
        v[f"{scope}/weights"] = _normal(rs, (k, k, ci, co))
        v[f"{scope}/biases"] = np.full((1, co, 1, 1), bias_init, np.float32)
        u[f"{scope}/u"] = _trunc_normal(rs, (1, co), 1.0)
        if prelu:
            v[f"{scope}/prelu/param"] = np.float32(0.2).reshape(())

    conv(f"{name}/Conv", 7, cfg.c_dim, 8, prelu=True)
    for t, (hd, fd) in enumerate(CLASSIFIER_UNITS, start=1):
        p = f"{name}/mru_conv_unit_t_{t}_layer_0"
        v[f"{p}/norm_activation_in/prelu/param"] = np.float32(0.2).reshape(())
        conv(f"{p}/update_gate", 3, hd + cfg.c_dim, hd, bias_init=0.5)
        conv(f"{p}/Conv", 3, cfg.c_dim, hd)
        v[f"{p}/norm_activation_merge_1/prelu/param"] = np.float32(0.2).reshape(())
        conv(f"{p}/Conv_1", 3, hd, fd, prelu=True)
        conv(f"{p}/Conv_2", 3, fd, fd)
        conv(f"{p}/Conv_3", 1, hd, fd)
    v[f"{name}/mru_conv_unit_last_norm/prelu/param"] = np.float32(0.2).reshape(())
    conv(f"{name}/Conv_1", 1, 768, 1)      # the discarded "disc" head (classifier.py:107-109)
    lim = math.sqrt(6.0 / (768 + cfg.num_classes))     # xavier uniform (linear.py:36)
    v[f"{name}/fully_connected/weights"] = rs.uniform(-lim, lim, (768, cfg.num_classes)).astype(np.float32)
    v[f"{name}/fully_connected/biases"] = np.zeros(cfg.num_classes, np.float32)
    u[f"{name}/fully_connected/u"] = _trunc_normal(rs, (1, cfg.num_classes), 1.0)
    return v, u


def init_variables(cfg: Config, seed: int = 0):
    """All trainables of the training graph (edgegan.py:132-177) + the frozen SN `u` vectors."""
    rs = np.random.RandomState(seed)
    v = OrderedDict()
    h, w = cfg.output_height, cfg.output_width
    v.update(generator_variables(cfg, "G1", rs))
    v.update(generator_variables(cfg, "G2", rs))
    v.update(discriminator_variables(cfg, "D", h, w, rs))
    v.update(discriminator_variables(cfg, "D_patch2", cfg.image_dis_size, cfg.image_dis_size, rs))
    v.update(discriminator_variables(cfg, "D_patch3", cfg.edge_dis_size, cfg.edge_dis_size, rs))
    v.update(encoder_variables(cfg, "E", rs))
    u = OrderedDict()
    if cfg.multiclasses:
        cv, u = classifier_variables(cfg, "D2", rs)
        v.update(cv)
    return v, u


def scope_vars(v, scope):
    return OrderedDict((k, t) for k, t in v.items() if k.startswith(scope + "/"))


# --------------------------------------------------------------------------------------
# ops (each follows one reference function + the TF-1.14 semantics of SURVEY appendix A)
# --------------------------------------------------------------------------------------

def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def conv2d(x, w, b=None, stride=2, pad="SAME"):
    """nn/modules/conv.py:13-36.  x NHWC, w [kh,kw,Cin,Cout] (A1, A7)."""
    k = w.shape[0]
    wt = w.permute(3, 2, 0, 1)
    xi = _nchw(x)
    if pad == "REFLECT":
        p = (k - 1) // 2
        if p:
            xi = TF.pad(xi, (p, p, p, p), mode="reflect")
        y = TF.conv2d(xi, wt, stride=stride)
    elif pad == "SAME":
        H, W = x.shape[1], x.shape[2]
        oh, ow = -(-H // stride), -(-W // stride)
        ph = max((oh - 1) * stride + k - H, 0)
        pw = max((ow - 1) * stride + k - W, 0)
        xi = TF.pad(xi, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        y = TF.conv2d(xi, wt, stride=stride)
    else:
        assert pad == "VALID"
        y = TF.conv2d(xi, wt, stride=stride)
    y = _nhwc(y)
    if b is not None:
        y = y + b
    return y


def deconv2d(x, w, b):
    """conv.py:39-58: tf.nn.conv2d_transpose SAME s2 k5, w [kh,kw,Cout,Cin]; oh = 2*ih + kh - 1 (A2)."""
    n_h, n_w = x.shape[1], x.shape[2]
    wt = w.permute(3, 2, 0, 1)               # torch conv_transpose weight: [Cin, Cout, kh, kw]
    y = TF.conv_transpose2d(_nchw(x), wt, stride=2, padding=1)[..., : 2 * n_h, : 2 * n_w]
    return _nhwc(y) + b


def instance_norm(x, eps=1e-5):
    """normalization.py:13-18: (x-mean)/(sqrt(var)+eps), biased var over H,W (A3)."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    return (x - mean) / (torch.sqrt(var) + eps)


def batch_norm(x, gamma, beta, eps=1e-5):
    """normalization.py:19-25: contrib fused BN, is_training=True always (A4)."""
    mean = x.mean(dim=(0, 1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(0, 1, 2), keepdim=True)
    return gamma * (x - mean) / torch.sqrt(var + eps) + beta


def act_lrelu_block(x):
    """activation.py:9: tf.maximum(x, 0.2x) -- gradient 1 at x == 0 (A8)."""
    return torch.where(x >= 0, x, 0.2 * x)


def lrelu(x, leak=0.2):
    """activation.py:30-32: tf.maximum(leak*x, x) -- gradient `leak` at 0 (A8)."""
    return torch.where(x > 0, x, leak * x)


def prelu(x, leak):
    """activation.py:23-27: tf.maximum(leak*x, x) with a learned scalar leak.
    Written with explicit masks so autograd reproduces TF's tie rule (first arg wins, A8)."""
    first = (leak * x >= x).to(x.dtype)          # tf.maximum(a,b): grad to a where a>=b
    return first * (leak * x) + (1 - first) * x


def bicubic_up2(x, size):
    """edgegan.py:211-213 tf.image.resize_images(method=2) legacy bicubic, A=-0.75 (A5).
    Same-size resize returns the input."""
    H, W = x.shape[1], x.shape[2]
    if H == size and W == size:
        return x
    assert size == 2 * H and size == 2 * W, "oracle covers the 2x case used by the configs"

    def up(t, axis):
        n = t.shape[axis]
        idx = torch.arange(n)

        def tk(o):
            return t.index_select(axis, (idx + o).clamp(0, n - 1))
        odd = -0.09375 * tk(-1) + 0.59375 * tk(0) + 0.59375 * tk(1) - 0.09375 * tk(2)
        st = torch.stack([t, odd], dim=axis + 1)
        shp = list(t.shape)
        shp[axis] = 2 * n
        return st.reshape(shp)
    return up(up(x, 1), 2)


def avg_pool_same(x, k):
    """encoder.py:68,70 tf.nn.avg_pool SAME, stride k: divisor = #in-bounds elements (A6)."""
    H, W = x.shape[1], x.shape[2]
    oh, ow = -(-H // k), -(-W // k)
    ph, pw = max((oh - 1) * k + k - H, 0), max((ow - 1) * k + k - W, 0)
    xi = TF.pad(_nchw(x), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    ones = TF.pad(torch.ones(1, 1, H, W, dtype=x.dtype), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    s = TF.avg_pool2d(xi, k, k) * (k * k)
    c = TF.avg_pool2d(ones, k, k) * (k * k)
    return _nhwc(s / c)


def mean_pool_nchw(x):
    """pooling.py:4-8."""
    return (x[:, :, ::2, ::2] + x[:, :, 1::2, ::2] + x[:, :, ::2, 1::2] + x[:, :, 1::2, 1::2]) / 4.0


def spectral_normed_weight(W, u):
    """normalization.py:38-76, num_iters=1, no stop_gradient, `u` never updated (A12)."""
    Wm = W.reshape(-1, W.shape[-1])

    def l2n(t):
        return t / (torch.sum(t ** 2) ** 0.5 + 1e-12)
    v1 = l2n(u @ Wm.t())
    u1 = l2n(v1 @ Wm)
    sigma = (v1 @ Wm @ u1.t())[0, 0]
    return (Wm / sigma).reshape(W.shape)


def conv2d2(x, w, b, u, k, act=None, leak=None):
    """conv.py:246-295: NCHW SAME stride-1 conv with SN weights + bias (+activation)."""
    wb = spectral_normed_weight(w, u)
    p = (k - 1) // 2
    y = TF.conv2d(x, wb.permute(3, 2, 0, 1), padding=p) + b
    if act == "prelu":
        y = prelu(y, leak)
    elif act == "lrelu":
        y = lrelu(y)
    return y


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------

def generator(v, name, z):
    """models/generator.py:35-74 (BN at h0 per SURVEY D4; IN + relu on dconv1-3; tanh)."""
    g = lambda s: v[f"{name}/{s}"]
    h = z @ g("g_lin_0/Matrix") + g("g_lin_0/bias")
    c = g("batch_norm/BatchNorm/gamma").shape[0]
    sp = h.shape[1] // c
    s = int(round(math.sqrt(sp)))
    # 64x64 -> 4x4; 128x128 -> 8x8.  generator.py:50 reshape [-1, s_h16, s_w16, 8*gf]
    h = h.reshape(-1, s, sp // s, c)
    h = torch.relu(batch_norm(h, g("batch_norm/BatchNorm/gamma"), g("batch_norm/BatchNorm/beta")))
    for i in (1, 2, 3):
        h = deconv2d(h, g(f"g_dconv_{i}/deconv2d/w"), g(f"g_dconv_{i}/deconv2d/b"))
        h = torch.relu(instance_norm(h))
    h = deconv2d(h, g("g_dconv_4/deconv2d/w"), g("g_dconv_4/deconv2d/b"))
    return torch.tanh(h)


def discriminator(v, name, x):
    """models/discriminator.py:58-81 -> (sigmoid(d), d)."""
    g = lambda s: v[f"{name}/{s}"]
    h = act_lrelu_block(conv2d(x, g("d_conv_0/conv2d/w"), None, 2, "SAME"))
    for l in ("d_conv_1", "d_conv_3", "d_conv_4"):
        h = act_lrelu_block(instance_norm(conv2d(h, g(f"{l}/conv2d/w"), None, 2, "SAME")))
    d = h.reshape(h.shape[0], -1) @ g("d_linear_5/Matrix") + g("d_linear_5/bias")
    return torch.sigmoid(d), d


def encoder(v, name, x, eps):
    """models/encoder.py:54-84; eps is the *scalar* noise -> (z, mu, log_sigma)."""
    g = lambda s: v[f"{name}/{s}"]
    h = torch.relu(conv2d(x, g("e_resnet_64_0/conv2d/w"), g("e_resnet_64_0/conv2d/b"), 2, "SAME"))
    blocks = []
    for k in v:
        part = k.split("/")
        if part[0] == name and part[1].startswith("e_resnet_") and part[1] != "e_resnet_64_0" and part[1] not in blocks:
            blocks.append(part[1])
    for p in sorted(blocks, key=lambda s: int(s.rsplit("_", 1)[1])):
        o = conv2d(h, g(f"{p}/res1/conv2d/w"), g(f"{p}/res1/conv2d/b"), 1, "REFLECT")
        o = torch.relu(instance_norm(o))
        o = conv2d(o, g(f"{p}/res2/conv2d/w"), g(f"{p}/res2/conv2d/b"), 1, "REFLECT")
        o = instance_norm(o)
        sc = conv2d(h, g(f"{p}/shortcut/conv2d/w"), g(f"{p}/shortcut/conv2d/b"), 1, "REFLECT")
        h = avg_pool_same(torch.relu(sc + o), 2)
    h = avg_pool_same(torch.relu(h), 8)
    h = h.reshape(h.shape[0], -1)
    mu = h @ g("FC8_mu/w") + g("FC8_mu/b")
    ls = h @ g("FC8_sigma/w") + g("FC8_sigma/b")
    return mu + eps * torch.exp(ls), mu, ls


def classifier(v, u, name, x_nchw):
    """models/classifier.py:12-119 -> logits [B, num_classes] (x is NCHW)."""
    g = lambda s: v[f"{name}/{s}"]
    pyr = [x_nchw]
    for _ in range(5):
        pyr.append(mean_pool_nchw(pyr[-1]))

    def c2(scope, t, k, act=None):
        leak = v.get(f"{name}/{scope}/prelu/param") if act == "prelu" else None
        return conv2d2(t, g(f"{scope}/weights"), g(f"{scope}/biases"), u[f"{name}/{scope}/u"], k, act, leak)

    ht = c2("Conv", pyr[0], 7, "prelu")
    for t in range(1, 5):
        p = f"mru_conv_unit_t_{t}_layer_0"
        inp = pyr[t - 1]
        hd, fd = CLASSIFIER_UNITS[t - 1]
        full = torch.cat([prelu(ht, g(f"{p}/norm_activation_in/prelu/param")), inp], dim=1)
        rg = c2(f"{p}/update_gate", full, 3, "lrelu")
        mn = rg.amin(dim=(2, 3), keepdim=True)
        mx = rg.amax(dim=(2, 3), keepdim=True)
        rg = (rg - mn) / (mx - mn)
        img_new = c2(f"{p}/Conv", inp, 3)
        ht_in = prelu(ht + rg * img_new, g(f"{p}/norm_activation_merge_1/prelu/param"))
        hn = c2(f"{p}/Conv_1", ht_in, 3, "prelu")
        hn = c2(f"{p}/Conv_2", hn, 3)
        ho = c2(f"{p}/Conv_3", ht, 1) if hd != fd else ht
        ht = mean_pool_nchw(ho + hn)
    ht = prelu(ht, g("mru_conv_unit_last_norm/prelu/param"))
    feat = ht.mean(dim=(2, 3))
    wb = spectral_normed_weight(g("fully_connected/weights"), u[f"{name}/fully_connected/u"])
    return feat @ wb + g("fully_connected/biases")


# --------------------------------------------------------------------------------------
# losses (nn/functional.py, models/edgegan.py:32-42)
# --------------------------------------------------------------------------------------

def penalty(v, dname, synthesized, real, alpha, weight):
    """edgegan.py:32-42 + functional.py:26-29.  tf.gradients of the *tuple* (sigmoid(d), d)
    sums both outputs."""
    a = alpha.reshape(-1, 1, 1, 1)
    xhat = (real + a * (synthesized - real)).detach().requires_grad_(True)
    p, d = discriminator(v, dname, xhat)
    (gx,) = torch.autograd.grad((p.sum() + d.sum()), xhat, create_graph=True)
    gl2 = torch.sqrt((gx ** 2).sum(dim=(1, 2, 3)))
    return weight * ((gl2 - 1) ** 2).mean()


def focal_loss_real(logits, labels):
    """functional.py:8-11 (ld1 = 1, ld_focal = 2)."""
    p = torch.softmax(logits, dim=1)
    py = p.gather(1, labels.view(-1, 1)).squeeze(1)
    ce = TF.cross_entropy(logits, labels, reduction="none")
    return ((1 - py) ** 2 * ce).mean()


def ce_loss_fake(logits, labels):
    """functional.py:13-15 (ld2 = 0.5)."""
    return 0.5 * TF.cross_entropy(logits, labels, reduction="none").mean()


# --------------------------------------------------------------------------------------
# the step (models/edgegan.py:87-130, 202-342)
# --------------------------------------------------------------------------------------

RUN_NAMES = ("d_optim", "d_optim_patch2", "d_optim_patch3", "d_optim2", "g_optim_u", "e_optim", "g_optim_b")


@dataclass
class StepInputs:
    images: np.ndarray                 # [B, H, W_pair, 3]  (edge | image)
    z: np.ndarray                      # [B, z_dim (+1 class id)]
    alpha: np.ndarray                  # [3, B]  one vector per discriminator run (joint, image, edge)
    eps: float = 0.0                   # scalar encoder noise for run 6


class OracleState:
    """Weights + RMSProp slots (tf.train.RMSPropOptimizer: rms0 = 1, decay .9, eps 1e-10, A9)."""

    def __init__(self, cfg: Config, variables, sn_u=None, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.v = OrderedDict((k, torch.tensor(np.asarray(a), dtype=dtype)) for k, a in variables.items())
        self.u = OrderedDict((k, torch.tensor(np.asarray(a), dtype=dtype)) for k, a in (sn_u or {}).items())
        # one slot set per minimize() op; runs 5 and 7 share ops -> share slots.  Slots are keyed by
        # (optimizer-scope, var) where each network is only ever touched by one optimizer.
        self.rms = OrderedDict((k, torch.ones_like(t)) for k, t in self.v.items())
        self.losses = {}

    def numpy(self):
        return OrderedDict((k, t.detach().numpy().copy()) for k, t in self.v.items())


def _rmsprop(state: OracleState, names, grads):
    lr = state.cfg.learning_rate
    for n, g in zip(names, grads):
        if g is None:
            continue
        state.rms[n] = 0.9 * state.rms[n] + 0.1 * g * g
        state.v[n] = (state.v[n] - lr * g / torch.sqrt(state.rms[n] + 1e-10)).detach()


def _with_grad(state, scopes):
    names = [k for k in state.v if any(k.startswith(s + "/") for s in scopes)]
    for k in state.v:
        state.v[k] = state.v[k].detach()
    for k in names:
        state.v[k].requires_grad_(True)
    return names


def _g_input(cfg, z):
    """edgegan.py:188-197."""
    if not cfg.multiclasses:
        return z, None
    labels = z[:, -1].to(torch.int64)
    onehot = TF.one_hot(labels, cfg.num_classes).to(z.dtype)
    return torch.cat([z[:, : cfg.z_dim], onehot], dim=1), labels


def forward_fakes(state, zin):
    return generator(state.v, "G1", zin), generator(state.v, "G2", zin)


def update_model(state: OracleState, inp: StepInputs, runs=None, collect=None):
    """One EdgeGAN.update_model (edgegan.py:126-130): the 7 (6 single-class) sequential RMSProp
    runs in construct_optimizers order (edgegan.py:109-124).  `collect` (dict) receives
    per-run gradients / losses for golden vectors."""
    cfg = state.cfg
    dt = state.dtype
    images = torch.tensor(inp.images, dtype=dt)
    z = torch.tensor(inp.z, dtype=dt)
    alpha = torch.tensor(inp.alpha, dtype=dt)
    half = cfg.output_width // 2
    zin, labels = _g_input(cfg, z)
    edges_real, pics_real = images[:, :, :half, :], images[:, :, half:, :]
    todo = runs or [r for r in RUN_NAMES if (r != "d_optim2" or cfg.multiclasses)]

    def record(run, names, grads, loss):
        state.losses[run] = float(loss.detach())
        if collect is not None:
            collect[run] = {"loss": float(loss.detach()),
                            "grads": OrderedDict((n, g.detach().numpy().copy()) for n, g in zip(names, grads)
                                                 if g is not None)}

    def d_run(run, dname, real_fn, fake_fn, a):
        names = _with_grad(state, [dname])
        with torch.no_grad():
            e, i = forward_fakes(state, zin)
        real, fake = real_fn(), fake_fn(e, i)
        _, d_real = discriminator(state.v, dname, real)
        _, d_fake = discriminator(state.v, dname, fake)
        loss = (d_fake - d_real).mean() + penalty(state.v, dname, fake, real, a, cfg.lambda_gp)
        grads = torch.autograd.grad(loss, [state.v[n] for n in names], allow_unused=True)
        record(run, names, grads, loss)
        _rmsprop(state, names, grads)

    def g_run(run):
        names = _with_grad(state, ["G1", "G2"])
        e, i = forward_fakes(state, zin)
        _, dj = discriminator(state.v, "D", torch.cat([e, i], dim=2))
        joint_g = (-dj).mean()
        edge_gloss = cfg.joint_dweight * joint_g
        image_gloss = cfg.joint_dweight * joint_g
        if cfg.use_edge_discriminator:
            _, de = discriminator(state.v, "D_patch3", bicubic_up2(e, cfg.edge_dis_size))
            edge_gloss = edge_gloss + cfg.edge_dweight * (-de).mean()
        if cfg.use_image_discriminator:
            _, di = discriminator(state.v, "D_patch2", bicubic_up2(i, cfg.image_dis_size))
            image_gloss = image_gloss + cfg.image_dweight * (-di).mean()
        if cfg.multiclasses:
            image_gloss = image_gloss + ce_loss_fake(classifier(state.v, state.u, "D2", _nchw(i)), labels)
        n1 = [n for n in names if n.startswith("G1/")]
        n2 = [n for n in names if n.startswith("G2/")]
        g1 = torch.autograd.grad(edge_gloss, [state.v[n] for n in n1], retain_graph=True, allow_unused=True)
        g2 = torch.autograd.grad(image_gloss, [state.v[n] for n in n2], allow_unused=True)
        record(run, n1 + n2, list(g1) + list(g2), edge_gloss + image_gloss)
        state.losses[run + "/edge_gloss"] = float(edge_gloss.detach())
        state.losses[run + "/image_gloss"] = float(image_gloss.detach())
        _rmsprop(state, n1 + n2, list(g1) + list(g2))

    for run in todo:
        if run == "d_optim":
            d_run(run, "D", lambda: images, lambda e, i: torch.cat([e, i], dim=2), alpha[0])
        elif run == "d_optim_patch2":
            d_run(run, "D_patch2", lambda: bicubic_up2(pics_real, cfg.image_dis_size),
                  lambda e, i: bicubic_up2(i, cfg.image_dis_size), alpha[1])
        elif run == "d_optim_patch3":
            d_run(run, "D_patch3", lambda: bicubic_up2(edges_real, cfg.edge_dis_size),
                  lambda e, i: bicubic_up2(e, cfg.edge_dis_size), alpha[2])
        elif run == "d_optim2":
            names = _with_grad(state, ["D2"])
            loss = focal_loss_real(classifier(state.v, state.u, "D2", _nchw(pics_real)), labels)
            grads = torch.autograd.grad(loss, [state.v[n] for n in names], allow_unused=True)
            record(run, names, grads, loss)
            _rmsprop(state, names, grads)
        elif run in ("g_optim_u", "g_optim_b"):
            g_run(run)
        elif run == "e_optim":
            names = _with_grad(state, ["E"])
            with torch.no_grad():
                e = generator(state.v, "G1", zin)
            zr, _, _ = encoder(state.v, "E", e, inp.eps)
            loss = cfg.stage1_zl_loss * (z[:, : cfg.z_dim] - zr).abs().mean()
            grads = torch.autograd.grad(loss, [state.v[n] for n in names], allow_unused=True)
            record(run, names, grads, loss)
            _rmsprop(state, names, grads)
    for k in state.v:
        state.v[k] = state.v[k].detach()
    return state


def test_forward(state: OracleState, inputs, classes=None, eps=0.0):
    """edgegan.py:492-517 (inference graph): E(left half) -> z (+onehot) -> G1, G2."""
    cfg = state.cfg
    x = torch.tensor(inputs, dtype=state.dtype)
    left = x[:, :, : x.shape[2] // 2, :]
    with torch.no_grad():
        z, _, _ = encoder(state.v, "E", left, eps)
        if cfg.multiclasses:
            oh = TF.one_hot(torch.tensor(classes, dtype=torch.int64), cfg.num_classes).to(z.dtype)
            z = torch.cat([z, oh], dim=1)
        return generator(state.v, "G1", z).numpy(), generator(state.v, "G2", z).numpy()


def make_inputs(cfg: Config, seed: int = 2333, sketch_like: bool = False) -> StepInputs:
    """Seeded synthetic step inputs."""
    rs = np.random.RandomState(seed)
    B = cfg.batch_size
    images = rs.uniform(-1, 1, (B, cfg.output_height, cfg.output_width, cfg.c_dim)).astype(np.float32)
    if sketch_like:
        half = cfg.output_width // 2
        strokes = rs.uniform(size=(B, cfg.output_height, half, 1)) < 0.05
        images[:, :, :half, :] = np.where(strokes, -1.0, 1.0)
    z = rs.normal(size=(B, cfg.z_dim)).astype(np.float32)
    if cfg.multiclasses:
        cls = rs.randint(0, cfg.num_classes, size=(B, 1)).astype(np.float32)
        z = np.concatenate([z, cls], axis=1)
    alpha = rs.uniform(0, 1, (3, B)).astype(np.float32)
    eps = float(rs.normal())
    return StepInputs(images, z, alpha, eps)
