"""Per-kernel parity (-m gpu): every C-ABI entry point, called through edgegan_b200.ops.DeviceOps, against
the fp64 CPU operator reference in tests/ref_ops.py on the same seeded inputs.

Tolerances: fp32 kernels must agree to ~1e-5 relative to the largest reference magnitude (fp32
accumulation-order noise); the tcgen05 TF32 path is tested separately in test_conv_tc_gpu.py."""
import numpy as np
import pytest
import torch

from ref_ops import RefOps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from edgegan_b200.ops import DeviceOps
    return DeviceOps()


@pytest.fixture(scope="module")
def ref():
    return RefOps(torch.float64)


def rnd(rs, *shape, scale=1.0):
    return (rs.standard_normal(shape) * scale).astype(np.float32)


def close(got, want, tol=2e-5, what=""):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.isfinite(got).all(), what
    err = np.abs(got - want).max() / (np.abs(want).max() + 1e-20)
    assert err < tol, (what, err)


def both(dev, ref, name, ins, outs, *args, tol=2e-5, **kw):
    """ins: list of numpy arrays or None; outs: list of shapes (or numpy arrays = initial contents)."""
    res = []
    for o in (dev, ref):
        ti = [None if a is None else o.from_numpy(a) for a in ins]
        to = [o.from_numpy(s) if isinstance(s, np.ndarray) else o.zeros(s) for s in outs]
        getattr(o, name)(*ti[:kw.get("n_in", len(ti))], *to, *args)
        res.append([o.to_numpy(t) for t in to])
    for k, (g, w) in enumerate(zip(*res)):
        close(g, w, tol, f"{name}[{k}]")
    return res[0]


CONV_CASES = [
    # N, H, W, Ci, Co, k, stride, pad, OH, OW
    (2, 16, 32, 3, 64, 4, 2, 1, 8, 16),        # d_conv_0 (thin Cin, scalar path)
    (3, 16, 16, 64, 128, 4, 2, 1, 8, 8),       # d_conv_1 style
    (2, 8, 8, 128, 64, 5, 2, 1, 4, 4),         # deconv seen as a conv: x = deconv output [8x8], y = deconv input [4x4]
    (2, 16, 16, 3, 64, 5, 2, 1, 8, 8),         # g_dconv_4 (Cout_g = 3)
    (2, 10, 10, 64, 128, 3, 1, 0, 8, 8),       # encoder 3x3 VALID on the reflect-padded map
    (2, 8, 8, 64, 128, 1, 1, 0, 8, 8),         # 1x1 shortcut
    (5, 1, 1, 100, 512, 1, 1, 0, 1, 1),        # linear as a 1x1 conv
    (2, 9, 7, 8, 20, 3, 1, 1, 9, 7),           # SAME 3x3 stride 1, odd sizes, ragged channels
    (1, 12, 12, 16, 8, 7, 1, 3, 12, 12),       # 7x7 SAME (classifier h0 style)
    (4, 6, 6, 512, 512, 3, 1, 0, 4, 4),        # encoder last block (few pixels, long K)
    (4, 4, 4, 512, 512, 1, 1, 0, 4, 4),        # its 1x1 shortcut
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_trio_simt(dev, ref, case):
    N, H, W, Ci, Co, k, s, p, OH, OW = case
    rs = np.random.RandomState(hash(case) % 2**31)
    x, w, b = rnd(rs, N, H, W, Ci), rnd(rs, k, k, Ci, Co, scale=0.05), rnd(rs, Co)
    dy = rnd(rs, N, OH, OW, Co)
    bi = rnd(rs, Ci)
    for o_name, ins, out, extra in (
            ("conv_fwd", [x, w, b], (N, OH, OW, Co), (s, p, "simt")),
            ("conv_fwd", [x, w, None], (N, OH, OW, Co), (s, p, "simt")),
            ("conv_bwd_data", [dy, w, None], (N, H, W, Ci), (s, p, "simt")),
            ("conv_bwd_data", [dy, w, bi], (N, H, W, Ci), (s, p, "simt"))):
        both(dev, ref, o_name, ins, [out], *extra)
    both(dev, ref, "conv_bwd_weight", [x, dy], [(k, k, Ci, Co)], s, p, False, "simt")
    both(dev, ref, "conv_bwd_weight", [x, dy], [rnd(rs, k, k, Ci, Co)], s, p, True, "simt")


SMALL_CASES = [
    # N, H, W, Ci, Co, k, pad  (stride 1): the classifier's stem, both channel counts <= 8 (conv_small.cu)
    (2, 64, 64, 3, 8, 7, 3),        # `Conv` 7x7 (classifier.py:27)
    (3, 40, 24, 3, 8, 7, 3),        # ragged tiles
    (2, 64, 64, 3, 8, 3, 1),        # unit-1 image conv / gate part
    (3, 33, 70, 3, 8, 3, 1),
    (2, 64, 64, 8, 8, 3, 1),        # unit-1 update gate on the hidden state
    (5, 17, 31, 8, 8, 3, 1),
    (2, 12, 12, 8, 8, 3, 0),        # VALID
]


@pytest.mark.parametrize("case", SMALL_CASES)
def test_conv_trio_small_channels(dev, ref, case, request):
    """direct kernels of the <= 8-channel layers against the CPU operator set and against the generic implicit GEMM
    (eg_debug_set(7, 2) switches them off)"""
    N, H, W, Ci, Co, k, p = case
    OH, OW = H + 2 * p - k + 1, W + 2 * p - k + 1
    rs = np.random.RandomState(hash(case) % 2**31)
    x, w, b = rnd(rs, N, H, W, Ci), rnd(rs, k, k, Ci, Co, scale=0.1), rnd(rs, Co)
    dy, bi, dw0 = rnd(rs, N, OH, OW, Co), rnd(rs, Ci), rnd(rs, k, k, Ci, Co)
    request.addfinalizer(lambda: dev.lib.eg_debug_set(7, 0))
    outs = {}
    for route in (0, 2):
        dev.lib.eg_debug_set(7, route)
        outs[route] = [
            both(dev, ref, "conv_fwd", [x, w, b], [(N, OH, OW, Co)], 1, p, "simt")[0],
            both(dev, ref, "conv_fwd", [x, w, None], [(N, OH, OW, Co)], 1, p, "simt")[0],
            both(dev, ref, "conv_bwd_data", [dy, w, None], [(N, H, W, Ci)], 1, p, "simt")[0],
            both(dev, ref, "conv_bwd_data", [dy, w, bi], [(N, H, W, Ci)], 1, p, "simt")[0],
            both(dev, ref, "conv_bwd_weight", [x, dy], [(k, k, Ci, Co)], 1, p, False, "simt", tol=5e-5)[0],
            both(dev, ref, "conv_bwd_weight", [x, dy], [dw0.copy()], 1, p, True, "simt", tol=5e-5)[0]]
    for a, g in zip(outs[0], outs[2]):
        close(a, g, 5e-5, "direct vs generic")


def test_conv_large_splitk(dev, ref):
    # big pixel count -> split-K wgrad with atomics
    rs = np.random.RandomState(5)
    N, H, W, Ci, Co = 8, 32, 32, 64, 64
    x, dy = rnd(rs, N, H, W, Ci), rnd(rs, N, 16, 16, Co)
    both(dev, ref, "conv_bwd_weight", [x, dy], [(4, 4, Ci, Co)], 2, 1, False, "simt", tol=5e-5)


# (2, 32, 32, 128): 1024-pixel slabs -> one block per slab with a narrower channel group; (1, 64, 64, 64), (2, 64, 64, 32),
# (1, 96, 96, 8), (1, 128, 128, 8): 4096+ pixels -> slabs split over thread-block clusters of 4-8 blocks (with clusters off:
# one big block per SM or the multi-pass kernels); (2, 4, 4, 40), (2, 5, 3, 6): odd channel counts -> multi-pass kernels
@pytest.mark.parametrize("shape", [(3, 8, 8, 64), (2, 16, 32, 128), (2, 4, 4, 40), (1, 2, 2, 512), (2, 32, 32, 128),
                                   (1, 96, 96, 8), (2, 5, 3, 6), (1, 64, 64, 64), (2, 64, 64, 32), (1, 128, 128, 8),
                                   (3, 24, 24, 96)])
@pytest.mark.parametrize("act", ["none", "relu", "lrelu"])
def test_instnorm(dev, ref, shape, act, request):
    """both block layouts: slabs split over thread-block clusters (default) and one block per slab (eg_norm_debug(-2))"""
    rs = np.random.RandomState(1)
    N, C = shape[0], shape[-1]
    x, gy, t, add = rnd(rs, *shape), rnd(rs, *shape), rnd(rs, *shape), rnd(rs, *shape)
    request.addfinalizer(lambda: dev.lib.eg_norm_debug(-3))
    for knob in (-3, -2):
        dev.lib.eg_norm_debug(knob)
        y, st = both(dev, ref, "instnorm_fwd", [x], [shape, (N, C, 2)], act)
        both(dev, ref, "instnorm_bwd", [x, st, gy, None], [shape], act, tol=1e-4)
        both(dev, ref, "instnorm_bwd", [x, st, gy, add], [shape], act, tol=1e-4)
        both(dev, ref, "instnorm_bwd2", [x, st, gy, t], [shape, shape], act, tol=2e-4)


@pytest.mark.parametrize("act", ["none", "relu", "lrelu", "tanh", "sigmoid"])
def test_activations(dev, ref, act):
    rs = np.random.RandomState(2)
    x, gy = rnd(rs, 1000), rnd(rs, 1000)
    both(dev, ref, "act_fwd", [x], [(1000,)], act)
    both(dev, ref, "act_bwd", [x, gy], [(1000,)], act)


def test_batchnorm(dev, ref):
    rs = np.random.RandomState(3)
    R, C = 96, 80
    x, gy = rnd(rs, R, C) + 0.3, rnd(rs, R, C)
    gamma, beta = rnd(rs, C) + 1.0, rnd(rs, C) * 0.1
    (sums,) = both(dev, ref, "bn_stats", [x], [(2 * C,)])
    count = float(R)
    # apply / bwd take (x, sums, count, gamma, beta, ...) with `count` a python scalar in the middle
    res = []
    for o in (dev, ref):
        X, S, G, Bt, GY = (o.from_numpy(a) for a in (x, sums, gamma, beta, gy))
        y, red, gx = o.zeros((R, C)), o.zeros((2 * C,)), o.zeros((R, C))
        o.bn_apply(X, S, count, G, Bt, y, "relu")
        o.bn_bwd_reduce(X, S, count, G, Bt, GY, red, "relu")
        o.bn_bwd_apply(X, S, count, G, Bt, GY, red, gx, "relu")
        res.append([o.to_numpy(t) for t in (y, red, gx)])
    for g, w, nm in zip(res[0], res[1], ("y", "red", "gx")):
        close(g, w, 5e-5, "bn_" + nm)


def test_rowdot(dev, ref):
    rs = np.random.RandomState(4)
    B, F = 6, 4096
    h, w, b, gd = rnd(rs, B, F), rnd(rs, F, 1, scale=0.02), rnd(rs, 1), rnd(rs, B)
    both(dev, ref, "rowdot_fwd", [h, w, b], [(B,)])
    both(dev, ref, "rowdot_fwd", [h, w, None], [(B,)])
    both(dev, ref, "rowdot_bwd_input", [gd, w], [(B, F)])
    both(dev, ref, "rowdot_bwd_weight", [gd, h], [(F, 1), (1,)], False)
    both(dev, ref, "rowdot_bwd_weight", [gd, h], [rnd(rs, F, 1), rnd(rs, 1)], True)


@pytest.mark.parametrize("shape", [(2, 8, 8, 3), (1, 5, 7, 4), (2, 64, 64, 3), (3, 6, 5, 1), (2, 2, 2, 3)])
def test_bicubic(dev, ref, shape):
    rs = np.random.RandomState(6)
    N, H, W, C = shape
    x, gy = rnd(rs, *shape), rnd(rs, N, 2 * H, 2 * W, C)
    (y,) = both(dev, ref, "bicubic_up2_fwd", [x], [(N, 2 * H, 2 * W, C)])
    assert np.array_equal(y[:, ::2, ::2, :], x)          # even output pixels are copies
    both(dev, ref, "bicubic_up2_bwd", [gy], [shape])


def test_slices_fill_axpby(dev, ref):
    rs = np.random.RandomState(7)
    src, dst = rnd(rs, 2, 4, 6, 3), rnd(rs, 2, 4, 10, 3)
    res = []
    for o in (dev, ref):
        S, D = o.from_numpy(src), o.from_numpy(dst)
        o.copy_wslice(S, 1, D, 3, 4)
        Y = o.from_numpy(src)
        o.axpby(o.from_numpy(src * 2), Y, 0.5, -1.5)
        Fz = o.zeros((7,))
        o.fill(Fz, 3.25)
        res.append([o.to_numpy(t) for t in (D, Y, Fz)])
    for g, w in zip(*res):
        close(g, w, 1e-6)


def test_gp_kernels(dev, ref):
    rs = np.random.RandomState(8)
    B, shape = 5, (5, 8, 16, 3)
    real, fake, alpha = rnd(rs, *shape), rnd(rs, *shape), rs.uniform(size=B).astype(np.float32)
    both(dev, ref, "gp_interpolate", [real, fake, alpha], [shape])
    d = rnd(rs, B)
    both(dev, ref, "gp_seed", [d], [(B,)])
    both(dev, ref, "gp_seed_bwd", [d, rnd(rs, B)], [(B,)])
    g = rnd(rs, *shape, scale=0.05)
    both(dev, ref, "gp_penalty", [g], [shape, (B,), np.array([0.5], np.float32)], 10.0, 1.0 / 7)
    res = []
    for o in (dev, ref):
        out = o.from_numpy(np.array([2.0], np.float32))
        o.sum_scaled(o.from_numpy(d), 0.25, out, True)
        out2 = o.from_numpy(np.array([2.0], np.float32))
        o.sum_scaled(o.from_numpy(d), 0.25, out2, False)
        res.append([o.to_numpy(out), o.to_numpy(out2)])
    for g_, w_ in zip(*res):
        close(g_, w_, 1e-5)


def test_encoder_pieces(dev, ref):
    rs = np.random.RandomState(9)
    shape = (2, 6, 8, 5)
    x = rnd(rs, *shape)
    both(dev, ref, "reflect_pad_fwd", [x], [(2, 8, 10, 5)], 1)
    both(dev, ref, "reflect_pad_bwd", [rnd(rs, 2, 8, 10, 5)], [shape], 1)
    both(dev, ref, "reflect_pad_bwd", [rnd(rs, 1, 4, 4, 3)], [(1, 2, 2, 3)], 1)     # smallest map (4x4 block)
    for sh in ((3, 6, 8, 64), (2, 2, 2, 8), (2, 5, 3, 12)):                          # 16-byte kernels (C % 4 == 0)
        N, H, W, Cn = sh
        both(dev, ref, "reflect_pad_fwd", [rnd(rs, *sh)], [(N, H + 2, W + 2, Cn)], 1)
        both(dev, ref, "reflect_pad_bwd", [rnd(rs, N, H + 2, W + 2, Cn)], [sh], 1)
    a, b = rnd(rs, *shape), rnd(rs, *shape)
    both(dev, ref, "addrelu_pool2_fwd", [a, b], [(2, 3, 4, 5)])
    both(dev, ref, "addrelu_pool2_bwd", [a, b, rnd(rs, 2, 3, 4, 5)], [shape])
    both(dev, ref, "relu_globalmean_fwd", [x], [(2, 5)])
    both(dev, ref, "relu_globalmean_bwd", [x, rnd(rs, 2, 5)], [shape])
    mu, ls = rnd(rs, 4, 10), rnd(rs, 4, 10, scale=0.3)
    res = []
    tgt = rnd(rs, 4, 11)
    for o in (dev, ref):
        M, L, T = o.from_numpy(mu), o.from_numpy(ls), o.from_numpy(tgt)
        z, gmu, gls, loss = o.zeros((4, 10)), o.zeros((4, 10)), o.zeros((4, 10)), o.zeros((1,))
        o.reparam_fwd(M, L, 0.7, z)
        o.zl1_loss_bwd(M, L, 0.7, T, 10.0, 1.0 / 40, gmu, gls, loss)
        res.append([o.to_numpy(t) for t in (z, gmu, gls, loss)])
    for g_, w_ in zip(*res):
        close(g_, w_, 1e-5)


def test_onehot_rmsprop_biasgrad(dev, ref):
    rs = np.random.RandomState(10)
    z = rnd(rs, 6, 11)
    z[:, 10] = rs.randint(0, 14, 6)
    res = []
    for o in (dev, ref):
        out = o.zeros((6, 24))
        o.onehot_concat(o.from_numpy(z), 10, 14, out)
        var, grad, ms = o.from_numpy(rnd(np.random.RandomState(1), 1000)), o.from_numpy(rnd(np.random.RandomState(2), 1000, scale=0.01)), o.zeros((1000,))
        o.fill(ms, 1.0)
        o.rmsprop(var, grad, ms, 2e-4)
        o.rmsprop(var, grad, ms, 2e-4)
        dy = o.from_numpy(rnd(np.random.RandomState(3), 3000, 40))
        db, db2 = o.zeros((40,)), o.from_numpy(np.ones(40, np.float32))
        o.bias_grad(dy, db, False)
        o.bias_grad(dy, db2, True)
        res.append([o.to_numpy(t) for t in (out, var, ms, db, db2)])
    for g_, w_ in zip(*res):
        close(g_, w_, 2e-5)


# ---- classifier kernels -------------------------------------------------------------------------------------
def test_prelu_lrelu2(dev, ref):
    rs = np.random.RandomState(20)
    x, gy = rnd(rs, 3, 8, 8, 16), rnd(rs, 3, 8, 8, 16)
    x.reshape(-1)[:5] = 0.0                                    # ties at zero
    leak = np.array(0.2, np.float32).reshape(1)
    res = []
    for o in (dev, ref):
        X, L, GY = o.from_numpy(x), o.from_numpy(leak), o.from_numpy(gy)
        y, gx, gl, gl2 = o.zeros(x.shape), o.zeros(x.shape), o.zeros((1,)), o.from_numpy(np.array([1.5], np.float32))
        o.prelu_fwd(X, L, y)
        o.prelu_bwd(X, L, GY, gx, gl, False)
        o.prelu_bwd(X, L, GY, None, gl2, True)
        a, ga = o.zeros(x.shape), o.zeros(x.shape)
        o.act_fwd(X, a, "lrelu2")
        o.act_bwd(X, GY, ga, "lrelu2")
        res.append([o.to_numpy(t) for t in (y, gx, gl, gl2, a, ga)])
    for g_, w_ in zip(*res):
        close(g_, w_, 2e-5)


@pytest.mark.parametrize("shape", [(2, 8, 8, 40), (3, 16, 16, 8)])
def test_minmax(dev, ref, shape):
    rs = np.random.RandomState(21)
    x, gy = rnd(rs, *shape), rnd(rs, *shape)
    x[0, 0, 0, :] = x[0, 1, 1, :] = x.reshape(shape[0], -1, shape[-1]).max(1)[0] + 0.5     # tied maxima in sample 0
    N, C = shape[0], shape[-1]
    y, st = both(dev, ref, "minmax_fwd", [x], [shape, (N, C, 2)])
    both(dev, ref, "minmax_bwd", [x, st, gy], [shape], tol=5e-5)


@pytest.mark.parametrize("shape", [(2, 8, 8, 40), (3, 16, 16, 8), (2, 32, 32, 128), (2, 4, 4, 12)])
def test_mru_gate_fused(dev, ref, shape):
    """eg_mru_gate_fwd / _bwd (the fused element-wise middle of an MRU unit) against the composition of the separate
    reference ops, with tied extrema and zeros in the gate pre-activation."""
    rs = np.random.RandomState(23)
    N, C = shape[0], shape[-1]
    cg, cgi, ht, img, g_hin, g_ht0 = (rnd(rs, *shape) for _ in range(6))
    top = (cg + cgi).reshape(N, -1, C)[0].max(0) + 0.5
    cgi[0, 0, 0, :] = cgi[0, 1, 1, :] = 0.0                                # two tied maxima in sample 0 (exact in fp32
    cg[0, 0, 0, :] = cg[0, 1, 1, :] = top                                  # and in fp64: the other summand is zero) ...
    cg[1, 2, 2, :] = cgi[1, 2, 2, :] = 0.0                                 # ... and exact zeros (lrelu tie) in sample 1
    leak = np.array([0.2], np.float32)
    res = []
    for o in (dev, ref):
        CG, CGI, HT, IMG, L = (o.from_numpy(a) for a in (cg, cgi, ht, img, leak))
        st, plus, hin = o.zeros((N, C, 4)), o.zeros(shape), o.zeros(shape)
        o.mru_gate_fwd(CG, CGI, HT, IMG, L, st, plus, hin)
        g_ht, g_img, g_cg, gl = o.from_numpy(g_ht0), o.zeros(shape), o.zeros(shape), o.zeros((1,))
        o.mru_gate_bwd(plus, o.from_numpy(g_hin), IMG, CG, st, L, g_ht, g_img, g_cg, gl, False)
        gl2 = o.from_numpy(np.array([2.0], np.float32))
        g_ht2, g_img2, g_cg2 = o.from_numpy(g_ht0), o.zeros(shape), o.zeros(shape)
        o.mru_gate_bwd(plus, o.from_numpy(g_hin), IMG, CG, st, L, g_ht2, g_img2, g_cg2, gl2, True)
        res.append([o.to_numpy(t) for t in (CG, st, plus, hin, g_ht, g_img, g_cg, gl, gl2, g_cg2)])
    for k, (g_, w_) in enumerate(zip(*res)):
        close(g_, w_, 5e-5, f"mru_gate[{k}]")
    assert (res[0][1][0, :, 3] == 2).all()                                  # both tied maxima counted


def test_prelu_fused_variants(dev, ref):
    rs = np.random.RandomState(24)
    for shape in ((3, 8, 8, 16), (1, 3, 5, 7)):                             # vector body and scalar tail
        x, gy, g0 = rnd(rs, *shape), rnd(rs, *shape), rnd(rs, *shape)
        l1, l2 = np.array([0.2], np.float32), np.array([-0.3], np.float32)
        res = []
        for o in (dev, ref):
            X, GY = o.from_numpy(x), o.from_numpy(gy)
            y, y2, gx, gl = o.zeros(shape), o.zeros(shape), o.from_numpy(g0), o.zeros((1,))
            o.prelu_fwd2(X, o.from_numpy(l1), y, o.from_numpy(l2), y2)
            o.prelu_bwd(X, o.from_numpy(l1), GY, gx, gl, False, accumulate_gx=True)
            res.append([o.to_numpy(t) for t in (y, y2, gx, gl)])
        for g_, w_ in zip(*res):
            close(g_, w_, 2e-5)
    a, b = rnd(rs, 2, 6, 8, 8), rnd(rs, 2, 6, 8, 8)
    res = []
    for o in (dev, ref):
        y, ya = o.zeros((2, 3, 4, 8)), o.zeros((2, 3, 4, 8))
        o.add_pool2_fwd(o.from_numpy(a), o.from_numpy(b), y, o.from_numpy(np.array([0.25], np.float32)), ya)
        res.append([o.to_numpy(y), o.to_numpy(ya)])
    for g_, w_ in zip(*res):
        close(g_, w_, 2e-5)


def test_fma_mul_pool_mean_cslice(dev, ref):
    rs = np.random.RandomState(22)
    shape = (2, 6, 8, 5)
    a, b, c = rnd(rs, *shape), rnd(rs, *shape), rnd(rs, *shape)
    both(dev, ref, "fma3", [a, b, c], [shape])
    both(dev, ref, "mul", [a, b], [shape])
    both(dev, ref, "add_pool2_fwd", [a, b], [(2, 3, 4, 5)])
    both(dev, ref, "add_pool2_fwd", [a, None], [(2, 3, 4, 5)])
    gy = rnd(rs, 2, 3, 4, 5)
    both(dev, ref, "pool2_bwd", [gy], [shape], False)
    both(dev, ref, "pool2_bwd", [gy], [rnd(rs, *shape)], True)
    a8, b8, gy8 = rnd(rs, 2, 6, 8, 8), rnd(rs, 2, 6, 8, 8), rnd(rs, 2, 3, 4, 8)      # float4 path
    both(dev, ref, "fma3", [a8, b8, a8], [a8.shape])
    both(dev, ref, "mul", [a8, b8], [a8.shape])
    both(dev, ref, "add_pool2_fwd", [a8, b8], [(2, 3, 4, 8)])
    both(dev, ref, "pool2_bwd", [gy8], [a8.shape], False)
    both(dev, ref, "pool2_bwd", [gy8], [rnd(rs, *a8.shape)], True)
    both(dev, ref, "globalmean_fwd", [a], [(2, 5)])
    both(dev, ref, "globalmean_bwd", [rnd(rs, 2, 5)], [shape])
    res = []
    for o in (dev, ref):
        S, D = o.from_numpy(a), o.from_numpy(rnd(np.random.RandomState(1), 2, 6, 8, 9))
        o.copy_cslice(S, 1, D, 4, 3)
        res.append(o.to_numpy(D))
    close(res[0], res[1], 1e-6)


@pytest.mark.parametrize("shape", [(3, 3, 11, 8), (3, 3, 128, 256), (1, 1, 768, 14), (7, 7, 3, 8)])
def test_spectral_norm(dev, ref, shape):
    rs = np.random.RandomState(23)
    W, u, G = rnd(rs, *shape, scale=0.02), rnd(rs, 1, shape[-1]), rnd(rs, *shape)
    Cn = shape[-1]
    K = int(np.prod(shape)) // Cn
    res = []
    for o in (dev, ref):
        Wt, ut, Gt = o.from_numpy(W), o.from_numpy(u), o.from_numpy(G)
        wb, ws, gw = o.zeros(shape), o.zeros((o.sn_ws_floats(K, Cn),)), o.zeros(shape)
        o.spectral_norm_fwd(Wt, ut, wb, ws)
        o.spectral_norm_bwd(Wt, ut, ws, Gt, gw)
        res.append([o.to_numpy(wb), o.to_numpy(gw)])
    close(res[0][0], res[1][0], 2e-5, "wbar")
    close(res[0][1], res[1][1], 1e-4, "gW")


def test_spectral_norm_set(dev, ref):
    """the whole-network table (eg_spectral_norm_set_*): tensors of very different sizes in one launch set, one of them
    split along its input-channel axis (the classifier's update gate), forward twice / backward twice per set"""
    rs = np.random.RandomState(29)
    shapes = [(7, 7, 3, 8), (3, 3, 11, 8), (3, 3, 128, 256), (1, 1, 768, 14), (3, 3, 131, 128), (3, 3, 768, 768)]
    split_at = {4: 128}
    host = [(rnd(rs, *sh, scale=0.02), rnd(rs, 1, sh[-1]), rnd(rs, *sh)) for sh in shapes]
    res = []
    for o in (dev, ref):
        items, outs = [], []
        for i, (sh, (W, u, G)) in enumerate(zip(shapes, host)):
            Cn = sh[-1]
            K = int(np.prod(sh)) // Cn
            it = dict(W=o.from_numpy(W), u=o.from_numpy(u), Wbar=o.zeros(sh), ws=o.zeros((o.sn_ws_floats(K, Cn),)), gW=o.zeros(sh))
            if i in split_at:
                hd = split_at[i]
                it.update(Wa=o.zeros(sh[:2] + (hd, Cn)), Wi=o.zeros(sh[:2] + (sh[2] - hd, Cn)), hd=hd,
                          Ga=o.from_numpy(np.ascontiguousarray(G[:, :, :hd])), Gi=o.from_numpy(np.ascontiguousarray(G[:, :, hd:])))
            else:
                it["G"] = o.from_numpy(G)
            items.append(it)
        st = o.spectral_norm_set(items)
        for _ in range(2):
            st.fwd()
            st.bwd()
        st.bwd()
        res.append([(o.to_numpy(it["Wbar"]), o.to_numpy(it["gW"]),
                     o.to_numpy(it["Wa"]) if "Wa" in it else None, o.to_numpy(it["Wi"]) if "Wi" in it else None) for it in items])
        if o is dev:
            # against the single-tensor entry points of the same library
            for it, (W, u, G) in zip(items, host):
                wb, gw = o.zeros(it["W"].shape), o.zeros(it["W"].shape)
                ws = o.zeros(it["ws"].shape)
                o.spectral_norm_fwd(it["W"], it["u"], wb, ws)
                o.spectral_norm_bwd(it["W"], it["u"], ws, o.from_numpy(G), gw)
                close(o.to_numpy(it["Wbar"]), o.to_numpy(wb), 1e-6, "set vs single wbar")     # r is summed by atomics
                close(o.to_numpy(it["gW"]), o.to_numpy(gw), 1e-5, "set vs single gW")
            st.close()
    for i, (a, b) in enumerate(zip(*res)):
        close(a[0], b[0], 2e-5, f"wbar {i}")
        close(a[1], b[1], 1e-4, f"gW {i}")
        if a[2] is not None:
            hd = split_at[i]
            np.testing.assert_array_equal(a[2], a[0][:, :, :hd])
            np.testing.assert_array_equal(a[3], a[0][:, :, hd:])


@pytest.mark.parametrize("focal", [False, True])
def test_softmax_ce(dev, ref, focal):
    rs = np.random.RandomState(24)
    B, Cn = 9, 14
    logits = rnd(rs, B, Cn, scale=2.0)
    z = rnd(rs, B, 101)
    z[:, 100] = rs.randint(0, Cn, B)
    res = []
    for o in (dev, ref):
        gl, loss = o.zeros((B, Cn)), o.from_numpy(np.array([0.25], np.float32))
        o.softmax_ce_bwd(o.from_numpy(logits), o.from_numpy(z), 100, focal, 0.5, 1.0 / 11, gl, loss)
        res.append([o.to_numpy(gl), o.to_numpy(loss)])
    for g_, w_ in zip(*res):
        close(g_, w_, 2e-5)
