"""tcgen05 (kind::tf32) implicit-GEMM conv kernels vs the fp64 operator reference (-m gpu).

Tolerance: TF32 keeps 10 explicit mantissa bits per operand, fp32 accumulation -> the error of a length-K dot
product of O(1) terms is ~ 2^-10 * sqrt(K) * |x||w|; asserted as max-abs error <= 4e-3 * max|reference|."""
import numpy as np
import pytest
import torch

from ref_ops import RefOps

pytestmark = pytest.mark.gpu

TOL = 4e-3

CASES = [
    # N, H, W, Ci, Co, k, stride, pad, OH, OW
    (4, 32, 64, 64, 128, 4, 2, 1, 16, 32),     # d_conv_1
    (4, 16, 32, 128, 256, 4, 2, 1, 8, 16),     # d_conv_3
    (8, 8, 16, 256, 512, 4, 2, 1, 4, 8),       # d_conv_4
    (4, 8, 8, 256, 512, 5, 2, 1, 4, 4),        # g_dconv_1 seen as a conv (x = deconv output)
    (2, 32, 32, 64, 128, 5, 2, 1, 16, 16),     # g_dconv_3 seen as a conv
    (2, 34, 34, 64, 128, 3, 1, 0, 32, 32),     # encoder res1 (VALID on reflect-padded input)
    (4, 6, 6, 512, 512, 3, 1, 0, 4, 4),        # encoder last block
    (2, 32, 32, 64, 128, 1, 1, 0, 32, 32),     # 1x1 shortcut
    (3, 16, 16, 128, 128, 3, 1, 1, 16, 16),    # SAME 3x3 stride 1 (classifier style), batch not a tile multiple
    (8, 8, 8, 256, 512, 5, 2, 1, 4, 4),        # few output tiles, long K -> split-K path (generator at small batch)
    # thin layers (conv_thin.cu: patch matrix + dense product on the same kernels)
    (3, 16, 32, 3, 64, 4, 2, 1, 8, 16),        # critic first layer, pixel count not a tile multiple
    (2, 20, 12, 3, 64, 5, 2, 1, 10, 6),        # generator last conv-transpose seen as a conv (75 -> 96 / 128 columns)
    (2, 9, 9, 3, 64, 3, 1, 1, 9, 9),           # stride 1, odd extent
    (2, 8, 8, 1, 128, 4, 2, 1, 4, 4),          # single channel
    (2, 12, 12, 8, 128, 3, 1, 1, 12, 12),      # classifier first layer (8 channels, 72 -> 96 / 128 columns)
    (2, 12, 12, 8, 128, 1, 1, 0, 12, 12),      # the first MRU block's 1x1 `Conv_3` (8 -> 128): streaming kernels of conv_small.cu
    (3, 7, 5, 8, 128, 1, 1, 0, 7, 5),          # the same, pixel count not a multiple of anything
    # filter gradient whose row side is below 128 channels (TMA zero-fills the missing rows)
    (2, 16, 16, 64, 64, 3, 1, 1, 16, 16),
    (2, 8, 8, 192, 64, 1, 1, 0, 8, 8),         # 1.5 row tiles
]


@pytest.fixture(scope="module")
def dev():
    from edgegan_b200.ops import DeviceOps
    return DeviceOps()


@pytest.fixture(scope="module")
def ref():
    return RefOps(torch.float64)


def rnd(rs, *shape, scale=1.0):
    return (rs.standard_normal(shape) * scale).astype(np.float32)


def relerr(got, want):
    return float(np.abs(got.astype(np.float64) - want).max() / (np.abs(want).max() + 1e-20))


def run(o, name, ins, out_shape, *args, init=None):
    ti = [None if a is None else o.from_numpy(a) for a in ins]
    out = o.from_numpy(init) if init is not None else o.zeros(out_shape)
    getattr(o, name)(*ti, out, *args)
    return o.to_numpy(out)


# tc: operands rounded to nearest TF32; tc3x: 3xTF32 split (hi*hi + lo*hi + hi*lo), fp32-class accuracy
ALGO_TOL = {"tc": 2e-3, "tc3x": 2e-5}


@pytest.mark.parametrize("algo", ["tc", "tc3x"])
@pytest.mark.parametrize("case", CASES)
def test_tc_conv_trio(dev, ref, case, algo, request):
    import ctypes as C
    TOL = ALGO_TOL[algo]
    N, H, W, Ci, Co, k, s, p, OH, OW = case
    rs = np.random.RandomState(abs(hash(case)) % 2**31)
    x, w, b = rnd(rs, N, H, W, Ci), rnd(rs, k, k, Ci, Co, scale=0.05), rnd(rs, Co)
    dy, bi = rnd(rs, N, OH, OW, Co), rnd(rs, Ci)
    cs = dev._cs(x.shape, w.shape, dy.shape, s, p)
    request.addfinalizer(lambda: dev.lib.eg_debug_set(5, 2 | 4 | 32))
    # eg_debug_set(5, mask): bit 0 / 1 / 2 = forward / input gradient / filter gradient through the patch-matrix route of
    # conv_thin.cu, bit 3 = forward gather route OFF, bit 4 = filter-gradient gather route ON.  Thin layers run twice:
    # bit 5 = input gradient scattered by the dense product's epilogue instead of product matrix + col2im pass.
    # (a) forward gathered by the conditioning warps (the default) + gathered filter gradient + scatter epilogue,
    # (b) everything through the patch matrix, col2im pass for the input gradient.  Default (2 | 4 | 32): gathered forward,
    # input gradient scattered by the dense product's epilogue, filter gradient through the patch matrix; the FFMA filter
    # gradient of these layers is covered by test_ops_gpu.py.
    # bit 6 = the dedicated 1x1 8 <-> 128 streaming kernels OFF (so that those shapes also run the routes above)
    routes = [2 | 16 | 32 | 64, 7 | 8 | 64, 2 | 4 | 32] if Ci <= 8 else [2 | 4 | 32]
    for route in routes:
        dev.lib.eg_debug_set(5, route)
        used = [dev.lib.eg_conv2d_algo_for(C.byref(cs), i, 2) for i in range(3)]
        # these shapes are the ones the tensor-core path must cover
        assert used == [2, 2, 2], (route, used)
        _trio(dev, ref, rs, x, w, b, dy, bi, case, algo, TOL)


def _trio(dev, ref, rs, x, w, b, dy, bi, case, algo, TOL):
    N, H, W, Ci, Co, k, s, p, OH, OW = case
    want = run(ref, "conv_fwd", [x, w, b], (N, OH, OW, Co), s, p)
    got = run(dev, "conv_fwd", [x, w, b], (N, OH, OW, Co), s, p, algo)
    assert relerr(got, want) < TOL, ("fwd", relerr(got, want))
    want = run(ref, "conv_bwd_data", [dy, w, bi], (N, H, W, Ci), s, p)
    got = run(dev, "conv_bwd_data", [dy, w, bi], (N, H, W, Ci), s, p, algo)
    assert relerr(got, want) < TOL, ("bwd_data", relerr(got, want))
    want = run(ref, "conv_bwd_weight", [x, dy], (k, k, Ci, Co), s, p, False)
    got = run(dev, "conv_bwd_weight", [x, dy], (k, k, Ci, Co), s, p, False, algo)
    assert relerr(got, want) < TOL, ("bwd_weight", relerr(got, want))
    init = rnd(rs, k, k, Ci, Co)
    want = run(ref, "conv_bwd_weight", [x, dy], None, s, p, True, init=init)
    got = run(dev, "conv_bwd_weight", [x, dy], None, s, p, True, algo, init=init)
    assert relerr(got, want) < TOL, ("bwd_weight+acc", relerr(got, want))


def test_tc_matches_simt_at_full_size(dev):
    """d_conv_3 at the bench batch (3*64 samples): tensor-core result vs the fp32 SIMT kernel on the GPU."""
    rs = np.random.RandomState(0)
    N, H, W, Ci, Co = 192, 16, 32, 128, 256
    x = dev.from_numpy(rnd(rs, N, H, W, Ci))
    w = dev.from_numpy(rnd(rs, 4, 4, Ci, Co, scale=0.02))
    dy = dev.from_numpy(rnd(rs, N, 8, 16, Co))
    for name, args, shape in (("conv_fwd", (x, w, None), (N, 8, 16, Co)),
                              ("conv_bwd_data", (dy, w, None), (N, H, W, Ci))):
        a, b = dev.zeros(shape), dev.zeros(shape)
        getattr(dev, name)(*args, a, 2, 1, "simt")
        getattr(dev, name)(*args, b, 2, 1, "tc")
        assert relerr(dev.to_numpy(b), dev.to_numpy(a).astype(np.float64)) < TOL, name
    a, b = dev.zeros((4, 4, Ci, Co)), dev.zeros((4, 4, Ci, Co))
    dev.conv_bwd_weight(x, dy, a, 2, 1, False, "simt")
    dev.conv_bwd_weight(x, dy, b, 2, 1, False, "tc")
    assert relerr(dev.to_numpy(b), dev.to_numpy(a).astype(np.float64)) < TOL
    # 3xTF32 at full size (24 576 pixels per filter tap): fp32-class agreement with the FFMA kernel
    dev.conv_bwd_weight(x, dy, b, 2, 1, False, "tc3x")
    assert relerr(dev.to_numpy(b), dev.to_numpy(a).astype(np.float64)) < 5e-5
    for name, args, shape in (("conv_fwd", (x, w, None), (N, 8, 16, Co)), ("conv_bwd_data", (dy, w, None), (N, H, W, Ci))):
        a, b = dev.zeros(shape), dev.zeros(shape)
        getattr(dev, name)(*args, a, 2, 1, "simt")
        getattr(dev, name)(*args, b, 2, 1, "tc3x")
        assert relerr(dev.to_numpy(b), dev.to_numpy(a).astype(np.float64)) < 2e-5, name


EPI_CASES = [
    (3, 16, 32, 3, 64, 4, 2, 1, 8, 16),        # critic first layer: forward with act (FFMA kernel), thin input gradient
    (4, 32, 64, 64, 128, 4, 2, 1, 16, 32),     # critic second layer: input gradient with the first layer's mask
    (8, 8, 8, 256, 512, 5, 2, 1, 4, 4),        # split-K launch: act forces a single split, mask works through red.add
    (2, 9, 9, 3, 64, 3, 1, 1, 9, 9),
]


@pytest.mark.parametrize("algo", ["simt", "tc", "tc3x"])
@pytest.mark.parametrize("case", EPI_CASES)
def test_fused_epilogues(dev, ref, case, algo):
    """eg_conv2d_fwd_ex / eg_conv2d_bwd_data_ex: out = act(conv + bias) and out = (conv + bias) * act'(mask)."""
    tol = {"simt": 2e-5, "tc": 2e-3, "tc3x": 2e-5}[algo]
    N, H, W, Ci, Co, k, s, p, OH, OW = case
    rs = np.random.RandomState(abs(hash(case)) % 2**31)
    x, w, b = rnd(rs, N, H, W, Ci), rnd(rs, k, k, Ci, Co, scale=0.05), rnd(rs, Co)
    dy = rnd(rs, N, OH, OW, Co)
    my, mx = rnd(rs, N, OH, OW, Co), rnd(rs, N, H, W, Ci)
    for act in ("lrelu", "relu"):
        want = run(ref, "conv_fwd", [x, w, b], (N, OH, OW, Co), s, p, None, act)
        got = run(dev, "conv_fwd", [x, w, b], (N, OH, OW, Co), s, p, algo, act)
        # a value within rounding of zero may take the other branch of the activation: compare where |pre| is clear
        pre = run(ref, "conv_fwd", [x, w, b], (N, OH, OW, Co), s, p)
        clear = np.abs(pre) > 10 * tol * np.abs(pre).max()
        assert relerr(got * clear, want * clear) < tol, ("fwd+act", act, relerr(got * clear, want * clear))
        ti = [ref.from_numpy(a) for a in (x, w, b)]
        out = ref.zeros((N, OH, OW, Co))
        ref.conv_fwd(*ti, out, s, p, None, act, ref.from_numpy(my))
        want = ref.to_numpy(out)
        td = [dev.from_numpy(a) for a in (x, w, b)]
        out = dev.zeros((N, OH, OW, Co))
        dev.conv_fwd(*td, out, s, p, algo, act, dev.from_numpy(my))
        assert relerr(dev.to_numpy(out), want) < tol, ("fwd*mask", act)
        ti = [ref.from_numpy(a) for a in (dy, w)]
        out = ref.zeros((N, H, W, Ci))
        ref.conv_bwd_data(ti[0], ti[1], None, out, s, p, None, act, ref.from_numpy(mx))
        want = ref.to_numpy(out)
        td = [dev.from_numpy(a) for a in (dy, w)]
        out = dev.zeros((N, H, W, Ci))
        dev.conv_bwd_data(td[0], td[1], None, out, s, p, algo, act, dev.from_numpy(mx))
        assert relerr(dev.to_numpy(out), want) < tol, ("dgrad*mask", act)


def test_prepared_filter_set(dev, ref):
    """eg_filter_set_*: the filters of a set are prepared by ONE launch after each write; conv calls on them launch no
    preparation kernel and always reflect the weights of the last prepare(); a freed set is never served."""
    rs = np.random.RandomState(5)
    N, H, W, Ci, Co, k, s, p = 2, 16, 16, 64, 128, 3, 1, 1
    x, w1, w2 = rnd(rs, N, H, W, Ci), rnd(rs, k, k, Ci, Co, scale=0.05), rnd(rs, k, k, Ci, Co, scale=0.05)
    wb = rnd(rs, 1, 1, 128, 64, scale=0.05)                              # a second filter in the same set
    xd, wd, wbd, y = dev.from_numpy(x), dev.from_numpy(w1), dev.from_numpy(wb), dev.zeros((N, H, W, Co))
    want1 = run(ref, "conv_fwd", [x, w1, None], (N, H, W, Co), s, p)
    want2 = run(ref, "conv_fwd", [x, w2, None], (N, H, W, Co), s, p)
    dev.set_default_algo("tc3x")
    fs = dev.filter_set([wd, wbd])
    fs.prepare()
    h0, l0 = dev.filter_set_hits(), dev.launches
    for _ in range(3):
        dev.conv_fwd(xd, wd, None, y, s, p, "tc3x")
    assert dev.filter_set_hits() == h0 + 3 and dev.launches == l0 + 3        # conv kernels only
    assert relerr(dev.to_numpy(y), want1) < 2e-5
    dx = dev.zeros((N, H, W, Ci))
    dev.conv_bwd_data(y, wd, None, dx, s, p, "tc3x")                     # the other prepared layout, same set
    assert dev.filter_set_hits() == h0 + 4
    want_dx = run(ref, "conv_bwd_data", [dev.to_numpy(y), w1, None], (N, H, W, Ci), s, p)
    assert relerr(dev.to_numpy(dx), want_dx) < 2e-5
    y1 = dev.zeros((N, H, W, 64))
    dev.conv_fwd(y, wbd, None, y1, 1, 0, "tc3x")                         # second member
    assert relerr(dev.to_numpy(y1), run(ref, "conv_fwd", [dev.to_numpy(y), wb, None], (N, H, W, 64), 1, 0)) < 2e-5
    dev.conv_fwd(xd, wd, None, y, s, p, "tc")                            # another mode ignores the 3x copies
    assert relerr(dev.to_numpy(y), want1) < 4e-3
    dev.upload(wd, w2)                                                   # same pointer, new weights ...
    fs.prepare()                                                         # ... the writer refreshes the set
    dev.conv_fwd(xd, wd, None, y, s, p, "tc3x")
    assert relerr(dev.to_numpy(y), want2) < 2e-5
    g, ms = dev.from_numpy(np.ones_like(w2)), dev.from_numpy(np.ones_like(w2))
    dev.rmsprop(wd, g, ms, 0.1)                                          # w -= 0.1 * 1 / sqrt(0.9 + 0.1 + 1e-10)
    fs.prepare()
    dev.conv_fwd(xd, wd, None, y, s, p, "tc3x")
    want3 = run(ref, "conv_fwd", [x, w2 - np.float32(0.1), None], (N, H, W, Co), s, p)
    assert relerr(dev.to_numpy(y), want3) < 2e-5
    # dropping a member destroys the set: the pointer is prepared per call again (and stays correct)
    h1 = dev.filter_set_hits()
    del wbd
    import gc
    gc.collect()
    dev.upload(wd, w1)
    dev.conv_fwd(xd, wd, None, y, s, p, "tc3x")
    assert dev.filter_set_hits() == h1 and relerr(dev.to_numpy(y), want1) < 2e-5


@pytest.mark.parametrize("Ci,Co,k,st,H", [(3, 64, 4, 2, 16), (8, 128, 3, 1, 12), (3, 128, 3, 1, 9), (3, 64, 5, 2, 16)])
def test_prepared_filter_set_thin_filters(dev, ref, Ci, Co, k, st, H):
    """thin (image-side) filters in a set: the gathered forward and the patch-matrix input gradient read the set's copies
    (no per-call preparation launch) and follow the weights of the last prepare()"""
    rs = np.random.RandomState(6)
    N, p = 3, (k - 1) // 2
    OH = (H + 2 * p - k) // st + 1
    x, dy = rnd(rs, N, H, H, Ci), rnd(rs, N, OH, OH, Co)
    w1, w2 = rnd(rs, k, k, Ci, Co, scale=0.1), rnd(rs, k, k, Ci, Co, scale=0.1)
    wbig = rnd(rs, 3, 3, 64, 64, scale=0.05)                             # a standard member next to the thin one
    dev.set_default_algo("tc3x")
    xd, dyd, wd, wbd = dev.from_numpy(x), dev.from_numpy(dy), dev.from_numpy(w1), dev.from_numpy(wbig)
    y, dx = dev.zeros((N, OH, OH, Co)), dev.zeros((N, H, H, Ci))
    # per-call preparation first: launches of the unmanaged route
    l0 = dev.launches
    dev.conv_fwd(xd, wd, None, y, st, p, "tc3x")
    dev.conv_bwd_data(dyd, wd, None, dx, st, p, "tc3x")
    unmanaged = dev.launches - l0
    y_un, dx_un = dev.to_numpy(y), dev.to_numpy(dx)
    fs = dev.filter_set([wbd, wd])
    fs.prepare()
    h0, l0 = dev.filter_set_hits(), dev.launches
    dev.conv_fwd(xd, wd, None, y, st, p, "tc3x")
    dev.conv_bwd_data(dyd, wd, None, dx, st, p, "tc3x")
    assert dev.filter_set_hits() == h0 + 2, "thin filter not served from the set"
    assert dev.launches - l0 == unmanaged - 3            # gather copy; transpose + operand preparation of the input gradient
    np.testing.assert_array_equal(dev.to_numpy(y), y_un)
    assert relerr(dev.to_numpy(dx), dx_un) < 1e-5        # scattered by red.add: same terms, order-dependent last bits
    assert relerr(y_un, run(ref, "conv_fwd", [x, w1, None], (N, OH, OH, Co), st, p)) < 2e-5
    assert relerr(dx_un, run(ref, "conv_bwd_data", [dy, w1, None], (N, H, H, Ci), st, p)) < 2e-5
    dev.upload(wd, w2)
    fs.prepare()
    dev.conv_fwd(xd, wd, None, y, st, p, "tc3x")
    dev.conv_bwd_data(dyd, wd, None, dx, st, p, "tc3x")
    assert relerr(dev.to_numpy(y), run(ref, "conv_fwd", [x, w2, None], (N, OH, OH, Co), st, p)) < 2e-5
    assert relerr(dev.to_numpy(dx), run(ref, "conv_bwd_data", [dy, w2, None], (N, H, H, Ci), st, p)) < 2e-5
    y2 = dev.zeros((N, H, H, 64))
    x2 = rnd(rs, N, H, H, 64)
    dev.conv_fwd(dev.from_numpy(x2), wbd, None, y2, 1, 1, "tc3x")        # the standard member is untouched by the thin one
    assert relerr(dev.to_numpy(y2), run(ref, "conv_fwd", [x2, wbig, None], (N, H, H, 64), 1, 1)) < 2e-5
    fs.close()
    h1 = dev.filter_set_hits()
    dev.conv_fwd(xd, wd, None, y, st, p, "tc3x")
    dev.conv_bwd_data(dyd, wd, None, dx, st, p, "tc3x")
    assert dev.filter_set_hits() == h1
    assert relerr(dev.to_numpy(y), run(ref, "conv_fwd", [x, w2, None], (N, OH, OH, Co), st, p)) < 2e-5
