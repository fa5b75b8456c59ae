"""edgegan_b200.nn -- the eager mirror of the reference op library -- assembled into the reference's networks
exactly as edgegan/models/{discriminator,generator,encoder}.py assemble edgegan.nn, on the CPU operator reference,
against the oracle."""
import numpy as np
import torch

from oracle import edgegan_oracle as O
from ref_ops import RefOps

from edgegan_b200 import nn


def test_discriminator_built_from_nn_functions():
    cfg = O.Config(batch_size=2, output_height=32, output_width=64, multiclasses=False)
    rs = np.random.RandomState(0)
    v = O.discriminator_variables(cfg, "D", 32, 64, rs)
    x = rs.uniform(-1, 1, (2, 32, 64, 3))
    ops = RefOps(torch.float64)
    with nn.variable_context(ops, variables=v) as ctx:
        with nn.variable_scope("D"):
            # discriminator.py:61-76
            D = nn.conv_block(ops.from_numpy(x), 64, "d_conv_0", 4, 2, True, False, None, "lrelu")
            D = nn.conv_block(D, 128, "d_conv_1", 4, 2, True, False, "instance", "lrelu")
            D = nn.conv_block(D, 256, "d_conv_3", 4, 2, True, False, "instance", "lrelu")
            D = nn.conv_block(D, 512, "d_conv_4", 4, 2, True, False, "instance", "lrelu")
            d = nn.linear(D.reshape(2, -1), 1, name="d_linear_5")
        assert set(ctx.variables) == set(v)
    vt = {k: torch.tensor(a, dtype=torch.float64) for k, a in v.items()}
    _, want = O.discriminator(vt, "D", torch.tensor(x))
    assert np.abs(d.numpy() - want.numpy()).max() < 1e-10


def test_generator_built_from_nn_functions_reproduces_the_batch_norm_quirk():
    cfg = O.Config(batch_size=3, output_height=32, output_width=64, multiclasses=False)
    rs = np.random.RandomState(1)
    v = O.generator_variables(cfg, "G1", rs)
    z = rs.normal(size=(3, 100))
    ops = RefOps(torch.float64)
    with nn.variable_context(ops, variables=v) as ctx:
        with nn.variable_scope("G1"):
            # generator.py:48-74; `nn.norm(h0, self._norm)` binds 'instance' to is_train -> BATCH norm
            z_ = nn.linear(ops.from_numpy(z), 64 * 8 * 2 * 2, name="g_lin_0")
            h0 = z_.reshape(-1, 2, 2, 512)
            h0 = nn.activation_fn(nn.norm(h0, "instance"), "relu")
            h1 = nn.deconv_block(h0, [3, 4, 4, 256], "g_dconv_1", 5, 2, True, False, "instance", "relu")
            h2 = nn.deconv_block(h1, [3, 8, 8, 128], "g_dconv_2", 5, 2, True, False, "instance", "relu")
            h3 = nn.deconv_block(h2, [3, 16, 16, 64], "g_dconv_3", 5, 2, True, False, "instance", "relu")
            h4 = nn.deconv_block(h3, [3, 32, 32, 3], "g_dconv_4", 5, 2, True, False, None, None)
            out = nn.activation_fn(h4, "tanh")
        assert set(ctx.variables) == set(v)
    vt = {k: torch.tensor(a, dtype=torch.float64) for k, a in v.items()}
    want = O.generator(vt, "G1", torch.tensor(z))
    assert np.abs(out.numpy() - want.numpy()).max() < 1e-10


def test_residual_and_mlp():
    rs = np.random.RandomState(2)
    x = rs.standard_normal((2, 8, 8, 16))
    ops = RefOps(torch.float64)
    with nn.variable_context(ops, seed=5) as ctx:
        y = nn.residual(ops.from_numpy(x), 32, "blk", True, False, "instance", bias=True)
        f = nn.mlp(y.reshape(2, -1)[:, :64].contiguous(), 10, "fc", True, False)
        pooled = nn.mean_pool(y)
    v = {k: torch.tensor(ops.to_numpy(t)) for k, t in ctx.variables.items()}
    xt = torch.tensor(x)
    o = O.conv2d(xt, v["blk/res1/conv2d/w"], v["blk/res1/conv2d/b"], 1, "REFLECT")
    o = torch.relu(O.instance_norm(o))
    o = O.instance_norm(O.conv2d(o, v["blk/res2/conv2d/w"], v["blk/res2/conv2d/b"], 1, "REFLECT"))
    sc = O.conv2d(xt, v["blk/shortcut/conv2d/w"], v["blk/shortcut/conv2d/b"], 1, "REFLECT")
    want = torch.relu(sc + o)
    assert np.abs(y.numpy() - want.numpy()).max() < 1e-10
    assert np.abs(f.numpy() - (want.reshape(2, -1)[:, :64] @ v["fc/w"] + v["fc/b"]).numpy()).max() < 1e-10
    assert np.abs(pooled.numpy() - O.avg_pool_same(want, 2).numpy()).max() < 1e-12


def test_classifier_built_from_nn_functions():
    """models/classifier.py:12-119 assembled from nn.conv2d2 / mru_conv / mean_pool / fully_connected / prelu: the
    auto-generated scope names (Conv, Conv_1, Conv_2, Conv_3 per unit) and the logits match the oracle."""
    cfg = O.Config(batch_size=2, output_height=32, output_width=64, multiclasses=True)
    v, u = O.init_variables(cfg, seed=4)
    v = {k: a for k, a in v.items() if k.startswith("D2/")}
    rs = np.random.RandomState(3)
    x = rs.uniform(-1, 1, (2, 32, 32, 3))
    ops = RefOps(torch.float64)
    given = dict(v)
    given.update(u)
    with nn.variable_context(ops, variables=given) as ctx:
        xt = ops.from_numpy(x)
        x_list = [xt]
        for _ in range(5):
            x_list.append(nn.mean_pool(x_list[-1], data_format="NCHW"))
        x_list = x_list[::-1]
        act, winit, size = nn.prelu, 0.02, 64
        with nn.variable_scope("D2"):
            h0 = nn.conv2d2(x_list[-1], 8, kernel_size=7, sn=True, stride=1, data_format="NCHW", activation_fn=act,
                            weights_initializer=winit)
            hts = [h0]
            for t, mult in enumerate((2, 4, 8, 12), start=1):
                hts = nn.mru_conv(x_list[-t], hts, size * mult, sn=True, stride=2, dilate_rate=1, data_format="NCHW",
                                  num_blocks=1, last_unit=(t == 4), activation_fn=act, weights_initializer=winit,
                                  unit_num=t)
            img = hts[-1]
            disc = nn.conv2d2(img, 1, kernel_size=1, sn=True, stride=1, data_format="NCHW", activation_fn=None,
                              weights_initializer=winit)
            n, H, W, C = img.shape
            feat = ops.empty((n, C))
            ops.globalmean_fwd(img, feat)
            logits = nn.fully_connected(feat, cfg.num_classes, sn=True, activation_fn=None)
        names = {k for k in ctx.variables}
        assert {k for k in names if not k.endswith("/u")} == set(v)
        assert {k for k in names if k.endswith("/u")} == set(u)
    vt = {k: torch.tensor(a, dtype=torch.float64) for k, a in v.items()}
    ut = {k: torch.tensor(a, dtype=torch.float64) for k, a in u.items()}
    want = O.classifier(vt, ut, "D2", torch.tensor(x).permute(0, 3, 1, 2))
    assert tuple(disc.shape) == (2, 2, 2, 1)          # 32 -> 2 after the four stride-2 units; the head is unused
    assert np.abs(logits.numpy() - want.numpy()).max() < 1e-9


def test_classifier_call_returns_the_reference_triple():
    """Classifier.__call__ (classifier.py:12-119): (disc, sigmoid(logits), logits); disc and logits equal the network
    assembled from the nn functions on the same variables."""
    from edgegan_b200.models.classifier import Classifier, classifier_specs
    from edgegan_b200.variables import ParamStore
    cfg = O.Config(batch_size=2, output_height=32, output_width=64, multiclasses=True)
    v, u = O.init_variables(cfg, seed=6)
    given = {k: a for k, a in v.items() if k.startswith("D2/")}
    given.update(u)
    rs = np.random.RandomState(8)
    x = rs.uniform(-1, 1, (2, 32, 32, 3))
    ops = RefOps(torch.float64)
    store = ParamStore(ops, classifier_specs("D2", cfg.num_classes), np.random.RandomState(0))
    clf = Classifier("D2", ops=ops, store=store, num_classes=cfg.num_classes)
    store.load({k: a for k, a in given.items() if k in store.offsets})
    clf.aux.load({k: a for k, a in given.items() if k in clf.aux.offsets}, strict=True)
    disc, prob, logits = clf(ops.from_numpy(x).permute(0, 3, 1, 2), cfg.num_classes)
    with nn.variable_context(ops, variables=given):
        xt = ops.from_numpy(x)
        pyr = [xt]
        for _ in range(3):
            pyr.append(nn.mean_pool(pyr[-1], data_format="NCHW"))
        with nn.variable_scope("D2"):
            hts = [nn.conv2d2(xt, 8, kernel_size=7, sn=True, activation_fn=nn.prelu, weights_initializer=0.02)]
            for t, mult in enumerate((2, 4, 8, 12), start=1):
                hts = nn.mru_conv(pyr[t - 1], hts, 64 * mult, sn=True, stride=2, num_blocks=1, last_unit=(t == 4),
                                  activation_fn=nn.prelu, weights_initializer=0.02, unit_num=t)
            want_disc = nn.conv2d2(hts[-1], 1, kernel_size=1, sn=True, activation_fn=None, weights_initializer=0.02)
    assert tuple(disc.shape) == tuple(want_disc.shape) == (2, 2, 2, 1)
    assert np.abs(disc.numpy() - want_disc.numpy()).max() < 1e-10
    assert np.abs(prob.numpy() - 1 / (1 + np.exp(-logits.numpy()))).max() < 1e-12
    vt = {k: torch.tensor(a, dtype=torch.float64) for k, a in v.items() if k.startswith("D2/")}
    ut = {k: torch.tensor(a, dtype=torch.float64) for k, a in u.items()}
    want = O.classifier(vt, ut, "D2", torch.tensor(x).permute(0, 3, 1, 2))
    assert np.abs(logits.numpy() - want.numpy()).max() < 1e-9
