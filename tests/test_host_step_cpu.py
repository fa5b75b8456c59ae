"""Host-side logic (edgegan_b200/models/*.py: hand-derived backward + WGAN-GP double backward, run order,
loss scaling) executed on the CPU reference operator set and compared with the oracle's autograd."""
import numpy as np
import pytest
import torch

from oracle import edgegan_oracle as O
from ref_ops import RefOps

from edgegan_b200.config import Flags
from edgegan_b200.models.edgegan import EdgeGAN


def small_cfg(multiclass=False, B=2):
    ocfg = O.Config(batch_size=B, output_height=32, output_width=64, multiclasses=multiclass,
                    image_dis_size=64, edge_dis_size=64)
    flags = Flags(batch_size=B, input_height=32, input_width=64, output_height=32, output_width=64,
                  multiclasses=multiclass, image_dis_size=64, edge_dis_size=64)
    if not multiclass:
        flags.num_classes = None
    return ocfg, flags


def run_both(multiclass, runs=None, B=2):
    torch.manual_seed(0)
    ocfg, flags = small_cfg(multiclass, B)
    v, u = O.init_variables(ocfg, seed=3)
    inp = O.make_inputs(ocfg, seed=11)
    st = O.OracleState(ocfg, v, u, dtype=torch.float64)
    collect = {}
    O.update_model(st, inp, runs=runs, collect=collect)

    ops = RefOps(torch.float64)
    m = EdgeGAN(None, flags, None, ops=ops)
    m.build_train_model()
    allv = dict(v)
    allv.update(u)
    m.load_variables(allv)
    grads = {}

    def hook(run, model):
        grads[run] = model.export_variables("grad")
    m.run_hook = hook
    m.update_model(ops.from_numpy(inp.images), ops.from_numpy(inp.z), ops.from_numpy(inp.alpha), inp.eps, runs=runs)
    return st, collect, m, grads


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


# gradients that are mathematically zero (a bias feeding an instance norm): numerical noise only
def cancelled(name):
    return (name.endswith("deconv2d/b") and "g_dconv_4" not in name) or ("/res" in name and name.endswith("conv2d/b"))


@pytest.mark.parametrize("multiclass,runs", [(False, ["d_optim"]), (False, ["d_optim_patch2"]), (False, ["g_optim_u"]),
                                             (False, ["e_optim"]), (True, ["d_optim2"]), (True, ["g_optim_u"])])
def test_single_run_gradients_match_oracle(multiclass, runs):
    st, collect, m, grads = run_both(multiclass, runs)
    run = runs[0]
    for name, g in collect[run]["grads"].items():
        mine = grads[run][name]
        # mathematically-zero gradients: biases feeding an instance norm, and the update-gate bias when the whole
        # gate pre-activation is positive (min-max normalisation is shift invariant)
        if cancelled(name) or np.abs(g).max() < 1e-12:
            assert np.abs(mine).max() < 1e-9, name
            continue
        assert rel_err(mine, g) < 1e-9, (run, name, rel_err(mine, g))


def test_full_multi_class_step_matches_oracle():
    """all 7 runs (incl. the classifier run d_optim2 and the CE term of the image generator), 14 classes"""
    st, collect, m, grads = run_both(True)
    new = m.export_variables("var")
    for name, t in st.v.items():
        if name not in new:                    # the unused disc head lives in the classifier's aux store
            continue
        tol = 1e-12 if (cancelled(name) or "update_gate/biases" in name) else 1e-9 * max(1.0, np.abs(t.numpy()).max())
        assert np.abs(new[name] - t.numpy()).max() <= tol, name
    losses = m.read_losses()
    assert abs(losses["loss_d_ac"] - st.losses["d_optim2"]) < 1e-9
    assert abs(losses["image_gloss_b"] - st.losses["g_optim_b/image_gloss"]) < 1e-9


def test_full_single_class_step_matches_oracle():
    st, collect, m, grads = run_both(False)
    new = m.export_variables("var")
    for name, t in st.v.items():
        if cancelled(name):      # zero gradient up to rounding noise: the variable stays (numerically) at its init
            assert np.abs(new[name] - t.numpy()).max() < 1e-12, name
            continue
        assert rel_err(new[name], t.numpy()) < 1e-9, name
    losses = m.read_losses()
    assert abs(losses["joint_dis_dloss"] - st.losses["d_optim"]) < 1e-9 * max(1, abs(st.losses["d_optim"]))
    assert abs(losses["image_dis_dloss"] - st.losses["d_optim_patch2"]) < 1e-9 * max(1, abs(st.losses["d_optim_patch2"]))
    assert abs(losses["edge_dis_dloss"] - st.losses["d_optim_patch3"]) < 1e-9 * max(1, abs(st.losses["d_optim_patch3"]))
    assert abs(losses["zl_loss"] - st.losses["e_optim"]) < 1e-9 * max(1, abs(st.losses["e_optim"]))
    assert abs(losses["edge_gloss"] - st.losses["g_optim_u/edge_gloss"]) < 1e-9
    assert abs(losses["image_gloss_b"] - st.losses["g_optim_b/image_gloss"]) < 1e-9


def test_teacher_forced_checker_on_the_cpu_operator_set():
    """tests/parity_util.teacher_forced_step (the whole-step checker of the -m gpu tests) exercised without a GPU: the
    fp64 CPU operator set must pass every run strictly (no tensor through the noise clause), and a deliberately
    corrupted run must be caught."""
    from parity_util import teacher_forced_step
    ocfg, flags = small_cfg(True, 2)
    v, u = O.init_variables(ocfg, seed=3)
    inp = O.make_inputs(ocfg, seed=11)

    def build():
        ops = RefOps(torch.float64)
        m = EdgeGAN(None, flags, None, ops=ops)
        m.build_train_model()
        allv = dict(v)
        allv.update(u)
        m.load_variables(allv)
        return m, ops

    m, ops = build()
    stats, fails = teacher_forced_step(m, ops, ocfg, v, u, inp, 1e-6, weight_tol=1e-4, sens_samples=0, log=lambda *_: None)
    assert not fails, fails[:3]
    assert len(stats) == 7 and all(r["noise_clause"] == 0 for r in stats.values())
    # corrupt the encoder gradient scale: the loss weight is read by run 6 only
    m, ops = build()
    m.config.stage1_zl_loss = 10.5
    stats, fails = teacher_forced_step(m, ops, ocfg, v, u, inp, 1e-6, weight_tol=1e-4, sens_samples=0, log=lambda *_: None)
    assert fails and all(f[1] == "e_optim" for f in fails), fails[:3]
