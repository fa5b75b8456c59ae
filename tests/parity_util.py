"""Shared parity helpers (test infrastructure).

The reference graph is fp32 and parts of it are ill-conditioned (the WGAN-GP double backward divides by the
per-channel standard deviation up to the third power), so an fp32 run is only reproducible up to its own
rounding noise.  We therefore evaluate the oracle twice -- fp64 ("truth") and fp32 (the reference precision) --
and accept a device result when its distance to the fp64 truth is within `tol` of the tensor's magnitude OR
within `noise_factor` x the fp32 oracle's own distance to that truth."""
import numpy as np
import torch

from oracle import edgegan_oracle as O


def cancelled(name):
    """gradients that are mathematically zero (a bias feeding an instance norm)"""
    return (name.endswith("deconv2d/b") and "g_dconv_4" not in name) or ("/res" in name and name.endswith("conv2d/b"))


def oracle_pair(ocfg, v, u, inp, runs=None):
    out = {}
    for dt in (torch.float64, torch.float32):
        st = O.OracleState(ocfg, v, u, dtype=dt)
        col = {}
        O.update_model(st, inp, runs=runs, collect=col)
        out[dt] = (st, col)
    return out[torch.float64], out[torch.float32]


def oracle_sensitivity(ocfg, v, u, inp, col64, rel=1e-5, seed=0, runs=None, samples=2, rms=None):
    """max over `samples` random perturbations of _oracle_sensitivity_once (the response is heavy-tailed: a
    single lrelu-mask or arg-max flip moves a whole filter gradient by a finite amount)."""
    out = None
    for k in range(samples):
        cur = _oracle_sensitivity_once(ocfg, v, u, inp, col64, rel, seed + k, runs, rms)
        w = cur.pop("__weights__")
        if out is None:
            out = cur
            out["__weights__"] = [w]
        else:
            for run in cur:
                for n in out[run]:
                    out[run][n] = max(out[run][n], cur[run][n])
            out["__weights__"].append(w)
    return out


def _oracle_sensitivity_once(ocfg, v, u, inp, col64, rel, seed, runs, rms=None):
    """How far the fp64 oracle's own gradients move when the images and z are perturbed by `rel` (relative): the
    critics' penalty gradient is discontinuous in the lrelu masks of low-variance instance-norm channels, so
    near such a point ANY two fp32 implementations disagree by a finite amount.  run -> name -> rel. change."""
    rs = np.random.RandomState(seed)
    img = (inp.images.astype(np.float64) * (1 + rel * rs.standard_normal(inp.images.shape))).astype(np.float32)
    # the latent too (runs 5-7 do not read the images at all); the class-id column stays exact
    z = inp.z.astype(np.float64).copy()
    z[:, :ocfg.z_dim] *= 1 + rel * rs.standard_normal((z.shape[0], ocfg.z_dim))
    inp2 = O.StepInputs(img, z.astype(np.float32), inp.alpha, inp.eps)
    st = O.OracleState(ocfg, v, u, dtype=torch.float64)
    if rms is not None:
        for n in st.rms:
            st.rms[n] = torch.tensor(np.asarray(rms[n], np.float64).reshape(tuple(st.rms[n].shape)), dtype=torch.float64)
    col = {}
    O.update_model(st, inp2, runs=runs, collect=col)
    out = {}
    for run, rec in col64.items():
        out[run] = {n: maxabs(col[run]["grads"][n] - g) / (maxabs(g) + 1e-30) for n, g in rec["grads"].items()}
        out[run] = dict(out[run])
    # the perturbed run's end weights (torch tensors): RMSProp normalises the step, so where |g| >> 1 a small relative
    # change of a gradient element moves its weight by O(lr) -- the weight check needs this response too
    out["__weights__"] = {n: t.detach().numpy().astype(np.float64) for n, t in st.v.items()}
    return out


def maxabs(a):
    return float(np.abs(np.asarray(a, np.float64)).max())


def check_grads(mine, truth, ref32, tol, noise_factor=10.0, sens=None, stats=None, run_level=False, abs_tol=None):
    """mine/truth/ref32: run -> name -> array; sens: oracle_sensitivity().  A tensor passes when its max-abs error
    relative to max|g| is <= tol, or <= noise_factor x max(fp32-oracle noise, oracle sensitivity) of that tensor.
    run_level=True adds a third clause for the large / realistic-input tests: <= noise_factor x the LARGEST such
    instability of any tensor of the same run.  Reason (measured, tools/classifier_diff.py): with 10^5-10^6 activations
    per layer some pre-activation always lies within the fp32 rounding error of a prelu / lrelu / relu kink, and which
    side it falls on differs between two correct fp32 implementations; one flipped element changes every gradient
    upstream of it by ~1e-3..1e-2 of max|g|.  The perturbed fp64 oracle flips other elements than the device does, so
    the per-tensor estimate can miss an effect that the run-wide estimate captures.
    abs_tol (name -> float, optional): a gradient whose ABSOLUTE max error is below abs_tol[name] passes ("abs_clause"):
    teacher_forced_step sets it to 0.005 * sqrt(0.9 * min rms), i.e. an error that moves no weight by more than 0.005
    learning rates through RMSProp -- ten times below the bar of the weight check.  This is for the scalar prelu
    leaks, whose gradients are heavily cancelling sums (|sum| << sum|terms|): any upstream noise is large relative to
    such a value and irrelevant to the update.
    `stats` (dict, optional) receives run -> {"tensors", "strict", "noise_clause", "run_clause", "worst", ...}: how many
    tensors passed on the plain tolerance and how many only through each noise clause."""
    fails, report = [], {}
    for run, rec in truth.items():
        if run.startswith("__"):
            continue
        worst = 0.0
        rows = []
        for name, g64 in rec["grads"].items():
            if cancelled(name):
                continue
            scale = maxabs(g64) + 1e-30
            e_abs = maxabs(np.asarray(mine[run][name], np.float64) - g64)
            e = e_abs / scale
            noise = maxabs(np.asarray(ref32[run]["grads"][name], np.float64) - g64) / scale
            if sens is not None:
                noise = max(noise, sens[run][name])
            rows.append((name, e, noise, e_abs))
        run_noise = max([r[2] for r in rows], default=0.0) if run_level else 0.0
        for name, e, noise, e_abs in rows:
            how = "strict" if e <= tol else ("noise_clause" if e <= noise_factor * noise else
                                             ("run_clause" if e <= noise_factor * run_noise else "fail"))
            if how == "fail" and abs_tol is not None and e_abs <= abs_tol.get(name, 0.0):
                how = "abs_clause"
            if stats is not None:
                st = stats.setdefault(run, {"tensors": 0, "strict": 0, "noise_clause": 0, "run_clause": 0, "abs_clause": 0, "worst": 0.0,
                                            "worst_name": "", "worst_noise": 0.0, "run_noise": run_noise})
                st["tensors"] += 1
                if how != "fail":
                    st[how] += 1
                if e >= st["worst"]:
                    st["worst"], st["worst_name"], st["worst_noise"] = e, name, noise
            worst = max(worst, e)
            if how == "fail":
                fails.append((run, name, e, noise))
        report[run] = worst
    return report, fails


def check_weights(new, st64, st32, lr, tol, noise_factor=10.0, only=None, stats=None, sens=None, run_level=False):
    """updated weights in lr-units (one RMSProp step moves a weight by at most ~lr / sqrt(0.1) ~ 3.2 lr): error vs the
    fp64 oracle <= tol, or <= noise_factor x max(the fp32 oracle's own distance to it, the fp64 oracle's response to
    the input perturbation of oracle_sensitivity -- `sens`); run_level: or <= noise_factor x the largest such
    instability among the checked tensors (see check_grads).  `only`: name-prefix filter."""
    fails, rows = [], []
    for name, t in st64.v.items():
        if cancelled(name) or (only is not None and not name.startswith(only)):
            continue
        e = maxabs(np.asarray(new[name], np.float64).reshape(t.shape) - t.numpy()) / lr
        noise = maxabs(st32.v[name].numpy().astype(np.float64) - t.numpy()) / lr
        if sens is not None:
            for w in sens.get("__weights__", []):
                noise = max(noise, maxabs(w[name] - t.numpy()) / lr)
        rows.append((name, e, noise))
    run_noise = max([r[2] for r in rows], default=0.0) if run_level else 0.0
    for name, e, noise in rows:
        how = "strict" if e <= tol else ("noise_clause" if e <= noise_factor * noise else
                                         ("run_clause" if e <= noise_factor * run_noise else "fail"))
        if stats is not None:
            stats["tensors"] = stats.get("tensors", 0) + 1
            if how != "fail":
                stats[how] = stats.get(how, 0) + 1
            if e >= stats.get("worst", 0.0):
                stats["worst"], stats["worst_name"], stats["worst_noise"] = e, name, noise
        if how == "fail":
            fails.append((name, e, noise))
    return fails


RUN_SCOPES = {"d_optim": ("D/",), "d_optim_patch2": ("D_patch2/",), "d_optim_patch3": ("D_patch3/",), "d_optim2": ("D2/",),
              "g_optim_u": ("G1/", "G2/"), "e_optim": ("E/",), "g_optim_b": ("G1/", "G2/")}


def teacher_forced_step(m, ops, ocfg, v, u, inp, grad_tol, weight_tol=0.05, sens_samples=1, log=print, run_level=False):
    """Whole update_model on the device, every run checked STRICTLY: the hook exports the device's weights, RMSProp
    slots and gradients right before each run's apply; afterwards each run is replayed by the fp64 oracle (truth) and
    the fp32 oracle (the reference precision) from exactly those device weights, so no run inherits the chaotic drift
    of an earlier one (which made the old whole-step bound for runs 5-7 meaningless) while the sequence itself
    (fresh / stale generator outputs, weights handed from run to run, slot sharing of runs 5 and 7) is still the
    device's.  Per run: gradients within grad_tol(run) * max|g| or the noise clause of check_grads; the weights that
    run's RMSProp wrote within weight_tol lr-units or 4x the fp32 oracle's own distance to the fp64 one.
    -> {run: stats} for the parity report."""
    snaps = []

    def hook(run, model):
        snaps.append((run, model.export_variables("var"), model.export_variables("ms"), model.export_variables("grad")))

    m.run_hook = hook
    m.update_model(ops.from_numpy(inp.images), ops.from_numpy(inp.z), ops.from_numpy(inp.alpha), inp.eps)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    m.run_hook = None
    final = m.export_variables("var")
    dev_losses = m.read_losses()
    loss_of = {"d_optim": ("joint_dis_dloss",), "d_optim_patch2": ("image_dis_dloss",), "d_optim_patch3": ("edge_dis_dloss",),
               "d_optim2": ("loss_d_ac",), "g_optim_u": ("edge_gloss", "image_gloss"), "e_optim": ("zl_loss",),
               "g_optim_b": ("edge_gloss_b", "image_gloss_b")}
    lr = ocfg.learning_rate
    out, all_fails = {}, []
    for k, (run, var, ms, grad) in enumerate(snaps):
        after = snaps[k + 1][1] if k + 1 < len(snaps) else final
        var = {n: np.asarray(var[n]).reshape(np.asarray(v[n]).shape) for n in v}
        pair = {}
        for dt in (torch.float64, torch.float32):
            st = O.OracleState(ocfg, var, u, dtype=dt)
            for n in st.rms:
                st.rms[n] = torch.tensor(np.asarray(ms[n], np.float64).reshape(tuple(st.rms[n].shape)), dtype=dt)
            col = {}
            O.update_model(st, inp, runs=[run], collect=col)
            pair[dt] = (st, col)
        (st64, col64), (st32, col32) = pair[torch.float64], pair[torch.float32]
        for name in list(col64[run]["grads"]):                 # mathematically-zero gradients
            if maxabs(col64[run]["grads"][name]) < 1e-9:
                for c in (col64, col32):
                    c[run]["grads"].pop(name)
        sens = oracle_sensitivity(ocfg, var, u, inp, col64, runs=[run], samples=sens_samples, rms=ms) if sens_samples else None
        mine = {run: {n: np.asarray(g).reshape(col64[run]["grads"][n].shape) for n, g in grad.items() if n in col64[run]["grads"]}}
        gstats = {}
        tol = grad_tol(run) if callable(grad_tol) else grad_tol
        abs_tol = {n: 0.005 * float(np.sqrt(0.9 * max(float(np.min(ms[n])), 0.0))) for n in mine[run]}
        _, fails = check_grads(mine, col64, col32, tol, sens=sens, stats=gstats, run_level=run_level, abs_tol=abs_tol)
        wstats = {}
        wfails = []
        for scope in RUN_SCOPES[run]:
            wfails += check_weights(after, st64, st32, lr, weight_tol, only=scope, stats=wstats, sens=sens, run_level=run_level)
        rec = dict(gstats.get(run, {}))
        loss_dev, loss_ref = sum(dev_losses[n] for n in loss_of[run]), st64.losses.get(run)
        rec.update(grad_tol=tol, loss_dev=loss_dev, loss_ref=loss_ref, weights=wstats)
        if not abs(loss_dev - loss_ref) <= 2e-3 * max(1.0, abs(loss_ref)):
            all_fails.append(("loss", run, loss_dev, loss_ref))
        out[f"{k + 1}:{run}"] = rec
        log(f"run {k + 1} {run}: grads {rec.get('tensors')} tensors, {rec.get('strict')} within {tol:g}, "
            f"{rec.get('noise_clause')} via the noise clause, {rec.get('run_clause')} via the run-level clause, {rec.get('abs_clause')} below the absolute bar; worst {rec.get('worst', 0):.2e} ({rec.get('worst_name')}, fp32-oracle "
            f"noise/sensitivity {rec.get('worst_noise', 0):.2e}); weights worst {wstats.get('worst', 0):.3f} lr-units "
            f"({wstats.get('worst_name')}, fp32-oracle noise {wstats.get('worst_noise', 0):.3f}), {wstats.get('noise_clause', 0)} via noise clause")
        all_fails += [("grad",) + f for f in fails] + [("weight", run) + f for f in wfails]
    return out, all_fails
