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


def oracle_sensitivity(ocfg, v, u, inp, col64, rel=1e-5, seed=0, runs=None, samples=2):
    """max over `samples` random perturbations of _oracle_sensitivity_once (the response is heavy-tailed: a
    single lrelu-mask or arg-max flip moves a whole filter gradient by a finite amount)."""
    out = None
    for k in range(samples):
        cur = _oracle_sensitivity_once(ocfg, v, u, inp, col64, rel, seed + k, runs)
        if out is None:
            out = cur
        else:
            for run in out:
                for n in out[run]:
                    out[run][n] = max(out[run][n], cur[run][n])
    return out


def _oracle_sensitivity_once(ocfg, v, u, inp, col64, rel, seed, runs):
    """How far the fp64 oracle's own gradients move when the images are perturbed by `rel` (relative): the
    critics' penalty gradient is discontinuous in the lrelu masks of low-variance instance-norm channels, so
    near such a point ANY two fp32 implementations disagree by a finite amount.  run -> name -> rel. change."""
    rs = np.random.RandomState(seed)
    img = (inp.images.astype(np.float64) * (1 + rel * rs.standard_normal(inp.images.shape))).astype(np.float32)
    inp2 = O.StepInputs(img, inp.z, inp.alpha, inp.eps)
    st = O.OracleState(ocfg, v, u, dtype=torch.float64)
    col = {}
    O.update_model(st, inp2, runs=runs, collect=col)
    out = {}
    for run, rec in col64.items():
        out[run] = {n: maxabs(col[run]["grads"][n] - g) / (maxabs(g) + 1e-30) for n, g in rec["grads"].items()}
        out[run] = dict(out[run])
    return out


def maxabs(a):
    return float(np.abs(np.asarray(a, np.float64)).max())


def check_grads(mine, truth, ref32, tol, noise_factor=10.0, sens=None):
    """mine/truth/ref32: run -> name -> array; sens: oracle_sensitivity().  A tensor passes when its max-abs error
    relative to max|g| is <= tol, or <= noise_factor x max(fp32-oracle noise, oracle sensitivity)."""
    fails, report = [], {}
    for run, rec in truth.items():
        worst = 0.0
        for name, g64 in rec["grads"].items():
            if cancelled(name):
                continue
            scale = maxabs(g64) + 1e-30
            e = maxabs(np.asarray(mine[run][name], np.float64) - g64) / scale
            noise = maxabs(np.asarray(ref32[run]["grads"][name], np.float64) - g64) / scale
            if sens is not None:
                noise = max(noise, sens[run][name])
            worst = max(worst, e)
            if not (e <= tol or e <= noise_factor * noise):
                fails.append((run, name, e, noise))
        report[run] = worst
    return report, fails


def check_weights(new, st64, st32, lr, tol, noise_factor=4.0):
    fails = []
    for name, t in st64.v.items():
        if cancelled(name):
            continue
        e = maxabs(np.asarray(new[name], np.float64) - t.numpy()) / lr
        noise = maxabs(st32.v[name].numpy().astype(np.float64) - t.numpy()) / lr
        if not (e <= tol or e <= noise_factor * noise):
            fails.append((name, e, noise))
    return fails
