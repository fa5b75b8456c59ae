"""Diagnostic: per-run / per-tensor gradient error of the device step vs the fp64 oracle, next to the fp32
oracle's own rounding noise, for each conv algorithm; and generator-output error of the inference path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import edgegan_oracle as O
from parity_util import oracle_pair, cancelled, maxabs
from test_step_gpu import make

B = int(os.environ.get("B", "4"))
SEED = int(os.environ.get("SEED", "11"))
WSEED = int(os.environ.get("WSEED", "3"))
algos = sys.argv[1:] or ["simt", "tc"]
ocfg = O.Config(batch_size=B, multiclasses=False)
v, u = O.init_variables(ocfg, seed=WSEED)
inp = O.make_inputs(ocfg, seed=SEED)
(st64, col64), (st32, col32) = oracle_pair(ocfg, v, u, inp)
for algo in algos:
    _, _, _, m, ops = make(B, False, algo.split(":")[0], seed=WSEED)
    if ":" in algo:       # e.g. simt:fwd=tc3x,wgrad=tc3x  -> per-pass override on top of the default
        for kv in algo.split(":")[1].split(","):
            k, val = kv.split("=")
            ops.pass_algo[k] = val
    grads = {}
    m.run_hook = lambda run, model: grads.__setitem__(run, model.export_variables("grad"))
    m.update_model(ops.from_numpy(inp.images), ops.from_numpy(inp.z), ops.from_numpy(inp.alpha), inp.eps)
    torch.cuda.synchronize()
    print(f"==== algo {algo}")
    for run, rec in col64.items():
        rows = []
        for name, g64 in rec["grads"].items():
            if cancelled(name):
                continue
            sc = maxabs(g64) + 1e-30
            e = maxabs(np.asarray(grads[run][name], np.float64) - g64) / sc
            nz = maxabs(np.asarray(col32[run]["grads"][name], np.float64) - g64) / sc
            rows.append((e, nz, name, sc))
        rows.sort(reverse=True)
        print(run, " worst:", ["%s e=%.2e noise=%.2e |g|=%.1e" % (r[2].split("/", 1)[1], r[0], r[1], r[3]) for r in rows[:3]])
    new = m.export_variables("var")
    lr = ocfg.learning_rate
    rows = []
    for name, t in st64.v.items():
        if cancelled(name):
            continue
        rows.append((maxabs(new[name].astype(np.float64) - t.numpy()) / lr, maxabs(st32.v[name].numpy().astype(np.float64) - t.numpy()) / lr, name))
    rows.sort(reverse=True)
    print("weights (x lr):", ["%s e=%.3f noise=%.3f" % (r[2], r[0], r[1]) for r in rows[:5]])
    # inference
    ocfg1 = O.Config(batch_size=1, multiclasses=False)
    _, _, _, m1, ops1 = make(1, False, algo.split(":")[0], seed=WSEED)
    x = np.random.RandomState(2333).uniform(-1, 1, (1, 64, 128, 3)).astype(np.float32)
    s64, s32 = O.OracleState(ocfg1, v, u, dtype=torch.float64), O.OracleState(ocfg1, v, u)
    for eps in (0.0, 1.0):
        e64, i64 = O.test_forward(s64, x, eps=eps)
        e32, i32 = O.test_forward(s32, x, eps=eps)
        e, i = m1.test_forward(ops1.from_numpy(x), eps=eps)
        e, i = ops1.to_numpy(e), ops1.to_numpy(i)
        print(f"inference eps={eps}: max-abs vs fp64 edge {np.abs(e-e64).max():.2e} image {np.abs(i-i64).max():.2e}; "
              f"fp32 oracle noise {np.abs(e32-e64).max():.2e}; mse {((e-e64)**2).mean():.2e}")
