"""Generates tests/golden/*.npz|json from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden vectors and TensorFlow 1.14 cannot run here, so these fixtures pin
the ORACLE (oracle/edgegan_oracle.py) against accidental change; they are not outputs of the reference itself
("parity unpinned").  Inputs are regenerated from seeds, only outputs are stored."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import edgegan_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def op_vectors():
    rs = np.random.RandomState(7)
    t = lambda *s: torch.tensor(rs.standard_normal(s).astype(np.float32))
    out = {}
    x, w, b = t(2, 4, 4, 8), t(5, 5, 6, 8) * 0.1, t(6)
    out["deconv"] = O.deconv2d(x, w, b).numpy()
    xc, wc = t(2, 8, 8, 3), t(4, 4, 3, 5) * 0.1
    out["conv_same_s2"] = O.conv2d(xc, wc, None, 2, "SAME").numpy()
    out["conv_reflect"] = O.conv2d(t(1, 6, 6, 4), t(3, 3, 4, 5) * 0.1, t(5), 1, "REFLECT").numpy()
    out["bicubic"] = O.bicubic_up2(t(1, 6, 6, 2), 12).numpy()
    out["instance_norm"] = O.instance_norm(t(2, 5, 5, 3)).numpy()
    out["avg_pool_8_on_2x2"] = O.avg_pool_same(t(2, 2, 2, 4), 8).numpy()
    return out


def step_summary(multiclass):
    cfg = O.Config(batch_size=2, output_height=32, output_width=64, multiclasses=multiclass,
                   image_dis_size=64, edge_dis_size=64)
    v, u = O.init_variables(cfg, seed=3)
    inp = O.make_inputs(cfg, seed=11)
    st = O.OracleState(cfg, v, u)
    O.update_model(st, inp)
    summ = {"losses": {k: float(x) for k, x in st.losses.items()}, "vars": {}}
    for k, tns in st.v.items():
        a = tns.numpy().astype(np.float64)
        d = a - np.asarray(v[k], np.float64)
        summ["vars"][k] = [float(a.sum()), float(np.abs(d).sum()), float(np.abs(d).max())]
    e, i = O.test_forward(st, inp.images[:1], classes=[3] if multiclass else None, eps=0.25)
    summ["test_forward"] = [float(e.sum()), float(np.abs(e).sum()), float(i.sum()), float(np.abs(i).sum())]
    return summ


if __name__ == "__main__":
    torch.set_num_threads(4)
    np.savez_compressed(os.path.join(HERE, "ops_small.npz"), **op_vectors())
    json.dump(step_summary(False), open(os.path.join(HERE, "step_small_single.json"), "w"), indent=1)
    if "--multi" in sys.argv:
        json.dump(step_summary(True), open(os.path.join(HERE, "step_small_multi.json"), "w"), indent=1)
    print("written")
