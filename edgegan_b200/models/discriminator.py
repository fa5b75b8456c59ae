"""Discriminator (reference edgegan/models/discriminator.py:6-21,58-81) with the explicit backward and
WGAN-GP double-backward passes that tf.gradients / RMSPropOptimizer.minimize derived implicitly
(reference edgegan/models/edgegan.py:32-42,277-308; nn/functional.py:26-29).

Network: 4x conv_block(4x4, stride 2, SAME, no bias) with 64/128/256/512 filters, the first without
norm, the others instance-norm, all followed by activation_fn('lrelu') = max(x, 0.2x); NHWC flatten;
linear -> logit d.  `__call__` returns (sigmoid(d), d) like the reference.

Notation used below for one critic step on a batch X = [real ; fake ; xhat] (3B samples):
  sweep 1  forward                         a_l = conv(h_{l-1}, W_l), h_l = lrelu(IN(a_l)), d = <h_4, w_5> + b_5
  sweep 2  first-order backward on xhat    dl_a_l ("delta"), g = d(sigmoid(d)+d)/dxhat           
  sweep 3  tangent sweep of the penalty    cotangents of the deltas ("db_*"), dW += wgrad(db_h_{l-1}, dl_a_l)
  sweep 4  ordinary backward, 3B samples   cotangent ab_l of a_l (+ IN second-order term on the xhat third)
"""
from __future__ import annotations

from ..variables import D_LAYERS


class Discriminator(object):
    def __init__(self, name, is_train=True, norm="instance", activation="lrelu", num_filters=64,
                 use_resnet=False, *, ops=None, store=None, in_hw=None):
        if use_resnet:
            raise NotImplementedError("if_resnet_d=True is outside the hot path (SURVEY.md 2.1)")
        if norm != "instance" or activation != "lrelu":
            raise NotImplementedError("only D_norm='instance' with lrelu is implemented")
        self.name = name
        self._is_train = is_train
        self._num_filters = num_filters
        self.ops = ops
        self.store = store
        self.in_hw = tuple(in_hw)
        self.var_list = store.names()
        self.W = [store.var[f"{name}/{l}/conv2d/w"] for l in D_LAYERS]
        self.gW = [store.g[f"{name}/{l}/conv2d/w"] for l in D_LAYERS]
        self.w5 = store.var[f"{name}/d_linear_5/Matrix"]
        self.b5 = store.var[f"{name}/d_linear_5/bias"]
        self.gw5 = store.g[f"{name}/d_linear_5/Matrix"]
        self.gb5 = store.g[f"{name}/d_linear_5/bias"]

    # ------------------------------------------------------------------------------------------
    def _shapes(self, n):
        H, W = self.in_hw
        nf = self._num_filters
        chans = [3, nf, nf * 2, nf * 4, nf * 8]
        return [(n, H >> l, W >> l, chans[l]) for l in range(5)]

    def forward(self, x, tag):
        """sweep 1.  x [n,H,W,3] -> cache with pre-norm conv outputs `a`, IN stats, activations `h`, logits `d`."""
        ops, n = self.ops, x.shape[0]
        shp = self._shapes(n)
        a, h, st = [None] * 5, [x] + [None] * 4, [None] * 5
        for l in range(1, 5):
            h[l] = ops.buf(f"{self.name}/{tag}/h{l}", shp[l])
            if l == 1:
                # no norm on the first layer: the activation is the conv's epilogue and a_1 is never stored -- lrelu
                # keeps the sign, so every later act'(a_1) is taken from h_1
                ops.conv_fwd(h[0], self.W[0], None, h[1], 2, 1, act="lrelu")
            else:
                a[l] = ops.buf(f"{self.name}/{tag}/a{l}", shp[l])
                ops.conv_fwd(h[l - 1], self.W[l - 1], None, a[l], 2, 1)
                st[l] = ops.buf(f"{self.name}/{tag}/st{l}", (n, shp[l][3], 2))
                ops.instnorm_fwd(a[l], h[l], st[l], "lrelu")
        d = ops.buf(f"{self.name}/{tag}/d", (n,))
        ops.rowdot_fwd(h[4], self.w5, self.b5, d)
        return {"a": a, "h": h, "st": st, "d": d, "n": n, "tag": tag}

    def __call__(self, input, reuse=False):
        """Reference call signature: returns (sigmoid(D), D) as [n,1] tensors."""
        c = self.forward(input, "call")
        prob = self.ops.buf(f"{self.name}/call/prob", (c["n"],))
        self.ops.act_fwd(c["d"], prob, "sigmoid")
        return prob.view(-1, 1), c["d"].view(-1, 1)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _sub(c, lo, hi):
        """view of a forward cache restricted to samples [lo:hi) (contiguous in NHWC)."""
        s = slice(lo, hi)
        return {"a": [None if t is None else t[s] for t in c["a"]],
                "h": [t[s] for t in c["h"]],
                "st": [None if t is None else t[s] for t in c["st"]],
                "d": c["d"][s], "n": hi - lo, "tag": c["tag"]}

    def backward(self, c, gd, tag, param_grads=True, accumulate=False, input_grad=False, addends=None):
        """sweep 4 (also the plain backward of the generator step).

        gd [n]: cotangent of the logits.  addends[l]: extra cotangent on a_l for the LAST
        `addends['n']` samples (the IN second-order term of the penalty).  Returns the input
        cotangent when input_grad."""
        ops, n = self.ops, c["n"]
        a, h, st = c["a"], c["h"], c["st"]
        if param_grads:
            ops.rowdot_bwd_weight(gd, h[4], self.gw5, self.gb5, accumulate)
        hb = ops.buf(f"{self.name}/{tag}/hb4", h[4].shape)
        ops.rowdot_bwd_input(gd, self.w5, hb)
        for l in range(4, 0, -1):
            if l == 1:
                ab = hb                      # the layer-2 input gradient below already carries act'(a_1)
            else:
                ab = ops.buf(f"{self.name}/{tag}/ab{l}", h[l].shape)
                if addends is None:
                    ops.instnorm_bwd(a[l], st[l], hb, None, ab, "lrelu")
                else:
                    m = n - addends["n"]
                    if m > 0:
                        ops.instnorm_bwd(a[l][:m], st[l][:m], hb[:m], None, ab[:m], "lrelu")
                    ops.instnorm_bwd(a[l][m:], st[l][m:], hb[m:], addends[l], ab[m:], "lrelu")
            if param_grads:
                ops.conv_bwd_weight(h[l - 1], ab, self.gW[l - 1], 2, 1, accumulate)
            if l > 1 or input_grad:
                hb = ops.buf(f"{self.name}/{tag}/hb{l - 1}", h[l - 1].shape)
                if l == 2:
                    ops.conv_bwd_data(ab, self.W[1], None, hb, 2, 1, act="lrelu", mask=h[1])
                else:
                    ops.conv_bwd_data(ab, self.W[l - 1], None, hb, 2, 1)
        return hb if input_grad else None

    def penalty_sweeps(self, c, weight, inv_global_batch, loss, tag="gp"):
        """sweeps 2 and 3 on the cache `c` of the interpolates.

        Adds weight*mean((||g||-1)^2) to loss[0], WRITES (not accumulates) the penalty's
        first-order-graph contributions into every parameter gradient, and returns
        (dbar [n]: cotangent on the logits, addends for sweep 4)."""
        ops, n = self.ops, c["n"]
        a, h, st, d = c["a"], c["h"], c["st"], c["d"]
        nm = self.name
        # ---- sweep 2: g = d(sigmoid(d)+d)/dxhat -------------------------------------------------
        dd = ops.buf(f"{nm}/{tag}/dd", (n,))
        ops.gp_seed(d, dd)
        dl_h = [None] * 5     # cotangent arriving at h_l (pre activation-mask)
        dl_a = [None] * 5
        dl_h[4] = ops.buf(f"{nm}/{tag}/dl_h4", h[4].shape)
        ops.rowdot_bwd_input(dd, self.w5, dl_h[4])
        for l in range(4, 0, -1):
            if l == 1:
                dl_a[1] = dl_h[1]            # act'(a_1) was applied by the layer-2 input gradient's epilogue
            else:
                dl_a[l] = ops.buf(f"{nm}/{tag}/dl_a{l}", h[l].shape)
                ops.instnorm_bwd(a[l], st[l], dl_h[l], None, dl_a[l], "lrelu")
            dl_h[l - 1] = ops.buf(f"{nm}/{tag}/dl_h{l - 1}", h[l - 1].shape)
            if l == 2:
                ops.conv_bwd_data(dl_a[2], self.W[1], None, dl_h[1], 2, 1, act="lrelu", mask=h[1])
            else:
                ops.conv_bwd_data(dl_a[l], self.W[l - 1], None, dl_h[l - 1], 2, 1)
        g = dl_h[0]
        gbar = ops.buf(f"{nm}/{tag}/gbar", g.shape)
        norms = ops.buf(f"{nm}/{tag}/norms", (n,))
        ops.gp_penalty(g, gbar, norms, loss, weight, inv_global_batch)
        # ---- sweep 3: cotangents of the deltas ----------------------------------------------------
        addends = {"n": n}
        db_h = gbar           # cotangent of dl_h[l-1]
        for l in range(1, 5):
            # dl_h[l-1] = dgrad(dl_a[l], W_l)  (bilinear)
            ops.conv_bwd_weight(db_h, dl_a[l], self.gW[l - 1], 2, 1, False)
            if l == 1:
                nxt = ops.buf(f"{nm}/{tag}/db_h{l}", h[l].shape)
                ops.conv_fwd(db_h, self.W[0], None, nxt, 2, 1, act="lrelu", mask=h[1])
                db_h = nxt
            else:
                db_a = ops.buf(f"{nm}/{tag}/db_a{l}", h[l].shape)
                ops.conv_fwd(db_h, self.W[l - 1], None, db_a, 2, 1)
                db_h = ops.buf(f"{nm}/{tag}/db_h{l}", h[l].shape)
                addends[l] = ops.buf(f"{nm}/{tag}/ax{l}", h[l].shape)
                ops.instnorm_bwd2(a[l], st[l], dl_h[l], db_a, db_h, addends[l], "lrelu")
        # dl_h[4] = dd (x) w5
        ops.rowdot_bwd_weight(dd, db_h, self.gw5, None, False)
        ddbar = ops.buf(f"{nm}/{tag}/ddbar", (n,))
        ops.rowdot_fwd(db_h, self.w5, None, ddbar)
        ops.fill(self.gb5, 0.0)
        return ddbar, addends, norms
