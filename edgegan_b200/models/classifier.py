"""Multi-class classifier D2 (reference edgegan/models/classifier.py:7-119; MRU block nn/modules/conv.py:133-243,
298-357; conv2d2 conv.py:246-295; fully_connected linear.py:34-76; spectral norm normalization.py:38-76) with
explicit backward passes.

Structure (hidden, filter) = (8,128), (128,256), (256,512), (512,768), one MRU unit per pyramid level:
    h0   = prelu(SNconv7x7(x, 3->8) + b)
    unit: a_in = prelu(ht);  full = concat(a_in, inp)                  inp = image mean-pooled to ht's resolution
          rg   = minmax_HW(lrelu(SNconv3x3(full -> hd) + b))           (update gate, bias init 0.5)
          img  = SNconv3x3(inp -> hd) + b
          hin  = prelu(ht + rg * img)
          hn   = SNconv3x3(prelu(SNconv3x3(hin -> fd) + b) -> fd) + b
          ht'  = mean_pool(SNconv1x1(ht -> fd) + b + hn)
    logits = SNfc(mean_HW(prelu(ht_4))) + b
The reference runs this network in NCHW (classifier.py:13); that is an API-only layout: EdgeGAN transposes its
NHWC tensors right before the call (edgegan.py:28-29,231), so the NHWC tensors are consumed directly here and
`__call__` accepts the reference's NCHW view.  The 1x1 "disc" head (classifier.py:107-109) is built but never
used by any loss; its variables exist (and stay at their initial values) for variable-list compatibility.

Every weight is spectrally normalised with ONE power iteration from a frozen `u` that the reference never
updates; the gradient flows through sigma (eg_spectral_norm_bwd).
"""
from __future__ import annotations

import math

import numpy as np

from ..variables import ParamStore, VarSpec

UNITS = ((8, 128), (128, 256), (256, 512), (512, 768))


def _conv_specs(scope, k, ci, co, bias_init=0.0, prelu=False):
    v = [VarSpec(f"{scope}/weights", (k, k, ci, co), "normal"),
         VarSpec(f"{scope}/biases", (1, co, 1, 1), "const", value=bias_init)]
    if prelu:
        v.append(VarSpec(f"{scope}/prelu/param", (), "const", value=0.2))
    return v


def classifier_specs(name, num_classes=14, c_dim=3):
    """trainables that receive gradients, in the reference's creation order"""
    v = _conv_specs(f"{name}/Conv", 7, c_dim, 8, prelu=True)
    for t, (hd, fd) in enumerate(UNITS, start=1):
        p = f"{name}/mru_conv_unit_t_{t}_layer_0"
        v.append(VarSpec(f"{p}/norm_activation_in/prelu/param", (), "const", value=0.2))
        v += _conv_specs(f"{p}/update_gate", 3, hd + c_dim, hd, bias_init=0.5)
        v += _conv_specs(f"{p}/Conv", 3, c_dim, hd)
        v.append(VarSpec(f"{p}/norm_activation_merge_1/prelu/param", (), "const", value=0.2))
        v += _conv_specs(f"{p}/Conv_1", 3, hd, fd, prelu=True)
        v += _conv_specs(f"{p}/Conv_2", 3, fd, fd)
        v += _conv_specs(f"{p}/Conv_3", 1, hd, fd)
    v.append(VarSpec(f"{name}/mru_conv_unit_last_norm/prelu/param", (), "const", value=0.2))
    lim = math.sqrt(6.0 / (768 + num_classes))
    v.append(VarSpec(f"{name}/fully_connected/weights", (768, num_classes), "uniform", std=lim))
    v.append(VarSpec(f"{name}/fully_connected/biases", (num_classes,), "zeros"))
    return v


def classifier_aux_specs(name, num_classes=14, c_dim=3):
    """variables no gradient reaches: the unused disc head and the frozen spectral-norm vectors u [1, Cout]"""
    v = _conv_specs(f"{name}/Conv_1", 1, 768, 1)
    for s in classifier_specs(name, num_classes, c_dim) + v[:1]:
        if s.name.endswith("/weights"):
            v.append(VarSpec(s.name[:-len("weights")] + "u", (1, s.shape[-1]), "trunc_normal", std=1.0))
    return v


class Classifier(object):
    def __init__(self, name, SPECTRAL_NORM_UPDATE_OPS="spectral_norm_update_ops", *, ops=None, store=None, rs=None,
                 num_classes=14, c_dim=3):
        self.name, self.ops, self.store = name, ops, store
        self.SPECTRAL_NORM_UPDATE_OPS = SPECTRAL_NORM_UPDATE_OPS
        self.num_classes, self.c_dim = num_classes, c_dim
        self.aux = ParamStore(ops, classifier_aux_specs(name, num_classes, c_dim), rs or np.random.RandomState(0),
                              conv_filter_set=False)
        self.var_list = store.names() + [n for n in self.aux.names() if not n.endswith("/u")]
        self.layers = [s.name[:-len("/weights")] for s in store.specs if s.name.endswith("/weights")]
        self._wbar_valid = False
        self.cache = None

    # ---- variables -----------------------------------------------------------------------------------
    def load_u(self, values):
        sub = {k: v for k, v in values.items() if k in self.aux.offsets}
        if sub:
            self.aux.load(sub, strict=False)
        self._wbar_valid = False

    def invalidate(self):
        self._wbar_valid = False

    def _v(self, s):
        return self.store.var[f"{self.name}/{s}"]

    def _g(self, s):
        return self.store.g[f"{self.name}/{s}"]

    def _normalise_weights(self):
        """Wbar = W / sigma(W) for every layer (normalization.py:38-76); cached until the weights change."""
        if self._wbar_valid:
            return
        ops = self.ops
        self.wbar, self.ws, self.gbar, items = {}, {}, {}, []
        for scope in self.layers:
            W = self.store.var[scope + "/weights"]
            Cn = W.shape[-1]
            K = W.numel() // Cn
            wb = ops.buf(scope + "/wbar", W.shape)
            ws = ops.buf(scope + "/sn_ws", (ops.sn_ws_floats(K, Cn),))
            self.wbar[scope], self.ws[scope] = wb, ws
            it = dict(W=W, u=self.aux.var[scope + "/u"], Wbar=wb, ws=ws, gW=self.store.g[scope + "/weights"])
            if scope.endswith("/update_gate"):
                # conv(concat(a, inp)) = conv(a, W[:, :, :hd]) + conv(inp, W[:, :, hd:]): two contiguous filter copies, so
                # the hd-channel part runs on the tensor cores (hd + 3 input channels would not) and no concat is built;
                # dL/dWbar likewise arrives as two filter gradients
                k, _, cin, co = W.shape
                hd = cin - self.c_dim
                wa, wi = ops.buf(scope + "/wbar_a", (k, k, hd, co)), ops.buf(scope + "/wbar_i", (k, k, self.c_dim, co))
                ga, gi = ops.buf(scope + "/gbar_a", (k, k, hd, co)), ops.buf(scope + "/gbar_i", (k, k, self.c_dim, co))
                self.wbar[scope + "#a"], self.wbar[scope + "#i"] = wa, wi
                self.gbar[scope + "#a"], self.gbar[scope + "#i"] = ga, gi
                it.update(Wa=wa, Wi=wi, Ga=ga, Gi=gi, hd=hd)
            else:
                self.gbar[scope] = it["G"] = ops.buf(scope + "/gbar", W.shape)      # dL/dWbar of this layer
            items.append(it)
        # one descriptor table for the whole network: 3 launches normalise all filters (4 more map all dL/dWbar to dL/dW)
        key = tuple(t.data_ptr() for it in items for t in it.values() if hasattr(t, "data_ptr"))
        if getattr(self, "_sn_key", None) != key:
            if getattr(self, "_sn", None) is not None:
                self._sn.close()
            self._sn, self._sn_key = ops.spectral_norm_set(items), key
        self._sn.fwd()
        # prepared copies (tensor-core operand layouts) of the normalised filters: one kernel for the whole network
        algo = getattr(ops, "default_algo", None)
        tc = [w for w in self.wbar.values() if ops.in_filter_set(w.shape)]
        ids = tuple(w.data_ptr() for w in tc)
        if getattr(self, "_fset_key", None) != (algo, ids):
            if getattr(self, "_fset", None) is not None:
                self._fset.close()
            self._fset, self._fset_key = ops.filter_set(tc), (algo, ids)
        if self._fset is not None:
            self._fset.prepare()
        self._wbar_valid = True

    # ---- forward -------------------------------------------------------------------------------------
    def _conv(self, scope, x, y, k):
        full = f"{self.name}/{scope}"
        self.ops.conv_fwd(x, self.wbar[full], self.store.var[full + "/biases"].view(-1), y, 1, (k - 1) // 2)

    def forward(self, x, tag="fwd"):
        """x [n,S,S,3] NHWC -> logits [n, num_classes]"""
        ops, nm = self.ops, self.name
        self._normalise_weights()
        n, S = x.shape[0], x.shape[1]
        B = lambda key, shape: ops.buf(f"{nm}/{tag}/{key}", shape)
        pyr = [x]
        for l in range(1, 4):
            p = B(f"pyr{l}", (n, S >> l, S >> l, self.c_dim))
            ops.add_pool2_fwd(pyr[-1], None, p)
            pyr.append(p)
        c0 = B("c0", (n, S, S, 8))
        self._conv("Conv", x, c0, 7)
        ht = B("h0", c0.shape)
        # h0 = prelu(c0) and the first unit's a_in = prelu(h0) in one pass; later units get their a_in from the pooling
        # kernel that produces their ht
        a_in = B("u1/a_in", c0.shape)
        ops.prelu_fwd2(c0, self._v("Conv/prelu/param"), ht, self._v("mru_conv_unit_t_1_layer_0/norm_activation_in/prelu/param"), a_in)
        units = []
        for t, (hd, fd) in enumerate(UNITS, start=1):
            p = f"mru_conv_unit_t_{t}_layer_0"
            inp = pyr[t - 1]
            H = inp.shape[1]
            U = lambda key, c: B(f"u{t}/{key}", (n, H, H, c))
            # update gate: conv(concat(a_in, inp)) as two convs; rgl = lrelu(sum) is written over cg by the fused kernel
            rgl, cg_i = U("rgl", hd), U("cg_i", hd)
            gate = f"{nm}/{p}/update_gate"
            ops.conv_fwd(a_in, self.wbar[gate + "#a"], self.store.var[gate + "/biases"].view(-1), rgl, 1, 1)
            ops.conv_fwd(inp, self.wbar[gate + "#i"], None, cg_i, 1, 1)
            img = U("img", hd)
            self._conv(f"{p}/Conv", inp, img, 3)
            mm = B(f"u{t}/mm", (n, hd, 4))
            plus, hin = U("plus", hd), U("hin", hd)
            ops.mru_gate_fwd(rgl, cg_i, ht, img, self._v(f"{p}/norm_activation_merge_1/prelu/param"), mm, plus, hin)
            c1, hn1, hn2, ho = U("c1", fd), U("hn1", fd), U("hn2", fd), U("ho", fd)
            self._conv(f"{p}/Conv_1", hin, c1, 3)
            ops.prelu_fwd(c1, self._v(f"{p}/Conv_1/prelu/param"), hn1)
            self._conv(f"{p}/Conv_2", hn1, hn2, 3)
            self._conv(f"{p}/Conv_3", ht, ho, 1)
            out = B(f"u{t}/out", (n, H // 2, H // 2, fd))
            units.append(dict(p=p, hd=hd, fd=fd, ht=ht, inp=inp, a_in=a_in, rgl=rgl, mm=mm, img=img, plus=plus, hin=hin,
                              c1=c1, hn1=hn1))
            if t < len(UNITS):
                a_in = B(f"u{t + 1}/a_in", out.shape)
                ops.add_pool2_fwd(ho, hn2, out, self._v(f"mru_conv_unit_t_{t + 1}_layer_0/norm_activation_in/prelu/param"), a_in)
            else:
                ops.add_pool2_fwd(ho, hn2, out)
            ht = out
        hl = B("hlast", ht.shape)
        ops.prelu_fwd(ht, self._v("mru_conv_unit_last_norm/prelu/param"), hl)
        feat = B("feat", (n, ht.shape[3]))
        ops.globalmean_fwd(hl, feat)
        logits = B("logits", (n, self.num_classes))
        C = feat.shape[1]
        ops.conv_fwd(feat.view(n, 1, 1, C), self.wbar[f"{nm}/fully_connected"].view(1, 1, C, self.num_classes),
                     self._v("fully_connected/biases"), logits.view(n, 1, 1, self.num_classes), 1, 0)
        self.cache = dict(x=x, pyr=pyr, c0=c0, units=units, ht_last=ht, hl=hl, feat=feat, logits=logits, n=n, tag=tag)
        return logits

    def __call__(self, x, num_classes=None, labels=None, reuse=False, data_format="NCHW"):
        """Reference signature (classifier.py:12): x NCHW view of an NHWC buffer -> (disc, prob, logits).
        `disc` is the 1x1 'discriminator end' head (classifier.py:104-107, [n, h, w, 1] here); nothing in the training
        step reads it, so only this call computes it."""
        assert data_format == "NCHW"
        xn = x.permute(0, 2, 3, 1)
        if not xn.is_contiguous():
            raise ValueError("pass the NCHW *view* of an NHWC buffer (edgegan.py:28-29 transposes right before the call)")
        ops, nm = self.ops, self.name
        logits = self.forward(xn, "call")
        prob = ops.buf(f"{nm}/call/prob", logits.shape)
        ops.act_fwd(logits, prob, "sigmoid")
        hl = self.cache["hl"]
        W, u = self.aux.var[f"{nm}/Conv_1/weights"], self.aux.var[f"{nm}/Conv_1/u"]
        wb = ops.buf(f"{nm}/Conv_1/wbar", W.shape)
        ws = ops.buf(f"{nm}/Conv_1/sn_ws", (ops.sn_ws_floats(W.numel() // W.shape[-1], W.shape[-1]),))
        ops.spectral_norm_fwd(W, u, wb, ws)
        disc = ops.buf(f"{nm}/call/disc", (hl.shape[0], hl.shape[1], hl.shape[2], 1))
        ops.conv_fwd(hl, wb, self.aux.var[f"{nm}/Conv_1/biases"].view(-1), disc, 1, 0)
        return disc, prob, logits

    # ---- backward ------------------------------------------------------------------------------------
    def _wgrad(self, scope, x, dy, k, tag):
        """dL/dWbar by the conv filter-gradient kernel; backward() maps all of them through the spectral norm at its end"""
        ops, full = self.ops, f"{self.name}/{scope}"
        ops.conv_bwd_weight(x, dy, self.gbar[full], 1, (k - 1) // 2, False)
        ops.bias_grad(dy, self.store.g[full + "/biases"].view(-1), False)

    def backward(self, glogits, param_grads, input_grad, tag="bwd"):
        ops, nm, c = self.ops, self.name, self.cache
        n = c["n"]
        B = lambda key, shape: ops.buf(f"{nm}/{tag}/{key}", shape)
        feat, hl, ht = c["feat"], c["hl"], c["ht_last"]
        C, K = feat.shape[1], self.num_classes
        fc = f"{nm}/fully_connected"
        if param_grads:
            ops.conv_bwd_weight(feat.view(n, 1, 1, C), glogits.view(n, 1, 1, K), self.gbar[fc].view(1, 1, C, K), 1, 0, False)
            ops.bias_grad(glogits, self._g("fully_connected/biases"), False)
        gfeat = B("gfeat", feat.shape)
        ops.conv_bwd_data(glogits.view(n, 1, 1, K), self.wbar[fc].view(1, 1, C, K), None, gfeat.view(n, 1, 1, C), 1, 0)
        ghl = B("ghl", hl.shape)
        ops.globalmean_bwd(gfeat, ghl)
        g_out = B("g_ht4", ht.shape)
        ops.prelu_bwd(ht, self._v("mru_conv_unit_last_norm/prelu/param"), ghl, g_out,
                      self._g("mru_conv_unit_last_norm/prelu/param") if param_grads else None)
        g_pyr = [None] * 4
        for t in range(4, 0, -1):
            u = c["units"][t - 1]
            p, hd, fd, htu, inp = u["p"], u["hd"], u["fd"], u["ht"], u["inp"]
            H = inp.shape[1]
            U = lambda key, ch: B(f"u{t}/{key}", (n, H, H, ch))
            g_sum = U("g_sum", fd)
            ops.pool2_bwd(g_out, g_sum, False)
            wb = lambda s: self.wbar[f"{nm}/{p}/{s}"]
            # ho = Conv_3(ht), hn2 = Conv_2(hn1)
            g_ht, g_hn1 = U("g_ht", hd), U("g_hn1", fd)
            if param_grads:
                self._wgrad(f"{p}/Conv_3", htu, g_sum, 1, tag)
                self._wgrad(f"{p}/Conv_2", u["hn1"], g_sum, 3, tag)
            ops.conv_bwd_data(g_sum, wb("Conv_3"), None, g_ht, 1, 0)
            ops.conv_bwd_data(g_sum, wb("Conv_2"), None, g_hn1, 1, 1)
            g_c1 = U("g_c1", fd)
            ops.prelu_bwd(u["c1"], self._v(f"{p}/Conv_1/prelu/param"), g_hn1, g_c1,
                          self._g(f"{p}/Conv_1/prelu/param") if param_grads else None)
            if param_grads:
                self._wgrad(f"{p}/Conv_1", u["hin"], g_c1, 3, tag)
            g_hin = U("g_hin", hd)
            ops.conv_bwd_data(g_c1, wb("Conv_1"), None, g_hin, 1, 1)
            # plus = ht + rg * img, hin = prelu(plus), rg = minmax(rgl), rgl = lrelu(cg): one fused kernel
            g_img, g_cg = U("g_img", hd), U("g_cg", hd)
            ops.mru_gate_bwd(u["plus"], g_hin, u["img"], u["rgl"], u["mm"], self._v(f"{p}/norm_activation_merge_1/prelu/param"),
                             g_ht, g_img, g_cg, self._g(f"{p}/norm_activation_merge_1/prelu/param") if param_grads else None)
            if param_grads:
                self._wgrad(f"{p}/Conv", inp, g_img, 3, tag)
            gate = f"{nm}/{p}/update_gate"
            if param_grads:
                # dL/dWbar of the two filter parts; the spectral-norm backward reads them as one [k, k, hd+3, hd] tensor
                ops.conv_bwd_weight(u["a_in"], g_cg, self.gbar[gate + "#a"], 1, 1, False)
                ops.conv_bwd_weight(inp, g_cg, self.gbar[gate + "#i"], 1, 1, False)
                ops.bias_grad(g_cg, self.store.g[gate + "/biases"].view(-1), False)
            g_ain = U("g_ain", hd)
            ops.conv_bwd_data(g_cg, self.wbar[gate + "#a"], None, g_ain, 1, 1)
            ops.prelu_bwd(htu, self._v(f"{p}/norm_activation_in/prelu/param"), g_ain, g_ht,
                          self._g(f"{p}/norm_activation_in/prelu/param") if param_grads else None, accumulate_gx=True)
            if input_grad:
                gi = B(f"g_pyr{t - 1}", inp.shape)
                ops.conv_bwd_data(g_img, wb("Conv"), None, gi, 1, 1)
                gi2 = U("g_inp2", self.c_dim)
                ops.conv_bwd_data(g_cg, self.wbar[gate + "#i"], None, gi2, 1, 1)
                ops.axpby(gi2, gi, 1.0, 1.0)
                g_pyr[t - 1] = gi
            g_out = g_ht
        # h0 = prelu(Conv(x))
        g_c0 = B("g_c0", c["c0"].shape)
        ops.prelu_bwd(c["c0"], self._v("Conv/prelu/param"), g_out, g_c0, self._g("Conv/prelu/param") if param_grads else None)
        if param_grads:
            self._wgrad("Conv", c["x"], g_c0, 7, tag)
            self._sn.bwd()                 # every dL/dWbar is in place: dL/dW of all 22 filters (through sigma, v and u')
        if not input_grad:
            return None
        for l in range(3, 0, -1):                      # pyramid: pyr[l] = mean_pool(pyr[l-1])
            ops.pool2_bwd(g_pyr[l], g_pyr[l - 1], True)
        gx = g_pyr[0]
        gx7 = B("g_x7", c["x"].shape)
        ops.conv_bwd_data(g_c0, self.wbar[f"{nm}/Conv"], None, gx7, 1, 3)
        ops.axpby(gx7, gx, 1.0, 1.0)
        return gx

    # ---- the two uses in the step ----------------------------------------------------------------------
    def real_loss_step(self, real_pic, z, inv_global_batch, loss):
        """run 4 (d_optim2): focal loss of the REAL image (functional.py:8-11), gradients of every D2 trainable."""
        ops = self.ops
        logits = self.forward(real_pic, "real")
        gl = ops.buf(f"{self.name}/real/glogits", logits.shape)
        ops.fill(loss, 0.0)
        ops.softmax_ce_bwd(logits, z, z.shape[1] - 1, True, 1.0, inv_global_batch, gl, loss)
        self.backward(gl, param_grads=True, input_grad=False, tag="real_bwd")
        self._wbar_valid = False          # the caller applies RMSProp next

    def fake_loss_input_grad(self, fake, z, weight, inv_global_batch, loss):
        """runs 5/7: weight * mean CE(C(G2(z)), class) (functional.py:13-15) and its gradient w.r.t. the image."""
        ops = self.ops
        logits = self.forward(fake, "fake")
        gl = ops.buf(f"{self.name}/fake/glogits", logits.shape)
        ops.fill(loss, 0.0)
        ops.softmax_ce_bwd(logits, z, z.shape[1] - 1, False, weight, inv_global_batch, gl, loss)
        return self.backward(gl, param_grads=False, input_grad=True, tag="fake_bwd")
