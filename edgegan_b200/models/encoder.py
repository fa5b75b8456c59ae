"""Encoder E (reference edgegan/models/encoder.py:8-25,54-84; nn.residual conv.py:70-85) with its
explicit backward.

conv 4x4 s2 SAME + bias + relu -> per block { residual( reflect-pad 3x3 conv + b, IN, relu,
reflect-pad 3x3 conv + b, IN ; 1x1 shortcut conv + b ; add ; relu ) ; avg-pool 2x2 } with
128/256/512/512 filters -> relu -> 8x8 SAME avg-pool (global mean of the <=8x8 map) -> flatten ->
mlp -> mu, mlp -> log_sigma ; z = mu + eps * exp(log_sigma) with ONE scalar eps.

tf.pad(REFLECT) has no TMA equivalent, so the padded activation is materialised by a small
HBM-bound kernel and the 3x3 convs run as VALID convolutions on it.
"""
from __future__ import annotations

from ..variables import encoder_blocks


class Encoder(object):
    def __init__(self, name, is_train=True, norm="instance", activation="relu", image_size=64,
                 latent_dim=100, use_resnet=True, *, ops=None, store=None):
        if not use_resnet:
            raise NotImplementedError("if_resnet_e=False is outside the hot path (SURVEY.md 2.1)")
        if norm != "instance" or activation != "relu":
            raise NotImplementedError("only E_norm='instance' with relu is implemented")
        self.name, self.ops, self.store = name, ops, store
        self._image_size, self._latent_dim = image_size, latent_dim
        self.var_list = store.names()
        self.blocks = encoder_blocks(image_size)
        self.cache = None

    def _v(self, s):
        return self.store.var[f"{self.name}/{s}"]

    def _g(self, s):
        return self.store.g[f"{self.name}/{s}"]

    def __call__(self, input, eps=0.0):
        return self.forward(input, eps)

    def forward(self, x, eps=0.0, tag="fwd"):
        """x [n,S,S,3] -> (z, mu, log_sigma) each [n, latent_dim]."""
        ops, nm = self.ops, self.name
        n, S = x.shape[0], x.shape[1]
        Wd = x.shape[2]
        B = lambda key, shape: ops.buf(f"{nm}/{tag}/{key}", shape)
        a0 = B("a0", (n, S // 2, Wd // 2, 64))
        ops.conv_fwd(x, self._v("e_resnet_64_0/conv2d/w"), self._v("e_resnet_64_0/conv2d/b"), a0, 2, 1)
        h = B("h0", a0.shape)
        ops.act_fwd(a0, h, "relu")
        blocks = []
        for blk, nf in self.blocks:
            _, hh, ww, cin = h.shape
            hp = B(f"{blk}/hp", (n, hh + 2, ww + 2, cin))
            ops.reflect_pad_fwd(h, hp, 1)
            o1 = B(f"{blk}/o1", (n, hh, ww, nf))
            ops.conv_fwd(hp, self._v(f"{blk}/res1/conv2d/w"), self._v(f"{blk}/res1/conv2d/b"), o1, 1, 0)
            st1, r1 = B(f"{blk}/st1", (n, nf, 2)), B(f"{blk}/r1", o1.shape)
            ops.instnorm_fwd(o1, r1, st1, "relu")
            r1p = B(f"{blk}/r1p", (n, hh + 2, ww + 2, nf))
            ops.reflect_pad_fwd(r1, r1p, 1)
            o2 = B(f"{blk}/o2", o1.shape)
            ops.conv_fwd(r1p, self._v(f"{blk}/res2/conv2d/w"), self._v(f"{blk}/res2/conv2d/b"), o2, 1, 0)
            st2, n2 = B(f"{blk}/st2", (n, nf, 2)), B(f"{blk}/n2", o1.shape)
            ops.instnorm_fwd(o2, n2, st2, "none")
            sc = B(f"{blk}/sc", o1.shape)
            ops.conv_fwd(h, self._v(f"{blk}/shortcut/conv2d/w"), self._v(f"{blk}/shortcut/conv2d/b"), sc, 1, 0)
            out = B(f"{blk}/out", (n, hh // 2, ww // 2, nf))
            ops.addrelu_pool2_fwd(sc, n2, out)
            blocks.append({"blk": blk, "h": h, "hp": hp, "o1": o1, "st1": st1, "r1p": r1p, "o2": o2,
                           "st2": st2, "n2": n2, "sc": sc})
            h = out
        C = h.shape[3]
        f = B("feat", (n, C))
        ops.relu_globalmean_fwd(h, f)
        Z = self._latent_dim
        mu, ls, z = B("mu", (n, Z)), B("ls", (n, Z)), B("z", (n, Z))
        ops.conv_fwd(f.view(n, 1, 1, C), self._v("FC8_mu/w").view(1, 1, C, Z), self._v("FC8_mu/b"), mu.view(n, 1, 1, Z), 1, 0)
        ops.conv_fwd(f.view(n, 1, 1, C), self._v("FC8_sigma/w").view(1, 1, C, Z), self._v("FC8_sigma/b"), ls.view(n, 1, 1, Z), 1, 0)
        ops.reparam_fwd(mu, ls, eps, z)
        self.cache = {"x": x, "a0": a0, "blocks": blocks, "hlast": h, "f": f, "mu": mu, "ls": ls, "n": n, "eps": eps}
        return z, mu, ls

    def backward(self, gmu, gls, tag="bwd"):
        """gmu, gls: cotangents of mu and log_sigma (the reparameterisation is folded into them)."""
        ops, nm, c = self.ops, self.name, self.cache
        n = c["n"]
        B = lambda key, shape: ops.buf(f"{nm}/{tag}/{key}", shape)
        f, hl = c["f"], c["hlast"]
        C, Z = f.shape[1], self._latent_dim
        f4 = f.view(n, 1, 1, C)
        gf, gf2 = B("gf", (n, C)), B("gf2", (n, C))
        for key, g, dst in (("FC8_mu", gmu, gf), ("FC8_sigma", gls, gf2)):
            w4 = self._v(f"{key}/w").view(1, 1, C, Z)
            ops.conv_bwd_weight(f4, g.view(n, 1, 1, Z), self._g(f"{key}/w").view(1, 1, C, Z), 1, 0, False)
            ops.bias_grad(g, self._g(f"{key}/b"), False)
            ops.conv_bwd_data(g.view(n, 1, 1, Z), w4, None, dst.view(n, 1, 1, C), 1, 0)
        ops.axpby(gf2, gf, 1.0, 1.0)
        gh = B("ghlast", hl.shape)
        ops.relu_globalmean_bwd(hl, gf, gh)
        for b in reversed(c["blocks"]):
            blk = b["blk"]
            g = B(f"{blk}/g", b["sc"].shape)
            ops.addrelu_pool2_bwd(b["sc"], b["n2"], gh, g)
            # shortcut branch
            ops.bias_grad(g, self._g(f"{blk}/shortcut/conv2d/b"), False)
            ops.conv_bwd_weight(b["h"], g, self._g(f"{blk}/shortcut/conv2d/w"), 1, 0, False)
            gh_in = B(f"{blk}/gh_in", b["h"].shape)
            ops.conv_bwd_data(g, self._v(f"{blk}/shortcut/conv2d/w"), None, gh_in, 1, 0)
            # residual branch
            go2 = B(f"{blk}/go2", b["o2"].shape)
            ops.instnorm_bwd(b["o2"], b["st2"], g, None, go2, "none")
            ops.bias_grad(go2, self._g(f"{blk}/res2/conv2d/b"), False)
            ops.conv_bwd_weight(b["r1p"], go2, self._g(f"{blk}/res2/conv2d/w"), 1, 0, False)
            gr1p = B(f"{blk}/gr1p", b["r1p"].shape)
            ops.conv_bwd_data(go2, self._v(f"{blk}/res2/conv2d/w"), None, gr1p, 1, 0)
            gr1 = B(f"{blk}/gr1", b["o1"].shape)
            ops.reflect_pad_bwd(gr1p, gr1, 1)
            go1 = B(f"{blk}/go1", b["o1"].shape)
            ops.instnorm_bwd(b["o1"], b["st1"], gr1, None, go1, "relu")
            ops.bias_grad(go1, self._g(f"{blk}/res1/conv2d/b"), False)
            ops.conv_bwd_weight(b["hp"], go1, self._g(f"{blk}/res1/conv2d/w"), 1, 0, False)
            ghp = B(f"{blk}/ghp", b["hp"].shape)
            ops.conv_bwd_data(go1, self._v(f"{blk}/res1/conv2d/w"), None, ghp, 1, 0)
            gh_res = B(f"{blk}/gh_res", b["h"].shape)
            ops.reflect_pad_bwd(ghp, gh_res, 1)
            ops.axpby(gh_res, gh_in, 1.0, 1.0)
            gh = gh_in
        ga0 = B("ga0", c["a0"].shape)
        ops.act_bwd(c["a0"], gh, ga0, "relu")
        ops.bias_grad(ga0, self._g("e_resnet_64_0/conv2d/b"), False)
        ops.conv_bwd_weight(c["x"], ga0, self._g("e_resnet_64_0/conv2d/w"), 2, 1, False)
