"""Generator (reference edgegan/models/generator.py:10-74) with its explicit backward.

linear(z) -> reshape [B, H/16, W/16, 512] -> BATCH norm (+gamma/beta, always batch statistics: the
reference's `nn.norm(h0, self._norm)` binds 'instance' to `is_train`) -> relu ->
3x deconv_block(5x5, stride 2, bias, instance norm, relu) -> deconv(5x5, stride 2, bias) -> tanh.

conv2d_transpose SAME s2 k5 is the input-gradient of a stride-2 conv with padding 1 before
(SURVEY A2), so the forward uses the conv *bwd_data* kernel and the backward the *fwd* / *bwd_weight*
kernels with the roles of input and output swapped.
"""
from __future__ import annotations


class Generator(object):
    def __init__(self, name, is_train=True, norm="instance", activation="relu", batch_size=64,
                 output_height=64, output_width=64, input_dim=64, output_dim=3, use_resnet=False,
                 *, ops=None, store=None, comm=None):
        if use_resnet:
            raise NotImplementedError("if_resnet_g=True is outside the hot path (SURVEY.md 2.1)")
        if norm != "instance" or activation != "relu":
            raise NotImplementedError("only G_norm='instance' with relu is implemented")
        self.name = name
        self._batch_size = batch_size
        self._output_height, self._output_width = output_height, output_width
        self._input_dim, self._output_dim = input_dim, output_dim
        self.ops, self.store, self.comm = ops, store, comm
        self.var_list = store.names()
        v, g = store.var, store.g
        self.Wl, self.bl = v[f"{name}/g_lin_0/Matrix"], v[f"{name}/g_lin_0/bias"]
        self.gWl, self.gbl = g[f"{name}/g_lin_0/Matrix"], g[f"{name}/g_lin_0/bias"]
        self.gamma, self.beta = v[f"{name}/batch_norm/BatchNorm/gamma"], v[f"{name}/batch_norm/BatchNorm/beta"]
        self.ggamma, self.gbeta = g[f"{name}/batch_norm/BatchNorm/gamma"], g[f"{name}/batch_norm/BatchNorm/beta"]
        self.W = [None] + [v[f"{name}/g_dconv_{i}/deconv2d/w"] for i in range(1, 5)]
        self.b = [None] + [v[f"{name}/g_dconv_{i}/deconv2d/b"] for i in range(1, 5)]
        self.gW = [None] + [g[f"{name}/g_dconv_{i}/deconv2d/w"] for i in range(1, 5)]
        self.gb = [None] + [g[f"{name}/g_dconv_{i}/deconv2d/b"] for i in range(1, 5)]
        self.cache = None

    def _shapes(self, n):
        gf, H, W = self._input_dim, self._output_height, self._output_width
        ch = [gf * 8, gf * 4, gf * 2, gf, self._output_dim]
        return [(n, H >> (4 - l), W >> (4 - l), ch[l]) for l in range(5)]

    def __call__(self, z):
        return self.forward(z)

    def forward(self, z, tag="fwd"):
        """z [n, input_dim(+classes)] -> tanh output [n, H, W, 3]; keeps the cache for backward()."""
        ops, nm, n = self.ops, self.name, z.shape[0]
        shp = self._shapes(n)
        K = z.shape[1]
        C0 = shp[0][3]
        F = shp[0][1] * shp[0][2] * C0
        lin = ops.buf(f"{nm}/{tag}/lin", (n, F))
        ops.conv_fwd(z.view(n, 1, 1, K), self.Wl.view(1, 1, K, F), self.bl, lin.view(n, 1, 1, F), 1, 0)
        sums = ops.buf(f"{nm}/{tag}/bn_sums", (2 * C0,))
        ops.bn_stats(lin.view(-1, C0), sums)
        rows = n * F // C0
        count = rows
        if self.comm is not None and self.comm.world_size > 1:
            self.comm.allreduce(sums)
            count = rows * self.comm.world_size
        a, h, st = [None] * 5, [None] * 5, [None] * 5
        h[0] = ops.buf(f"{nm}/{tag}/h0", shp[0])
        ops.bn_apply(lin.view(-1, C0), sums, count, self.gamma, self.beta, h[0].view(-1, C0), "relu")
        for l in range(1, 5):
            a[l] = ops.buf(f"{nm}/{tag}/a{l}", shp[l])
            ops.conv_bwd_data(h[l - 1], self.W[l], self.b[l], a[l], 2, 1)
            h[l] = ops.buf(f"{nm}/{tag}/h{l}", shp[l])
            if l < 4:
                st[l] = ops.buf(f"{nm}/{tag}/st{l}", (n, shp[l][3], 2))
                ops.instnorm_fwd(a[l], h[l], st[l], "relu")
            else:
                ops.act_fwd(a[l], h[l], "tanh")
        self.cache = {"z": z, "lin": lin, "sums": sums, "count": count, "a": a, "h": h, "st": st, "n": n,
                      "tag": tag, "C0": C0, "F": F}
        return h[4]

    def backward(self, gout, tag="bwd"):
        """gout: cotangent of the tanh output.  Writes every parameter gradient of this generator."""
        ops, nm, c = self.ops, self.name, self.cache
        a, h, st, n = c["a"], c["h"], c["st"], c["n"]
        gh = gout
        for l in range(4, 0, -1):
            ga = ops.buf(f"{nm}/{tag}/ga{l}", a[l].shape)
            if l == 4:
                ops.act_bwd(a[4], gh, ga, "tanh")
            else:
                ops.instnorm_bwd(a[l], st[l], gh, None, ga, "relu")
            ops.bias_grad(ga, self.gb[l], False)
            ops.conv_bwd_weight(ga, h[l - 1], self.gW[l], 2, 1, False)
            gh = ops.buf(f"{nm}/{tag}/gh{l - 1}", h[l - 1].shape)
            ops.conv_fwd(ga, self.W[l], None, gh, 2, 1)
        C0, F, K = c["C0"], c["F"], c["z"].shape[1]
        lin2 = c["lin"].view(-1, C0)
        red = ops.buf(f"{nm}/{tag}/bn_red", (2 * C0,))
        ops.bn_bwd_reduce(lin2, c["sums"], c["count"], self.gamma, self.beta, gh.view(-1, C0), red, "relu")
        # local sums are this rank's share of d(beta), d(gamma); the gradient all-reduce adds the ranks up
        ops.copy(red[:C0], self.gbeta)
        ops.copy(red[C0:], self.ggamma)
        if self.comm is not None and self.comm.world_size > 1:
            self.comm.allreduce(red)
        glin = ops.buf(f"{nm}/{tag}/glin", (n, F))
        ops.bn_bwd_apply(lin2, c["sums"], c["count"], self.gamma, self.beta, gh.view(-1, C0), red,
                         glin.view(-1, C0), "relu")
        ops.conv_bwd_weight(c["z"].view(n, 1, 1, K), glin.view(n, 1, 1, F), self.gWl.view(1, 1, K, F), 1, 0, False)
        ops.bias_grad(glin, self.gbl, False)
