"""EdgeGAN step orchestration (reference edgegan/models/edgegan.py:76-130,132-342,492-517).

`update_model(images, z)` performs the reference's sequential RMSProp runs in `construct_optimizers`
order (edgegan.py:109-124): d_optim, d_optim_patch2, d_optim_patch3, [d_optim2], g_optim_u, e_optim,
g_optim_b -- each run seeing the weights left by the previous one.  What TensorFlow did implicitly
(re-executing the pruned forward graph per sess.run, tf.gradients, minimize) is explicit here:

  * the generators are only re-run when their weights changed (runs 1-5 share one G forward; the
    reference recomputes the identical values in every sess.run);
  * each critic sees real, fake and interpolated samples as ONE 3B batch (identical arithmetic,
    fewer and larger kernels), and the penalty's double backward is hand-derived
    (models/discriminator.py);
  * the random inputs the reference samples inside the graph -- alpha ~ U[0,1) per critic
    (edgegan.py:32-35) and the scalar encoder noise (encoder.py:78) -- are explicit arguments.

Data parallelism: the batch is sharded over ranks, all loss means divide by the GLOBAL batch, and the
flat gradient buffer of the network being optimised is all-reduced (sum) before its RMSProp update.
"""
from __future__ import annotations

import numpy as np

from ..config import Flags
from ..variables import (ParamStore, discriminator_specs, encoder_specs, generator_specs)
from .discriminator import Discriminator
from .encoder import Encoder
from .generator import Generator

RUN_NAMES = ("d_optim", "d_optim_patch2", "d_optim_patch3", "d_optim2", "g_optim_u", "e_optim", "g_optim_b")
LOSS_SLOTS = {"joint_dis_dloss": 0, "image_dis_dloss": 1, "edge_dis_dloss": 2, "loss_d_ac": 3,
              "edge_gloss": 4, "image_gloss": 5, "zl_loss": 6, "edge_gloss_b": 7, "image_gloss_b": 8,
              "loss_g_ac": 9}


class LocalComm:
    """Single-process stand-in for the NCCL communicator."""
    world_size, rank = 1, 0

    def allreduce(self, t):
        return t


class _GraphStep:
    """update_model as a CUDA-graph replay for the training loop (see EdgeGAN.train)."""

    def __init__(self, model, eager_steps=2):
        self.m, self.eager_steps, self.n, self.graph, self.shape = model, eager_steps, 0, None, None
        self.stream = None

    def __call__(self, images, z):
        """runs on a dedicated stream (the one the graph is captured on, so that the library's per-stream scratch exists
        before the capture), ordered after / before the caller's current stream"""
        import torch
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self._step(images, z)
        cur.wait_stream(self.stream)

    def _step(self, images, z):
        import torch
        m, ops = self.m, self.m.ops
        rs = getattr(m, "_rs", None) or np.random.RandomState(m.seed + 1)
        m._rs = rs
        nb = images.shape[0]
        alpha = rs.uniform(0, 1, (3, nb)).astype(np.float32)
        eps = float(rs.normal())
        shape = (tuple(images.shape), tuple(z.shape))
        if self.shape is not None and shape != self.shape:      # e.g. another batch size: back to eager launches
            self.graph, self.n, self.shape = None, 0, None
        self.n += 1
        if self.graph is None and self.n <= self.eager_steps:
            m.update_model(images, z, ops.from_numpy(alpha), eps)
            return
        if self.graph is None:
            self.shape = shape
            self.s_img, self.s_z = torch.empty_like(images), torch.empty_like(z)
            self.s_alpha = ops.empty((3, nb))
            self.s_eps = ops.empty((1,))
            self.h_alpha = torch.empty((3, nb), dtype=torch.float32).pin_memory()
            self.h_eps = torch.empty((1,), dtype=torch.float32).pin_memory()
            self.graph = m.capture_step(self.s_img, self.s_z, self.s_alpha, self.s_eps, warmup=0, stream=self.stream)
        self.s_img.copy_(images, non_blocking=True)
        self.s_z.copy_(z, non_blocking=True)
        self.h_alpha.numpy()[...] = alpha
        self.h_eps[0] = eps
        self.s_alpha.copy_(self.h_alpha, non_blocking=True)
        self.s_eps.copy_(self.h_eps, non_blocking=True)
        self.graph.replay()
        torch.cuda.current_stream().synchronize()               # the pinned alpha / eps staging is reused next iteration


class EdgeGAN(object):
    def __init__(self, sess=None, config: Flags = None, dataset=None, z_dim=100, gf_dim=64, df_dim=64,
                 gfc_dim=1024, dfc_dim=1024, c_dim=3, *, ops=None, comm=None, seed=0):
        self.sess = sess                 # kept for signature compatibility; unused
        self.config = config.validate()
        self.dataset = dataset
        self.z_dim, self.gf_dim, self.df_dim, self.c_dim = z_dim, gf_dim, df_dim, c_dim
        if ops is None:
            from ..ops import DeviceOps
            ops = DeviceOps()
        self.ops = ops
        self.comm = comm or LocalComm()
        self.seed = seed
        self.stores = {}
        self._built = None
        self.run_hook = None             # test hook: called as run_hook(run_name, self) before each RMSProp apply

    # ---- construction -----------------------------------------------------------------------------
    @property
    def multiclass(self):
        return bool(self.config.multiclasses)

    def _store(self, name, specs, rs, conv_filter_set=True):
        st = ParamStore(self.ops, specs, rs, conv_filter_set)
        self.stores[name] = st
        return st

    def build_networks(self, train=True):
        """edgegan.py:132-177 (train) / 519-549 (test: E, G1, G2 only)."""
        cfg, ops = self.config, self.ops
        rs = np.random.RandomState(self.seed)
        H, Wp = cfg.output_height, cfg.output_width
        half = Wp // 2
        g_in = self.z_dim + (cfg.num_classes if self.multiclass else 0)
        mk_g = lambda nm: Generator(nm, True, cfg.G_norm, batch_size=cfg.batch_size, output_height=H,
                                    output_width=half, input_dim=self.gf_dim, output_dim=self.c_dim,
                                    ops=ops, comm=self.comm,
                                    store=self._store(nm, generator_specs(nm, g_in, H, half, self.gf_dim, self.c_dim), rs))
        self.edge_generator = mk_g("G1")
        self.image_generator = mk_g("G2")
        if train:
            mk_d = lambda nm, h, w: Discriminator(nm, True, cfg.D_norm, num_filters=self.df_dim, ops=ops, in_hw=(h, w),
                                                  store=self._store(nm, discriminator_specs(nm, h, w, self.df_dim, self.c_dim), rs))
            self.joint_discriminator = mk_d("D", H, Wp)
            if cfg.use_image_discriminator:
                self.image_discriminator = mk_d("D_patch2", cfg.image_dis_size, cfg.image_dis_size)
            if cfg.use_edge_discriminator:
                self.edge_discriminator = mk_d("D_patch3", cfg.edge_dis_size, cfg.edge_dis_size)
        self.encoder = Encoder("E", True, cfg.E_norm, image_size=cfg.input_height, latent_dim=self.z_dim,
                               ops=ops, store=self._store("E", encoder_specs("E", cfg.input_height, self.z_dim, self.c_dim), rs))
        if train and self.multiclass:
            try:
                from .classifier import Classifier, classifier_specs
            except ImportError as e:
                raise NotImplementedError("the multi-class classifier run (d_optim2, BASELINE configs[2:]) is not "
                                          "built yet; use multiclasses=False") from e
            st = self._store("D2", classifier_specs("D2", cfg.num_classes, self.c_dim), rs, conv_filter_set=False)
            self.classifier = Classifier("D2", cfg.SPECTRAL_NORM_UPDATE_OPS, ops=ops, store=st, rs=rs,
                                         num_classes=cfg.num_classes)
            self.stores["D2/aux"] = self.classifier.aux      # unused disc head + frozen spectral-norm vectors
        self.losses = ops.zeros((16,))
        self._built = "train" if train else "test"
        # the tensor-core kernels' prepared filter copies are refreshed by ONE kernel per network right after each
        # write to its weights (ParamStore.load / .rmsprop, Classifier._normalise_weights): ops.filter_set

    def build_train_model(self):
        self.build_networks(True)

    def build_test_model(self):
        self.build_networks(False)

    # ---- variables -------------------------------------------------------------------------------
    def load_variables(self, values, strict=True):
        for st in self.stores.values():
            sub = {k: v for k, v in values.items() if k in st.offsets}
            st.load(sub, strict=strict)
        if hasattr(self, "classifier"):
            self.classifier.invalidate()

    def export_variables(self, what="var"):
        out = {}
        for st in self.stores.values():
            out.update(st.export(what))
        return out

    # ---- train / test loops (edgegan.py:425-489, 551-633) ----------------------------------------------------------
    def train(self, max_steps=None, prefetch_workers=None, log=print, use_graph=None):
        """edgegan.py:425-489.  Restores the latest checkpoint if there is one, then per epoch: shuffle, and per batch
        one `update_model` and the loss line of the reference.  The reference re-evaluates five loss tensors with three
        extra graph runs after every step (:461-477); here the line is printed from the scalars the step itself
        produced (`read_losses`, one 64-byte read-back), i.e. each loss is the value its own run minimised.
        Scalar summaries go to <logdir>/events.out.tfevents.* under the reference's tags (edgegan_b200/summary.py; image
        and histogram summaries are not written).  `max_steps` bounds the run (tests); `prefetch_workers` > 0 feeds the
        device through DevicePrefetcher (default: 8 on a GPU operator set, 0 otherwise).
        `use_graph` (default: on for a CUDA operator set): the first two iterations launch eagerly (they allocate every
        buffer of the step), the third is captured into a CUDA graph over static input buffers and from then on each
        iteration is a refill of those buffers plus one replay -- ~1 200 kernel launches per step without host work.
        alpha ~ U[0,1) and the encoder noise are drawn on the host per iteration exactly as in the eager path."""
        import time
        from ..utils.data import DevicePrefetcher
        cfg, ops = self.config, self.ops
        if self._built != "train":
            self.build_train_model()
        counter = 1
        start_time = time.time()
        loaded, checkpoint_counter = self.load(None, cfg.checkpoint_dir)
        if loaded:
            counter = checkpoint_counter
            log(" [*] Load SUCCESS")
        else:
            log(" [!] Load failed...")
        if prefetch_workers is None:
            prefetch_workers = 8 if getattr(getattr(ops, "device", None), "type", "cpu") == "cuda" else 0
        writer = None
        if cfg.logdir and self.comm.rank == 0:              # edgegan.py:443 nn.SummaryWriter(logdir, graph): scalars only here
            from ..summary import SummaryWriter
            writer = SummaryWriter(cfg.logdir)
        steps = 0
        if use_graph is None:
            use_graph = getattr(getattr(ops, "device", None), "type", "cpu") == "cuda"
        replay = _GraphStep(self) if use_graph else None
        for epoch in range(cfg.epoch):
            if self.comm.world_size > 1:
                # one global permutation per epoch (seed drawn on rank 0), disjoint shards per rank
                seed = ops.from_numpy(np.asarray([np.random.randint(0, 2 ** 24) if self.comm.rank == 0 else 0], np.float32))
                self.comm.allreduce(seed)
                self.dataset.shuffle(int(ops.to_numpy(seed)[0]), self.comm.rank, self.comm.world_size)
            else:
                self.dataset.shuffle()
            if prefetch_workers > 0:
                batches = DevicePrefetcher(self.dataset, ops, workers=prefetch_workers)
            else:
                batches = ((ops.from_numpy(im), ops.from_numpy(z.astype(np.float32)), f)
                           for im, z, f in (self.dataset[i] for i in range(len(self.dataset))))
            for idx, (batch_images, batch_z, _files) in enumerate(batches):
                if replay is not None:
                    replay(batch_images, batch_z)
                else:
                    self.update_model(batch_images, batch_z)
                L = self.read_losses()
                discriminator_err = L["joint_dis_dloss"] + L["image_dis_dloss"] + L["edge_dis_dloss"]
                generator_err = L["edge_gloss"] + L["image_gloss"]
                if writer is not None:
                    writer.add_losses(L, counter)
                counter += 1
                steps += 1
                log("Epoch: [%2d/%2d] [%4d/%4d] time: %4.4f, joint_dis_dloss: %.8f, joint_dis_gloss: %.8f"
                    % (epoch, cfg.epoch, idx, len(self.dataset), time.time() - start_time,
                       2 * discriminator_err, generator_err))
                if np.mod(counter, cfg.save_checkpoint_frequency) == 2:
                    self.save(None, cfg.checkpoint_dir, counter)
                    if writer is not None:
                        writer.flush()
                if max_steps is not None and steps >= max_steps:
                    if writer is not None:
                        writer.close()
                    return counter
        if writer is not None:
            writer.close()
        return counter

    def test(self, log=print):
        """edgegan.py:551-633: restore, then for every test picture E(left half) -> G1, G2 and write the combination
        chosen by `output_combination` below <test_output_dir>/<dataset>/<path below 'test'>.  Like the reference the
        edge and the image output come from two separate passes (two draws of the encoder noise).  The reference's
        'outputL_inputR' branch reads an undefined name; here it is the right half of the input."""
        import os
        from ..utils import pathsplit, save_images
        cfg, ops = self.config, self.ops
        if self._built is None:
            self.build_test_model()
        loaded, _ = self.load(None, cfg.checkpoint_dir)
        if not loaded:
            log(" [!] Load failed...")
            return 0
        log(" [*] Load SUCCESS")
        written = 0
        for idx in range(len(self.dataset)):
            batch_images, filenames = self.dataset[idx]
            keep = np.ones(len(filenames), bool)
            classes = None
            if self.multiclass:
                ids = []
                for k, path in enumerate(filenames):
                    try:
                        cid = int(pathsplit(path)[-2])
                        if cid >= cfg.num_classes:
                            raise ValueError
                        ids.append(cid)
                    except ValueError:
                        keep[k] = False
                if not ids:
                    continue
                classes = ops.from_numpy(np.asarray(ids, np.float32))
            x = ops.from_numpy(np.ascontiguousarray(batch_images[keep]))
            half = int(cfg.output_width / 2)
            outputL = ops.to_numpy(self.test_forward(x, classes, eps=float(np.random.normal()))[0])
            outputR = ops.to_numpy(self.test_forward(x, classes, eps=float(np.random.normal()))[1])
            inputs = batch_images[keep]
            if cfg.output_combination == "inputL_outputR":
                results = np.append(inputs[:, :, 0:half, :], outputR, axis=2)
            elif cfg.output_combination == "outputL_inputR":
                results = np.append(outputL, inputs[:, :, half:, :], axis=2)
            elif cfg.output_combination == "outputR":
                results = outputR
            else:
                results = np.append(np.append(inputs, outputL, axis=2), outputR, axis=2)
            names = [f for f, k in zip(filenames, keep) if k]
            for fname, img in zip(names, results):
                parts = pathsplit(fname)
                name = os.path.join(*parts[parts.index("test") + 1:])
                out = os.path.join(cfg.test_output_dir, cfg.dataset, name)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                save_images(img[np.newaxis, ...], [1, 1], out)
                written += 1
            log("Test: [%4d/%4d]" % (idx, len(self.dataset)))
        return written

    # ---- checkpoints (edgegan.py:421,547 tf.train.Saver over all global variables; :635-657 save / load) ---------
    @property
    def model_name(self):
        """edgegan.py:659-661."""
        return "EdgeGAN-Model"

    @staticmethod
    def _tf_name(name):
        """internal variable name -> the name tf.train.Saver stores.  Only the spectral-norm vectors differ: the
        reference creates `u` under variable_scope(W.name.rsplit('/', 1)[0]) opened INSIDE the layer scope
        (nn/modules/normalization.py:40-42), which doubles the path: D2/Conv/u -> D2/Conv/D2/Conv/u."""
        if name.endswith("/u"):
            scope = name[:-2]
            return f"{scope}/{scope}/u"
        return name

    def checkpoint_tensors(self):
        """Everything the reference's Saver writes: variables, the generators' batch-norm moving statistics,
        spectral-norm vectors, and per trained variable the RMSProp slots `<var>/RMSProp` (rms) and `<var>/RMSProp_1`
        (momentum, dead state at momentum = 0).
        The moving statistics are written at their INITIAL values (0 / 1).  The reference does update them in place on
        every generator forward (`tf.contrib.layers.batch_norm(..., updates_collections=None, is_training=True)`,
        nn/modules/normalization.py:21-25), so a reference-written checkpoint holds other numbers there; but
        `is_training` is hard-wired to True, so no graph -- training or test -- ever READS them and the difference
        cannot reach any output.  (How often the reference updates them depends on how many sess.run calls re-execute
        the generator per iteration, which this implementation deliberately does not imitate.)"""
        out = {}
        for key, st in self.stores.items():
            var = st.export("var")
            for n, a in var.items():
                out[self._tf_name(n)] = a
            if self._built == "train" and not key.endswith("/aux"):
                for n, a in st.export("ms").items():
                    out[n + "/RMSProp"] = a
                    out[n + "/RMSProp_1"] = np.zeros_like(a)
        for g in ("G1", "G2"):
            c = self.gf_dim * 8
            out[f"{g}/batch_norm/BatchNorm/moving_mean"] = np.zeros((c,), np.float32)
            out[f"{g}/batch_norm/BatchNorm/moving_variance"] = np.ones((c,), np.float32)
        return out

    def save(self, saver, checkpoint_dir, step):
        """edgegan.py:635-639: writes <checkpoint_dir>/EdgeGAN-Model-<step>.{index,data-00000-of-00001} in the
        TensorFlow bundle format and updates <checkpoint_dir>/checkpoint.  `saver` is unused (signature parity)."""
        from .. import checkpoint as ckpt
        import os
        print(" [*] Saving checkpoints...")
        os.makedirs(checkpoint_dir, exist_ok=True)
        prefix = os.path.join(checkpoint_dir, f"{self.model_name}-{int(step)}")
        if self.comm.rank == 0:
            ckpt.write_bundle(prefix, self.checkpoint_tensors())
            ckpt.update_checkpoint_state(checkpoint_dir, prefix)
        return prefix

    def load(self, saver, checkpoint_dir):
        """edgegan.py:641-657 -> (found, step).  Variables are matched by name; a checkpoint without the RMSProp slots
        (or a test-time model, which has none) loads the variables only.  Missing variables raise, like Saver.restore."""
        from .. import checkpoint as ckpt
        import os
        print(" [*] Reading checkpoints {}...".format(checkpoint_dir))
        state = ckpt.get_checkpoint_state(checkpoint_dir)
        if not state:
            print(" [*] Failed to find a checkpoint")
            return False, 0
        name = os.path.basename(state["model_checkpoint_path"])
        rd = ckpt.BundleReader(os.path.join(checkpoint_dir, name))
        try:
            for key, st in self.stores.items():
                vals, slots = {}, {}
                for n in st.offsets:
                    tn = self._tf_name(n)
                    if tn not in rd:
                        raise KeyError(f"{tn} not found in checkpoint {name}")
                    vals[n] = rd.tensor(tn)
                    if n + "/RMSProp" in rd:
                        slots[n] = rd.tensor(n + "/RMSProp")
                st.load(vals, strict=True)
                if slots and self._built == "train":
                    st.load(slots, strict=False, what="ms")
        finally:
            rd.close()
        if hasattr(self, "classifier"):
            self.classifier.invalidate()
        print(" [*] Success to read {}".format(name))
        return True, ckpt.step_of(name)

    def num_params(self):
        return {k: st.num_params for k, st in self.stores.items()}

    # ---- helpers -----------------------------------------------------------------------------------
    def _global_batch(self, n):
        return n * self.comm.world_size

    def _g_input(self, z):
        """edgegan.py:188-197: z[:, :z_dim] ++ one_hot(int(z[:, -1]))."""
        if not self.multiclass:
            return z
        n = z.shape[0]
        zin = self.ops.buf("step/z_onehot", (n, self.z_dim + self.config.num_classes))
        self.ops.onehot_concat(z, self.z_dim, self.config.num_classes, zin)
        return zin

    def _resize(self, x, size, key):
        """tf.image.resize_images(method=2) (edgegan.py:211-213); same-size resize is the identity."""
        n, H, W, C = x.shape
        if H == size and W == size:
            return x
        if size != 2 * H or size != 2 * W:
            raise NotImplementedError("only the 2x bicubic resize used by the reference configs is implemented")
        y = self.ops.buf(key, (n, size, size, C))
        self.ops.bicubic_up2_fwd(x, y)
        return y

    def _resize_bwd(self, gy, like, key):
        if gy.shape == like.shape:
            return gy
        gx = self.ops.buf(key, like.shape)
        self.ops.bicubic_up2_bwd(gy, gx)
        return gx

    def _apply(self, run, store_names):
        for s in store_names:
            st = self.stores[s]
            if self.comm.world_size > 1:
                self.comm.allreduce(st.grad)
        if self.run_hook is not None:
            self.run_hook(run, self)
        for s in store_names:
            self.stores[s].rmsprop(self.config.learning_rate)

    # ---- the critic run (runs 1-3) ---------------------------------------------------------------
    def _critic_run(self, run, D, X, alpha, slot):
        """X [3n,...]: rows [0:n) real and [n:2n) fake already filled; alpha [n]."""
        ops, cfg = self.ops, self.config
        n = X.shape[0] // 3
        inv_b = 1.0 / self._global_batch(n)
        loss = self.losses[slot:slot + 1]
        ops.gp_interpolate(X[0:n], X[n:2 * n], alpha, X[2 * n:])
        c = D.forward(X, "critic")
        ops.sum_scaled(c["d"][n:2 * n], inv_b, loss, False)          # mean(D(fake))
        ops.sum_scaled(c["d"][0:n], -inv_b, loss, True)              # - mean(D(real))   functional.py:32-33
        cx = D._sub(c, 2 * n, 3 * n)
        ddbar, addends, _ = D.penalty_sweeps(cx, cfg.lambda_gp, inv_b, loss)
        gd = ops.buf(f"{D.name}/critic/gd", (3 * n,))
        ops.fill(gd[0:n], -inv_b)
        ops.fill(gd[n:2 * n], inv_b)
        ops.gp_seed_bwd(cx["d"], ddbar, gd[2 * n:])
        D.backward(c, gd, "critic", param_grads=True, accumulate=True, addends=addends)
        self._apply(run, [D.name])

    # ---- generator run (runs 5 and 7) ------------------------------------------------------------
    def _generator_run(self, run, zin, labels_z, fresh_forward):
        ops, cfg = self.ops, self.config
        G1, G2 = self.edge_generator, self.image_generator
        if fresh_forward:
            e, i = G1.forward(zin), G2.forward(zin)
        else:
            e, i = G1.cache["h"][4], G2.cache["h"][4]
        n, H, half, C = e.shape
        inv_b = 1.0 / self._global_batch(n)
        sfx = "" if run == "g_optim_u" else "_b"
        l_e = self.losses[LOSS_SLOTS["edge_gloss" + sfx]:][:1]
        l_i = self.losses[LOSS_SLOTS["image_gloss" + sfx]:][:1]
        # joint critic on concat([edge, image], axis=2)
        joint = ops.buf("g/joint", (n, H, 2 * half, C))
        ops.copy_wslice(e, 0, joint, 0, half)
        ops.copy_wslice(i, 0, joint, half, half)
        D = self.joint_discriminator
        c = D.forward(joint, "gen")
        ops.sum_scaled(c["d"], -cfg.joint_dweight * inv_b, l_e, False)
        ops.sum_scaled(c["d"], -cfg.joint_dweight * inv_b, l_i, False)
        gd = ops.buf("g/gd_joint", (n,))
        ops.fill(gd, -cfg.joint_dweight * inv_b)
        gj = D.backward(c, gd, "gen", param_grads=False, input_grad=True)
        ge, gi = ops.buf("g/ge", e.shape), ops.buf("g/gi", i.shape)
        ops.copy_wslice(gj, 0, ge, 0, half)
        ops.copy_wslice(gj, half, gi, 0, half)
        for use, Dp, x, gx, w, lslot, key in (
                (cfg.use_edge_discriminator, getattr(self, "edge_discriminator", None), e, ge, cfg.edge_dweight, l_e, "edge"),
                (cfg.use_image_discriminator, getattr(self, "image_discriminator", None), i, gi, cfg.image_dweight, l_i, "image")):
            if not use:
                continue
            size = cfg.edge_dis_size if key == "edge" else cfg.image_dis_size
            up = self._resize(x, size, f"g/{key}_up")
            cp = Dp.forward(up, "gen")
            ops.sum_scaled(cp["d"], -w * inv_b, lslot, True)
            gdp = ops.buf(f"g/gd_{key}", (n,))
            ops.fill(gdp, -w * inv_b)
            gup = Dp.backward(cp, gdp, "gen", param_grads=False, input_grad=True)
            gxp = self._resize_bwd(gup, x, f"g/{key}_gdown")
            ops.axpby(gxp, gx, 1.0, 1.0)
        if self.multiclass:
            l_ac = self.losses[LOSS_SLOTS["loss_g_ac"]:][:1]
            gcls = self.classifier.fake_loss_input_grad(i, labels_z, 0.5, inv_b, l_ac)   # functional.py:13-15
            ops.axpby(l_ac, l_i, 1.0, 1.0)
            ops.axpby(gcls, gi, 1.0, 1.0)
        G1.backward(ge)
        G2.backward(gi)
        self._apply(run, ["G1", "G2"])

    # ---- encoder run (run 6) -----------------------------------------------------------------------
    def _encoder_run(self, run, zin, z, eps):
        ops, cfg = self.ops, self.config
        e = self.edge_generator.forward(zin)          # G1 was just updated by run 5
        E = self.encoder
        _, mu, ls = E.forward(e, eps)
        n = e.shape[0]
        gmu, gls = ops.buf("e/gmu", mu.shape), ops.buf("e/gls", ls.shape)
        loss = self.losses[LOSS_SLOTS["zl_loss"]:][:1]
        ops.fill(loss, 0.0)
        ops.zl1_loss_bwd(mu, ls, eps, z, cfg.stage1_zl_loss, 1.0 / (self._global_batch(n) * self.z_dim), gmu, gls, loss)
        E.backward(gmu, gls)
        self._apply(run, ["E"])

    # ---- the step ------------------------------------------------------------------------------------
    def update_model(self, images, z, alpha=None, eps=None, runs=None):
        """One training iteration (edgegan.py:126-130).

        images [n,H,W_pair,3] (edge | image), z [n, z_dim(+1)], alpha [3,n] (one U[0,1) vector per
        critic), eps: the scalar encoder noise.  All device float32 tensors (eps a python float).
        When alpha / eps are None they are drawn from a host RNG like the reference's in-graph
        sampling."""
        if self._built != "train":
            self.build_train_model()
        ops, cfg = self.ops, self.config
        n, H, Wp, C = images.shape
        half = Wp // 2
        if alpha is None:
            rs = getattr(self, "_rs", None) or np.random.RandomState(self.seed + 1)
            self._rs = rs
            alpha = ops.from_numpy(rs.uniform(0, 1, (3, n)).astype(np.float32))
            eps = float(rs.normal()) if eps is None else eps
        if eps is None:
            eps = 0.0
        elif not hasattr(eps, "data_ptr"):
            eps = float(eps)          # python scalar; a 1-element device tensor keeps a captured graph re-playable
        todo = runs or [r for r in RUN_NAMES if r != "d_optim2" or self.multiclass]
        zin = self._g_input(z)
        G1, G2 = self.edge_generator, self.image_generator
        fakes_fresh = False

        def fakes():
            nonlocal fakes_fresh
            if not fakes_fresh:
                G1.forward(zin)
                G2.forward(zin)
                fakes_fresh = True
            return G1.cache["h"][4], G2.cache["h"][4]

        for run in todo:
            if run == "d_optim":
                e, i = fakes()
                D = self.joint_discriminator
                X = ops.buf("D/X", (3 * n, H, Wp, C))
                ops.copy(images, X[0:n])
                ops.copy_wslice(e, 0, X[n:2 * n], 0, half)
                ops.copy_wslice(i, 0, X[n:2 * n], half, half)
                self._critic_run(run, D, X, alpha[0], LOSS_SLOTS["joint_dis_dloss"])
            elif run in ("d_optim_patch2", "d_optim_patch3"):
                img = run == "d_optim_patch2"
                if not (cfg.use_image_discriminator if img else cfg.use_edge_discriminator):
                    continue
                e, i = fakes()
                D = self.image_discriminator if img else self.edge_discriminator
                size = cfg.image_dis_size if img else cfg.edge_dis_size
                X = ops.buf(f"{D.name}/X", (3 * n, size, size, C))
                real_half = ops.buf("step/real_half", (n, H, half, C))
                ops.copy_wslice(images, half if img else 0, real_half, 0, half)
                fake = i if img else e
                if size == H and size == half:
                    ops.copy(real_half, X[0:n])
                    ops.copy(fake, X[n:2 * n])
                else:
                    ops.bicubic_up2_fwd(real_half, X[0:n])
                    ops.bicubic_up2_fwd(fake, X[n:2 * n])
                self._critic_run(run, D, X, alpha[1 if img else 2],
                                 LOSS_SLOTS["image_dis_dloss" if img else "edge_dis_dloss"])
            elif run == "d_optim2":
                real_pic = ops.buf("step/real_pic", (n, H, half, C))
                ops.copy_wslice(images, half, real_pic, 0, half)
                loss = self.losses[LOSS_SLOTS["loss_d_ac"]:][:1]
                self.classifier.real_loss_step(real_pic, z, 1.0 / self._global_batch(n), loss)
                self._apply(run, ["D2"])
            elif run == "g_optim_u":
                fresh = not fakes_fresh
                fakes_fresh = False
                self._generator_run(run, zin, z, fresh_forward=fresh)
            elif run == "e_optim":
                self._encoder_run(run, zin, z, eps)
                fakes_fresh = False
            elif run == "g_optim_b":
                self._generator_run(run, zin, z, fresh_forward=True)
                fakes_fresh = False

    def capture_step(self, images, z, alpha, eps, warmup=2, stream=None):
        """Capture one update_model into a CUDA graph over STATIC device buffers (images, z, alpha and a 1-element
        device tensor eps): refill the buffers, then `graph.replay()`.  The step allocates nothing after its first
        call and takes every scalar input from device memory, so the ~650 launches replay without host work.
        Runs `warmup` real steps first (they update the weights like any other step); they also let the library size its
        per-stream scratch memory on the capture stream, so with warmup=0 pass the `stream` earlier steps already ran on."""
        import torch
        s = stream or torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.update_model(images, z, alpha, eps)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # data parallel: the NCCL all-reduces (gradients, sync-BN sums) are captured with the kernels -- the communicator
        # exists by now (the warm-up steps above ran them eagerly).  NCCL's watchdog thread polls CUDA events while we
        # capture, which the default "global" capture mode would turn into a capture error.
        mode = "thread_local" if self.comm.world_size > 1 else "global"
        with torch.cuda.graph(g, stream=s, capture_error_mode=mode):
            self.update_model(images, z, alpha, eps)
        return g

    def read_losses(self):
        """host copy of the loss scalars of the last step (one device->host read)."""
        t = self.losses
        if self.comm.world_size > 1:      # each rank holds its shard's share (already divided by the global batch)
            t = self.ops.buf("step/losses_global", self.losses.shape)
            self.ops.copy(self.losses, t)
            self.comm.allreduce(t)
        host = self.ops.to_numpy(t)
        return {k: float(host[i]) for k, i in LOSS_SLOTS.items()}

    # ---- inference (edgegan.py:492-517) -----------------------------------------------------------
    def test_forward(self, inputs, classes=None, eps=0.0):
        """inputs [n,H,W_pair,3]: E(left half) -> z (++ one_hot(class)) -> (G1(z), G2(z))."""
        if self._built is None:
            self.build_test_model()
        ops, cfg = self.ops, self.config
        n, H, Wp, C = inputs.shape
        half = Wp // 2
        left = ops.buf("test/left", (n, H, half, C))
        ops.copy_wslice(inputs, 0, left, 0, half)
        z, _, _ = self.encoder.forward(left, eps, tag="test")
        if self.multiclass:
            zc = ops.buf("test/zc", (n, self.z_dim + 1))
            ops.copy_wslice(z.view(n, 1, self.z_dim, 1), 0, zc.view(n, 1, self.z_dim + 1, 1), 0, self.z_dim)
            ops.copy_wslice(classes.view(n, 1, 1, 1), 0, zc.view(n, 1, self.z_dim + 1, 1), self.z_dim, 1)
            zin = ops.buf("test/z_onehot", (n, self.z_dim + cfg.num_classes))
            ops.onehot_concat(zc, self.z_dim, cfg.num_classes, zin)
        else:
            zin = z
        return self.edge_generator.forward(zin, tag="test"), self.image_generator.forward(zin, tag="test")
