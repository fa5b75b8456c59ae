#!/usr/bin/env python
"""bench.py -- EdgeGAN G + 3xD(+GP) + E training-step throughput (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--algo tc|tc3x|simt]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
              --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one `EdgeGAN.update_model` (reference edgegan/models/edgegan.py:126-130): the 7 sequential
RMSProp runs of the 14-class model (6 without the classifier) on one synthetic batch.  Default workload =
BASELINE.json configs[2]: 14-class 64x64 training, batch 128 per GPU (the reference's own default,
train.py:44-45, and the config the north-star targets are stated on); weak scaling for N > 1, so `--gpus 8`
IS configs[3] (global batch 1024; gradients all-reduced over NCCL, sync-BN sums all-reduced).  `--config 2`
= single-class 64x64 batch 64 (configs[1]); `--config 5` = 14-class 128x128 batch 64 per GPU (configs[4], 4 GPUs).
At N = 1 the default run also times configs[1] as a second leg and reports it under the extra key "config2".

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = images/s through the
public API with pinned HOST buffers (H2D of images/z/alpha and D2H of the losses inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic GFLOP per image per update_model (2*MAC of conv / conv-transpose / linear)
GFLOP_PER_IMAGE = {2: 43.49, 3: 93.76, 5: 295.32}
# BASELINE.json configs -> (multiclass, per-GPU batch, height, pair width, label)
CONFIGS = {2: (False, 64, 64, 128, "single-class 64x64"), 3: (True, 128, 64, 128, "14-class 64x64"),
           5: (True, 64, 128, 256, "14-class 128x128")}
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "tc", "tc3x", "simt"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 5], help="BASELINE.json config number (1-based)")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--multiclass", action="store_true", help="(kept for compatibility) same as --config 3")
    ap.add_argument("--no-second-leg", action="store_true", help="skip the configs[1] leg of the default N = 1 run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile-pass", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--ncu", action="store_true", help="profiling run: warmup/steps as given, no e2e / cpu / roofline passes; the printed number is NOT a bench value")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": (sorted(power)[len(power) // 2] if power else None), "samples": len(sm),
                "reasons": sorted(reasons)}


def resolve(args):
    """-> (config number, multiclass, per-GPU batch, H, pair width, label)"""
    cfg = 3 if args.multiclass else args.config
    multiclass, B, H, W, label = CONFIGS[cfg]
    if args.batch is not None:
        B = args.batch
    return cfg, multiclass, B, H, W, label


def workload_config(cfg, label, B, world):
    """`config` of the JSON line: identical in both arms (the reference arm runs on OUR arm's config)."""
    return {"workload": f"{label} EdgeGAN update_model (G1+G2, 3 critics with WGAN-GP, "
                        f"{'classifier D2, ' if CONFIGS[cfg][0] else ''}E; BASELINE.json configs[{cfg - 1}]), batch {B}/GPU",
            "baseline_config": cfg, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "gflop_per_image": GFLOP_PER_IMAGE[cfg],
            "l2": "no flush: one step rewrites a multi-GB activation working set (>> 126 MB L2) and the inputs rotate over 4 sets"}


def oracle_step_time(batch, multiclass, steps, warmup, threads, H=64, W=128):
    """The reference's CPU path restated (oracle/edgegan_oracle.py, torch-CPU fp32) -- TF 1.14 itself cannot run
    here.  Returns seconds per step (mean of `steps`)."""
    import torch
    from oracle import edgegan_oracle as O
    torch.set_num_threads(threads)
    dis = 128
    cfg = O.Config(batch_size=batch, multiclasses=multiclass, output_height=H, output_width=W,
                   image_dis_size=dis, edge_dis_size=dis)
    v, u = O.init_variables(cfg, seed=0)
    st = O.OracleState(cfg, v, u)
    ts = []
    for i in range(warmup + steps):
        inp = O.make_inputs(cfg, seed=100 + i)
        t0 = time.perf_counter()
        O.update_model(st, inp)
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    return sum(ts) / len(ts)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the restated oracle: TF 1.14 cannot be
    installed), all host threads, SAME config and batch as our arm; each step is one full update_model."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cfg, multiclass, B, H, W, label = resolve(args)
    # bounded run: the whole arm has to end within a few minutes (driver limit 1800 s).  A batch-128 14-class step
    # costs ~10 s on 16-32 cores, so long --steps/--warmup requests are served by fewer timed repetitions of the
    # SAME full-batch step (stated in `sample`); the per-step workload is never shrunk.
    probe_t0 = time.perf_counter()
    sec_probe = oracle_step_time(B, multiclass, 1, 0, threads, H, W)          # also the warm-up of the process
    budget = 600.0 - (time.perf_counter() - probe_t0)
    steps = max(1, min(args.steps, int(budget / max(sec_probe, 1e-3)) - 1))
    sec = oracle_step_time(B, multiclass, steps, 1 if steps > 1 else 0, threads, H, W)
    val = B / sec
    sample = (f"{steps} full update_model steps at batch {B} of the same {label} workload after warm-up "
              f"({sec:.1f} s per step, torch-CPU fp32 restatement of the reference, {threads} threads)")
    line = {
        "impl": "reference", "metric": "G+D+GP step images/sec at 64x64", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, label, B, max(1, args.gpus)),
        "timed_steps": steps,
        "note": "restated-reference CPU path (torch-CPU fp32 oracle); TensorFlow 1.14 cannot be installed here (SURVEY.md D9)",
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Leg:
    """One workload (a BASELINE config) built on this rank: model, synthetic inputs, graph; timing helpers."""

    def __init__(self, args, cfg, multiclass, B, H, W, ops, comm, world, rank):
        import numpy as np
        import torch
        from edgegan_b200.config import Flags
        from edgegan_b200.models.edgegan import EdgeGAN
        self.args, self.ops, self.comm, self.world, self.rank, self.B = args, ops, comm, world, rank, B
        flags = Flags(batch_size=B, multiclasses=multiclass, input_height=H, input_width=W, output_height=H,
                      output_width=W)
        if not multiclass:
            flags.num_classes = None
        self.model = model = EdgeGAN(None, flags, None, ops=ops, comm=comm, seed=1234)
        model.build_train_model()
        # synthetic data: images ~ U(-1,1), z ~ N(0,1) (+ class id), alpha ~ U(0,1), scalar eps
        rs = np.random.RandomState(2333 + rank)
        self.n_sets = n_sets = 4
        self.host = host = []
        for _ in range(n_sets):
            img = torch.from_numpy(rs.uniform(-1, 1, (B, H, W, 3)).astype(np.float32)).pin_memory()
            z = rs.normal(size=(B, 100)).astype(np.float32)
            if multiclass:
                z = np.concatenate([z, rs.randint(0, 14, (B, 1)).astype(np.float32)], 1)
            z = torch.from_numpy(z).pin_memory()
            al = torch.from_numpy(rs.uniform(0, 1, (3, B)).astype(np.float32)).pin_memory()
            host.append((img, z, al, float(rs.normal())))
        self.dev_sets = [(i.cuda(non_blocking=True), z.cuda(non_blocking=True), a.cuda(non_blocking=True), e)
                         for i, z, a, e in host]
        self.d_in = tuple(torch.empty_like(t, device=ops.device) for t in host[0][:3])
        self.loss_host = torch.empty(16, dtype=torch.float32).pin_memory()
        self.h2d_bytes = sum(t.numel() * 4 for t in host[0][:3])
        self.d2h_bytes = self.loss_host.numel() * 4
        # the step as a CUDA graph over static input buffers (falls back to eager launches if capture is refused)
        self.graph, self.graph_note = None, "eager launches"
        self.s_in = tuple(torch.empty_like(t, device=ops.device) for t in host[0][:3])
        self.s_eps = torch.zeros(1, dtype=torch.float32, device=ops.device)
        self.eps_host = [torch.tensor([e], dtype=torch.float32).pin_memory() for _, _, _, e in host]
        self.eps_dev = [t.cuda() for t in self.eps_host]
        if not args.no_graph and not args.ncu:
            try:
                for s, d in zip(self.s_in, self.dev_sets[0][:3]):
                    s.copy_(d)
                self.graph = model.capture_step(*self.s_in, self.s_eps)
                self.graph_note = "CUDA graph replay" + (" (NCCL all-reduces captured)" if world > 1 else "")
            except Exception as e:          # noqa: BLE001
                self.graph, self.graph_note = None, f"eager launches (graph capture failed: {type(e).__name__}: {e})"[:200]
                torch.cuda.synchronize()
        # kernels per step, counted by one eager step (the graph replays exactly these)
        l0 = ops.launches
        model.update_model(*self.dev_sets[0])
        torch.cuda.synchronize()
        self.launches_per_step = ops.launches - l0

    def barrier(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            self.comm.barrier()
            torch.cuda.synchronize()

    def step_resident(self, k):
        i, z, a, e = self.dev_sets[k % self.n_sets]
        if self.graph is None:
            self.model.update_model(i, z, a, e)
        else:                            # inputs already in HBM: device-to-device refill of the static buffers
            for s, d in zip(self.s_in, (i, z, a)):
                s.copy_(d, non_blocking=True)
            self.s_eps.copy_(self.eps_dev[k % self.n_sets], non_blocking=True)
            self.graph.replay()

    def step_e2e(self, k):
        import torch
        i, z, a, e = self.host[k % self.n_sets]
        if self.graph is None:
            for d, h in zip(self.d_in, (i, z, a)):
                d.copy_(h, non_blocking=True)
            self.model.update_model(*self.d_in, e)
        else:
            for s, h in zip(self.s_in, (i, z, a)):
                s.copy_(h, non_blocking=True)
            self.s_eps.copy_(self.eps_host[k % self.n_sets], non_blocking=True)
            self.graph.replay()
        self.loss_host.copy_(self.model.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the user reads the losses every step

    def timed(self, fn, steps, warmup):
        import torch
        for k in range(warmup):
            fn(k)
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.ops.launches
        ev0.record()
        for k in range(steps):
            fn(warmup + k)
        ev1.record()
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        launches = self.ops.launches - l0
        if self.graph is not None:       # replayed launches do not pass through ops.*: count what the graph holds
            launches = steps * self.launches_per_step
        if self.world > 1:
            t = torch.tensor([ms], device=self.ops.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches

    def release(self):
        import torch
        self.graph = None
        self.model = None
        self.ops._bufs.clear()
        torch.cuda.empty_cache()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch

    from edgegan_b200.models.edgegan import LocalComm
    from edgegan_b200.ops import DeviceOps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: edgegan_b200 needs a CUDA device (B200, sm_100a); it has no CPU path "
                 "(`--impl reference` times the CPU restatement of the reference)")
    torch.cuda.set_device(local)
    comm = LocalComm()
    if world > 1:
        from edgegan_b200.comm import TorchDistComm
        comm = TorchDistComm("nccl")
    cfg, multiclass, B, H, W, label = resolve(args)
    ops = DeviceOps(f"cuda:{local}")
    algo = "tc3x" if args.algo == "auto" else args.algo
    ops.set_default_algo(algo)
    leg = Leg(args, cfg, multiclass, B, H, W, ops, comm, world, rank)
    model = leg.model

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if args.ncu:
        for k in range(args.warmup + args.steps):
            leg.step_resident(k)
        torch.cuda.synchronize()
        print(json.dumps({"ncu_run": True, "launches_per_step": leg.launches_per_step}))
        return
    warm = max(args.warmup, 3)
    ms_step, launches = leg.timed(leg.step_resident, args.steps, warm)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = leg.timed(leg.step_e2e, args.steps, 1)
    losses = model.read_losses()
    finite = all(np.isfinite(v) for v in losses.values())

    # ---- roofline of the dominant kernel family (the implicit-GEMM conv kernels), timed live with CUDA events
    roof = None
    if not args.no_profile_pass:
        roof = conv_roofline(model, ops, lambda k: model.update_model(*leg.dev_sets[k % leg.n_sets]), algo)
    ws_gb = ops.bytes_allocated() / 1e9
    graph_note = leg.graph_note
    h2d_bytes, d2h_bytes = leg.h2d_bytes, leg.d2h_bytes

    # ---- second leg (default N = 1 run only): BASELINE configs[1], reported under "config2"
    second = None
    if world == 1 and cfg == 3 and args.batch is None and not args.no_second_leg:
        leg.release()
        mc2, B2, H2, W2, label2 = CONFIGS[2]
        leg2 = Leg(args, 2, mc2, B2, H2, W2, ops, comm, world, rank)
        ms2, l2n = leg2.timed(leg2.step_resident, args.steps, warm)
        ms2e, _ = leg2.timed(leg2.step_e2e, args.steps, 1)
        second = {"config": workload_config(2, label2, B2, 1), "value": B2 / (ms2 * 1e-3), "unit": "images/s",
                  "ms_per_step": ms2, "e2e": {"value": B2 / (ms2e * 1e-3), "unit": "images/s", "ms_per_step": ms2e,
                                              "h2d_bytes_per_step": leg2.h2d_bytes, "d2h_bytes_per_step": leg2.d2h_bytes},
                  "gpu_launches": l2n, "launch": leg2.graph_note, "step_tflops": B2 * GFLOP_PER_IMAGE[2] / ms2}
        leg2.release()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:       # reported at N = 1 only (the other ranks would idle in a barrier)
        threads = os.cpu_count() or 1
        sec = oracle_step_time(B, multiclass, 1, 1, threads, H, W)     # ~20-30 s of CPU work on 16 cores
        cpu = {"value": B / sec, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"1 update_model step (after 1 warm-up step) at batch {B} of the same {label} "
                         f"workload ({sec:.1f} s per step, torch-CPU fp32 oracle, {threads} threads)"}

    if rank == 0:
        gimg = B * world
        line = {
            "metric": "G+D+GP step images/sec at 64x64", "value": gimg / (ms_step * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tc": "tf32", "tc3x": "3xtf32", "simt": "f32"}[algo], "data": "synthetic",
            "config": workload_config(cfg, label, B, world),
            "impl_detail": {"conv_algo": algo, "launch": graph_note, "working_set_gb": round(ws_gb, 2),
                            "losses_finite": finite},
            "e2e": {"value": gimg / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "step_tflops": gimg * GFLOP_PER_IMAGE[cfg] / ms_step,      # GFLOP / ms = TFLOP/s (whole job)
            "config2": second,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down in a fixed order: the captured graph (it references NCCL kernels) first, then a last rendezvous so
        # that no rank leaves while another still needs it.  destroy_process_group() has been seen to hang after a graph
        # with captured collectives (the bench line was out, the launcher then waited for its timeout), so the
        # communicator is left to process exit.
        leg.release()
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def conv_roofline(model, ops, step_fn, algo):
    """Instrument every conv launch of one extra step with CUDA events on the launching stream; aggregate the
    tensor-core implicit-GEMM launches: achieved = sum(2*MAC) / sum(duration)."""
    import torch
    pk, src = peaks()
    recs = []
    orig = {n: getattr(ops, n) for n in ("conv_fwd", "conv_bwd_data", "conv_bwd_weight")}

    def wrap(name, pass_id):
        f = orig[name]

        def g(*a, **kw):
            if name == "conv_fwd":
                x, w, y = a[0], a[1], a[3]
                macs = y.numel() * w.shape[0] * w.shape[1] * w.shape[2]
                stride, pad = a[4], a[5]
                s = ops._cs(x.shape, w.shape, y.shape, stride, pad)
            elif name == "conv_bwd_data":
                dy, w, dx = a[0], a[1], a[3]
                macs = dy.numel() * w.shape[0] * w.shape[1] * w.shape[2]
                s = ops._cs(dx.shape, w.shape, dy.shape, a[4], a[5])
            else:
                x, dy, dw = a[0], a[1], a[2]
                macs = dy.numel() * dw.shape[0] * dw.shape[1] * dw.shape[2]
                s = ops._cs(x.shape, dw.shape, dy.shape, a[3], a[4])
            import ctypes as C
            used = ops.lib.eg_conv2d_algo_for(C.byref(s), pass_id, 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f(*a, **kw)
            e1.record()
            shape = f"N{s.N} {s.H}x{s.W}x{s.Ci}->{s.OH}x{s.OW}x{s.Co} k{s.KH} s{s.stride}"
            thin = min(s.Ci, s.Co) <= 8
            nbytes = 4.0 * (s.N * s.H * s.W * s.Ci + s.N * s.OH * s.OW * s.Co)      # algorithmic: input + output once
            recs.append((name, used, 2.0 * macs, e0, e1, shape, thin, nbytes))
        return g

    for pid, n in enumerate(("conv_fwd", "conv_bwd_data", "conv_bwd_weight")):
        setattr(ops, n, wrap(n, pid))
    try:
        step_fn(0)
        torch.cuda.synchronize()
    finally:
        for n, f in orig.items():
            setattr(ops, n, f)
    # Three classes of conv launches: "tc" = tensor-bound tcgen05 layers; "thin" = the image-side layers (<= 8 channels
    # on one side): tensor cores or FFMA, but HBM-bound by arithmetic intensity (d_conv_0 AI 20, g_dconv_4
    # 31, classifier h0 / unit-1 34-53 flop/B) -> reported in GB/s against the HBM roofline; "simt" = the rest on FFMA.
    agg, by_shape = {}, {}
    for name, used, fl, e0, e1, shape, thin, nbytes in recs:
        path = "thin" if thin else ("tc" if used in (2, 3) else "simt")
        key = (name, path)
        ms = e0.elapsed_time(e1)
        for d, k in ((agg, key), (by_shape, key + (shape,))):
            a = d.setdefault(k, [0.0, 0.0, 0, 0.0])
            a[0] += fl
            a[1] += ms
            a[2] += 1
            a[3] += nbytes
    tc_fl = sum(v[0] for k, v in agg.items() if k[1] == "tc")
    tc_ms = sum(v[1] for k, v in agg.items() if k[1] == "tc")
    thin_b = sum(v[3] for k, v in agg.items() if k[1] == "thin")
    thin_ms = sum(v[1] for k, v in agg.items() if k[1] == "thin")
    all_ms = sum(v[1] for v in agg.values())
    def num(v):                      # plain number, or an object carrying it under "value"
        if isinstance(v, dict):
            v = v.get("value")
        try:
            return float(v)
        except (TypeError, ValueError):
            return None
    peak = num(pk.get("bf16_tflops_sustained")) or num(pk.get("bf16_tflops")) or FALLBACK_PEAKS["bf16_tflops_sustained"]
    if peak > 1e5:                   # given in GFLOP/s
        peak /= 1e3
    # kind::tf32 runs at half the bf16 rate: the peak of the MMA kind actually used
    peak_kind = peak / 2.0
    hbm = num(pk.get("hbm_gbs")) or FALLBACK_PEAKS["hbm_gbs"]
    detail = {f"{k[0]}/{k[1]}": {"launches": v[2], "ms": round(v[1], 3), "tflops": (v[0] / v[1] / 1e9 if v[1] > 0 else None),
                                 **({"gbs": v[3] / v[1] / 1e6, "hbm_frac": v[3] / v[1] / 1e6 / hbm} if k[1] == "thin" and v[1] > 0 else {})}
              for k, v in agg.items()}
    top = sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:24]
    shapes = [{"pass": k[0], "path": k[1], "shape": k[2], "launches": v[2], "ms": round(v[1], 3),
               "tflops": round(v[0] / v[1] / 1e9, 1) if v[1] > 0 else None,
               "gbs": round(v[3] / v[1] / 1e6, 0) if v[1] > 0 else None} for k, v in top]
    thin_shapes = [{"pass": k[0], "shape": k[2], "launches": v[2], "ms": round(v[1], 3),
                    "gbs": round(v[3] / v[1] / 1e6, 0) if v[1] > 0 else None,
                    "hbm_frac": round(v[3] / v[1] / 1e6 / hbm, 3) if v[1] > 0 else None}
                   for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][1]) if k[1] == "thin"]
    # DRAM traffic of the dominant launch class from the committed ncu --set full capture (tools/ncu_evidence.sh), next to its
    # live CUDA-event throughput of this run
    traffic, dominant = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        live = [x for x in shapes if x["shape"] == tj["shape"] and x["pass"] == "conv_fwd" and x["path"] == "tc"]
        traffic = tj["dram_bytes_per_launch"]
        dominant = {"kernel": tj["kernel"], "shape": tj["shape"], "algorithmic_bytes_per_launch": tj["algorithmic_bytes_per_launch"],
                    "dram_bytes_per_launch": tj["dram_bytes_per_launch"], "tensor_pipe_active_pct_ncu": tj["tensor_pipe_active_pct"],
                    "live_tflops": live[0]["tflops"] if live else None,
                    "live_frac_of_tf32_peak": (live[0]["tflops"] / peak_kind) if live else None, "source": tj["source"]}
    except Exception:
        pass
    if tc_ms > 0:
        ach = tc_fl / tc_ms / 1e9
        mult = 3.0 if algo == "tc3x" else 1.0
        return {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv (fwd/dgrad/wgrad), kind::tf32",
                "achieved": ach, "peak": peak_kind, "unit": "TFLOP/s", "frac": ach / peak_kind, "traffic": traffic,
                "dominant_launch": dominant,
                "peak_source": f"{src} bf16 sustained / 2 (tf32 rate)", "conv_ms_per_step": all_ms,
                "mma_per_algorithmic_flop": mult,
                "tensor_pipe_work_frac": mult * ach / peak_kind,
                "note": ("achieved counts ALGORITHMIC conv FLOPs; the 3xTF32 mode issues 3 tensor-core MMAs per algorithmic "
                         "MMA (hi*hi + lo*hi + hi*lo), so the tensor pipe is busy tensor_pipe_work_frac of its tf32 peak"),
                "thin_layers": {"bound": "hbm", "achieved": thin_b / thin_ms / 1e6 if thin_ms > 0 else None, "peak": hbm, "unit": "GB/s",
                                "frac": thin_b / thin_ms / 1e6 / hbm if thin_ms > 0 else None, "ms_per_step": thin_ms,
                                "note": "image-side layers (<= 8 channels on one side): algorithmic bytes = 4*(input + output) per launch",
                                "by_shape": thin_shapes},
                "detail": detail, "by_shape": shapes}
    simt_fl = sum(v[0] for v in agg.values())
    ach = simt_fl / all_ms / 1e9 if all_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "fp32 FFMA implicit-GEMM conv (SIMT path; tensor cores not used)",
            "achieved": ach, "peak": peak_kind, "unit": "TFLOP/s", "frac": ach / peak_kind, "traffic": None,
            "peak_source": f"{src} bf16 sustained / 2 (tf32 rate)", "conv_ms_per_step": all_ms, "detail": detail}


if __name__ == "__main__":
    main()
