#!/usr/bin/env python
"""bench.py -- EdgeGAN G + 3xD(+GP) + E training-step throughput (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--algo tc|tc3x|simt]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
              --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one `EdgeGAN.update_model` (reference edgegan/models/edgegan.py:126-130): the 6 sequential
RMSProp runs of the single-class model (7 with the classifier) on one synthetic batch.  Workload at N = 1:
BASELINE.json configs[1] = single-class 64x64 training, batch 64 per GPU; weak scaling for N > 1 (batch 64
per rank, gradients all-reduced over NCCL, sync-BN sums all-reduced).

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
GFLOP_PER_IMAGE = {"single": 43.49, "multi": 93.76}
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "tc", "tc3x", "simt"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--multiclass", action="store_true")
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


def oracle_step_time(batch, multiclass, steps, warmup, threads):
    """The reference's CPU path restated (oracle/edgegan_oracle.py, torch-CPU fp32) -- TF 1.14 itself cannot run
    here.  Returns seconds per step (mean of `steps`)."""
    import torch
    from oracle import edgegan_oracle as O
    torch.set_num_threads(threads)
    cfg = O.Config(batch_size=batch, multiclasses=multiclass)
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
    """--impl reference: the reference's own CPU implementation of the path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    b = 64
    while b > 4 and b * total > 640:      # ~3 images/s on 8 cores -> keep the whole run within a few minutes
        b //= 2
    mode = "multi" if args.multiclass else "single"
    sec = oracle_step_time(b, args.multiclass, args.steps, args.warmup, threads)
    val = b / sec
    sample = f"{args.steps} steps of one update_model at batch {b} (bounded sample of the batch-{args.batch} workload)"
    line = {
        "impl": "reference", "metric": "G+D+GP step images/sec at 64x64", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{mode}-class 64x64 EdgeGAN update_model, batch {b} (CPU sample)",
                   "note": "restated-reference CPU path (torch-CPU fp32 oracle); TensorFlow 1.14 cannot be installed here"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch

    from edgegan_b200.config import Flags
    from edgegan_b200.models.edgegan import EdgeGAN, LocalComm
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
    B = args.batch
    multiclass = args.multiclass
    mode = "multi" if multiclass else "single"
    flags = Flags(batch_size=B, multiclasses=multiclass)
    if not multiclass:
        flags.num_classes = None
    ops = DeviceOps(f"cuda:{local}")
    algo = "tc3x" if args.algo == "auto" else args.algo
    ops.set_default_algo(algo)
    model = EdgeGAN(None, flags, None, ops=ops, comm=comm, seed=1234)
    model.build_train_model()

    # synthetic data: images ~ U(-1,1), z ~ N(0,1) (+ class id), alpha ~ U(0,1), scalar eps
    rs = np.random.RandomState(2333 + rank)
    n_sets = 4
    host = []
    for _ in range(n_sets):
        img = torch.from_numpy(rs.uniform(-1, 1, (B, 64, 128, 3)).astype(np.float32)).pin_memory()
        z = rs.normal(size=(B, 100)).astype(np.float32)
        if multiclass:
            z = np.concatenate([z, rs.randint(0, 14, (B, 1)).astype(np.float32)], 1)
        z = torch.from_numpy(z).pin_memory()
        al = torch.from_numpy(rs.uniform(0, 1, (3, B)).astype(np.float32)).pin_memory()
        host.append((img, z, al, float(rs.normal())))
    dev_sets = [(i.cuda(non_blocking=True), z.cuda(non_blocking=True), a.cuda(non_blocking=True), e) for i, z, a, e in host]
    d_img, d_z, d_al = (torch.empty_like(t, device=ops.device) for t in host[0][:3])
    loss_host = torch.empty(16, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(t.numel() * 4 for t in host[0][:3])
    d2h_bytes = loss_host.numel() * 4

    # the step as a CUDA graph over static input buffers (falls back to eager launches if capture is refused)
    graph, graph_note = None, "eager launches"
    s_img, s_z, s_al = (torch.empty_like(t, device=ops.device) for t in host[0][:3])
    s_eps = torch.zeros(1, dtype=torch.float32, device=ops.device)
    eps_host = [torch.tensor([e], dtype=torch.float32).pin_memory() for _, _, _, e in host]
    eps_dev = [t.cuda() for t in eps_host]
    if world > 1:
        graph_note = "eager launches (NCCL all-reduces between the runs are not captured)"
    elif not args.no_graph and not args.ncu:
        try:
            s_img.copy_(dev_sets[0][0]); s_z.copy_(dev_sets[0][1]); s_al.copy_(dev_sets[0][2])
            graph = model.capture_step(s_img, s_z, s_al, s_eps)
            graph_note = "CUDA graph replay"
        except Exception as e:          # noqa: BLE001
            graph, graph_note = None, f"eager launches (graph capture failed: {type(e).__name__})"
            torch.cuda.synchronize()
    launches_per_step = [None]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            comm.barrier()
            torch.cuda.synchronize()

    def step_resident(k):
        i, z, a, e = dev_sets[k % n_sets]
        if graph is None:
            model.update_model(i, z, a, e)
        else:                            # inputs already in HBM: device-to-device refill of the static buffers
            s_img.copy_(i, non_blocking=True); s_z.copy_(z, non_blocking=True); s_al.copy_(a, non_blocking=True)
            s_eps.copy_(eps_dev[k % n_sets], non_blocking=True)
            graph.replay()

    def step_e2e(k):
        i, z, a, e = host[k % n_sets]
        if graph is None:
            d_img.copy_(i, non_blocking=True)
            d_z.copy_(z, non_blocking=True)
            d_al.copy_(a, non_blocking=True)
            model.update_model(d_img, d_z, d_al, e)
        else:
            s_img.copy_(i, non_blocking=True); s_z.copy_(z, non_blocking=True); s_al.copy_(a, non_blocking=True)
            s_eps.copy_(eps_host[k % n_sets], non_blocking=True)
            graph.replay()
        loss_host.copy_(model.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the user reads the losses every step

    def timed(fn, steps, warmup):
        for k in range(warmup):
            fn(k)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launches
        ev0.record()
        for k in range(steps):
            fn(warmup + k)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = ops.launches - l0
        if graph is not None:            # replayed launches do not pass through ops.*: count what the graph holds
            launches = steps * launches_per_step[0]
        if world > 1:
            t = torch.tensor([ms], device=ops.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches

    # kernels per step, counted by one eager step (the graph replays exactly these)
    l0 = ops.launches
    i0, z0, a0, e0 = dev_sets[0]
    model.update_model(i0, z0, a0, e0)
    torch.cuda.synchronize()
    launches_per_step[0] = ops.launches - l0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if args.ncu:
        for k in range(args.warmup + args.steps):
            step_resident(k)
        torch.cuda.synchronize()
        print(json.dumps({"ncu_run": True, "launches_per_step": launches_per_step[0]}))
        return
    ms_step, launches = timed(step_resident, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, args.steps, 1)
    losses = model.read_losses()
    finite = all(np.isfinite(v) for v in losses.values())

    # ---- roofline of the dominant kernel family (the implicit-GEMM conv kernels), timed live with CUDA events
    roof = None
    if not args.no_profile_pass:
        roof = conv_roofline(model, ops, lambda k: model.update_model(*dev_sets[k % n_sets]), algo)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sb, ssteps = min(B, 64), (4 if not multiclass else 2)     # ~10-20 s of CPU work on 16 cores
        sec = oracle_step_time(sb, multiclass, ssteps, 1, threads)
        cpu = {"value": sb / sec, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{ssteps} update_model steps (after 1 warm-up) at batch {sb} of the same {mode}-class 64x64 "
                         f"workload ({sec:.1f} s per step, torch-CPU fp32 oracle, {threads} threads)"}

    if rank == 0:
        gimg = B * world
        ws_gb = ops.bytes_allocated() / 1e9
        line = {
            "metric": "G+D+GP step images/sec at 64x64", "value": gimg / (ms_step * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tc": "tf32", "tc3x": "3xtf32", "simt": "f32"}[algo], "data": "synthetic",
            "config": {"workload": f"{mode}-class 64x64 EdgeGAN update_model (G1+G2, 3 critics with WGAN-GP, "
                                   f"{'classifier, ' if multiclass else ''}E), batch {B}/GPU",
                       "global_batch": gimg, "parallelism": f"dp{world}", "conv_algo": algo, "launch": graph_note,
                       "l2": f"step working set {ws_gb:.2f} GB >> 126 MB L2 (every activation is rewritten each step); no flush",
                       "gflop_per_image": GFLOP_PER_IMAGE[mode], "losses_finite": finite},
            "e2e": {"value": gimg / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "step_tflops": gimg * GFLOP_PER_IMAGE[mode] / ms_step,      # GFLOP / ms = TFLOP/s (whole job)
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


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
            recs.append((name, used, 2.0 * macs, e0, e1))
        return g

    for pid, n in enumerate(("conv_fwd", "conv_bwd_data", "conv_bwd_weight")):
        setattr(ops, n, wrap(n, pid))
    try:
        step_fn(0)
        torch.cuda.synchronize()
    finally:
        for n, f in orig.items():
            setattr(ops, n, f)
    agg = {}
    for name, used, fl, e0, e1 in recs:
        key = (name, "tc" if used in (2, 3) else "simt")
        a = agg.setdefault(key, [0.0, 0.0, 0])
        a[0] += fl
        a[1] += e0.elapsed_time(e1)
        a[2] += 1
    tc_fl = sum(v[0] for k, v in agg.items() if k[1] == "tc")
    tc_ms = sum(v[1] for k, v in agg.items() if k[1] == "tc")
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
    detail = {f"{k[0]}/{k[1]}": {"launches": v[2], "ms": round(v[1], 3), "tflops": (v[0] / v[1] / 1e9 if v[1] > 0 else None)}
              for k, v in agg.items()}
    if tc_ms > 0:
        ach = tc_fl / tc_ms / 1e9
        mult = 3.0 if algo == "tc3x" else 1.0
        return {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv (fwd/dgrad/wgrad), kind::tf32",
                "achieved": ach, "peak": peak_kind, "unit": "TFLOP/s", "frac": ach / peak_kind, "traffic": None,
                "peak_source": f"{src} bf16 sustained / 2 (tf32 rate)", "conv_ms_per_step": all_ms,
                "mma_per_algorithmic_flop": mult,
                "tensor_pipe_work_frac": mult * ach / peak_kind,
                "note": ("achieved counts ALGORITHMIC conv FLOPs; the 3xTF32 mode issues 3 tensor-core MMAs per algorithmic "
                         "MMA (hi*hi + lo*hi + hi*lo), so the tensor pipe is busy tensor_pipe_work_frac of its tf32 peak"),
                "detail": detail}
    simt_fl = sum(v[0] for v in agg.values())
    ach = simt_fl / all_ms / 1e9 if all_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "fp32 FFMA implicit-GEMM conv (SIMT path; tensor cores not used)",
            "achieved": ach, "peak": peak_kind, "unit": "TFLOP/s", "frac": ach / peak_kind, "traffic": None,
            "peak_source": f"{src} bf16 sustained / 2 (tf32 rate)", "conv_ms_per_step": all_ms, "detail": detail}


if __name__ == "__main__":
    main()
