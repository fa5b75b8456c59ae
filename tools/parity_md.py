"""gpurun_out/parity_report.jsonl (written by tests/test_step_gpu.py on the GPU box) -> profiles/r02_parity.md:
per test and per optimizer run, how many gradient tensors met the plain tolerance, how many passed only through the
noise clause (tests/parity_util.check_grads), the worst tensor with the fp32 oracle's own noise beside it, and the
weight error the run's RMSProp left.  usage: python tools/parity_md.py [in.jsonl] [out.md]"""
import json
import sys


def main(src, dst):
    recs = [json.loads(l) for l in open(src) if l.strip()]
    latest = {}
    for r in recs:                       # keep the last record per test name
        latest[r["test"]] = r
    with open(dst, "w") as f:
        f.write("# Step parity on the B200 (teacher-forced per-run check, `tests/test_step_gpu.py`)\n\n"
                "Each run of `update_model` on the device is replayed by the fp64 oracle (truth) and the fp32 oracle (the\n"
                "reference's precision) from the device's own pre-run weights.  `tol` = plain bar on max|g_dev - g_fp64| / max|g_fp64|\n"
                "per tensor; a tensor above it passes only if it is within 10x max(fp32-oracle distance to fp64, fp64 response to a\n"
                "1e-5 relative input perturbation) -- the column `via noise clause` counts those; in the large / realistic-input tests a\n"
                "tensor may also pass within 10x the largest such instability of its RUN (`via run-level clause`: activation-mask flips, see\n"
                "tests/parity_util.check_grads), and a gradient whose absolute error cannot move its weight by more than 0.005 lr through\n"
                "RMSProp counts as `below absolute bar` (the scalar prelu leaks: heavily cancelling sums).  Weight errors are in units of\n"
                "the learning rate (one RMSProp step moves a weight by <= ~3.2 lr).\n\n")
        for name, r in latest.items():
            f.write(f"## {name}\n\n| run | tol | grad tensors | within tol | via noise clause | via run-level clause | below absolute bar | worst rel. error (tensor; fp32-oracle noise) | "
                    "loss device / fp64 oracle | worst weight error (lr-units; fp32-oracle noise) | weights via noise clause |\n"
                    "|---|---|---|---|---|---|---|---|---|---|---|\n")
            for run, s in r["runs"].items():
                w = s.get("weights", {})
                f.write(f"| {run} | {s.get('grad_tol'):g} | {s.get('tensors')} | {s.get('strict')} | {s.get('noise_clause')} | {s.get('run_clause', 0)} | {s.get('abs_clause', 0)} | "
                        f"{s.get('worst', 0):.2e} ({s.get('worst_name')}; {s.get('worst_noise', 0):.2e}) | "
                        f"{s.get('loss_dev'):.6f} / {s.get('loss_ref'):.6f} | {w.get('worst', 0):.3f} ({w.get('worst_name')}; "
                        f"{w.get('worst_noise', 0):.3f}) | {w.get('noise_clause', 0)} |\n")
            if r.get("extra"):
                f.write("\nEnd state of the whole step vs the fp64 oracle's own whole step, per network (lr-units; drift of the fp32 "
                        "oracle's whole step beside it):\n\n| network | worst weight error | tensor | fp32 oracle drift |\n|---|---|---|---|\n")
                for net, e in r["extra"].items():
                    f.write(f"| {net} | {e['worst']:.3f} | {e['worst_name']} | {e['fp32_oracle_noise']:.3f} |\n")
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_report.jsonl",
         sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_parity.md")
