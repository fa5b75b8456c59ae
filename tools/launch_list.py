"""ncu launch-list CSV (gpu__time_duration.sum) of `bench.py --ncu --warmup 1 --steps 1` -> markdown table of the
LAST step's launches, aggregated per kernel.  usage: launch_list.py in.csv launches_per_step out.md [title]"""
import collections, csv, re, sys


def main(path, per_step, out, title):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)
        name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        seq.append((name, v))
    step = seq[-per_step:]
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for k, v in step:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tc = sum(t for k, (n, t) in agg.items() if k.startswith("conv_tc_"))
    with open(out, "w") as f:
        f.write(f"# {title}\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --ncu --warmup 1 --steps 1`, last step\n"
                "(eager launches; per-launch times are cold-cache and serialised: compare SHARES, not absolutes)\n\n")
        f.write(f"total {tot/1000:.2f} ms over {len(step)} launches; tcgen05 conv kernels {tc/1000:.2f} ms = {100*tc/tot:.0f} % of the step here\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {100*t/tot:.1f}% |\n")
    print("wrote", out, f"{tot/1000:.2f} ms", len(step))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu launch list")
