"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/ (run here, no GPU needed)."""
import csv, io, subprocess, sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of sustained-active peak)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_bytes.sum", "L2 bytes (all traffic)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory LSU wavefronts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) CTAs/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main(path, out):
    if path.endswith(".csv"):            # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
        raw = open(path, errors="replace").read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{path}`\n\n")
        for r in data:
            f.write(f"## {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m, label in METRICS:
                if m in col:
                    f.write(f"| {label} (`{m}`) | {r[col[m]]} | {units[col[m]]} |\n")
            f.write("\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
