"""SASS evidence for profiles/: per kernel of the built library, how many tcgen05 / TMA / TMEM instructions it holds
(UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, REDG / RED = red.global.add).  usage: python tools/sass_excerpt.py [lib.so] [out.md]"""
import collections, re, subprocess, sys

PAT = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "RED", "HMMA", "FFMA"]


def short_name(d):
    """demangled kernel name without the parameter list: 'void ns::k<4, true>(A, B)' -> 'k<4, true>'"""
    d = d.replace("(anonymous namespace)::", "")
    d = re.sub(r"^void\s+", "", d)
    depth, cut = 0, len(d)
    for i, ch in enumerate(d):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return d[:cut]


def main(lib, out):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, counts, order = None, collections.defaultdict(collections.Counter), []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for p in PAT:
                if op.startswith(p):
                    counts[cur][p] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    with open(out, "w") as f:
        f.write(f"# SASS excerpt of `{lib}` (sm_100a), `cuobjdump -sass | grep` counts per kernel\n\n"
                "UTCHMMA = `tcgen05.mma` (kind::tf32), UTMALDG = TMA tensor load, LDTM / STTM = `tcgen05.ld` / `tcgen05.st`,\n"
                "UTCBAR = `tcgen05.commit`, SYNCS = mbarrier operations, RED = `red.global.add`.  Only kernels that hold at\n"
                "least one tcgen05 / TMA instruction are listed in the first table; the second gives the instruction totals of all\n"
                "kernels (FFMA = fp32 FMA).\n\n| kernel | instructions | " + " | ".join(PAT[:9]) + " |\n|---|---|" + "---|" * 9 + "\n")
        for name, d in zip(order, dem):
            c = counts[name]
            if c["UTCHMMA"] + c["UTMALDG"] + c["LDTM"] + c["STTM"] == 0:
                continue
            short = short_name(d)
            f.write(f"| `{short}` | {c['_all']} | " + " | ".join(str(c[p]) for p in PAT[:9]) + " |\n")
        f.write("\n| kernel | instructions | FFMA | RED |\n|---|---|---|---|\n")
        for name, d in zip(order, dem):
            c = counts[name]
            f.write(f"| `{short_name(d)}` | {c['_all']} | {c['FFMA']} | {c['RED']} |\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "edgegan_b200/libedgegan_b200.so",
         sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_sass_excerpt.md")
