"""Phase table of the forward kernel from an ncu report (`--import-source on`, -lineinfo): share of issued warp
instructions, average active lanes and share of stall samples per phase of trace_fwd.cu (the phases are the source
ranges between its `// ---- <name>` banner comments) and per inlined helper of trace_common.cuh.

    python scripts/ncu_phases.py REPORT.ncu-rep [path/to/trace_fwd.cu [path/to/trace_common.cuh]]
"""
import bisect, collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
fwd = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "irgs_b200", "csrc", "trace_fwd.cu")
com = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "irgs_b200", "csrc", "trace_common.cuh")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:trace_forward"], capture_output=True, text=True).stdout
fname = hdr = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or row[0] in ("Function Name", "Kernel Name") or row[0] == "":
        continue
    try:
        ln = int(row[0])
    except ValueError:
        continue

    def col(n):
        i = [k for k, h in enumerate(hdr) if h == n][0]
        try:
            return float(row[i].replace(",", ""))
        except ValueError:
            return 0.0
    a = agg[(fname, ln)]
    a[0] += col("Instructions Executed"); a[1] += col("Thread Instructions Executed"); a[2] += col("# Samples")
tot = sum(a[0] for a in agg.values()) or 1.0
tots = sum(a[2] for a in agg.values()) or 1.0
marks, names = [], []
for i, l in enumerate(open(fwd), 1):
    m = re.match(r"\s*// -{40,} (.*)", l)
    if m:
        marks.append(i); names.append(m.group(1).strip())
ph = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (f, ln), a in agg.items():
    if f == "trace_fwd.cu":
        i = bisect.bisect_right(marks, ln) - 1
        key = names[i] if i >= 0 else "kernel prologue"
    elif f == "trace_common.cuh":
        continue
    else:
        key = "inlined: " + f
    for k in range(3):
        ph[key][k] += a[k]
# helpers of trace_common.cuh: ranges between `__device__` function heads
heads = [(i, re.search(r"(\w+)\(", l).group(1)) for i, l in enumerate(open(com), 1)
         if l.startswith("__device__") and re.search(r"(\w+)\(", l)]
for (f, ln), a in agg.items():
    if f == "trace_common.cuh":
        i = bisect.bisect_right([h[0] for h in heads], ln) - 1
        key = "helper: " + (heads[i][1] if i >= 0 else "?")
        for k in range(3):
            ph[key][k] += a[k]
print(f"# forward kernel: {tot:.3e} warp instructions issued, {sum(a[1] for a in agg.values()) / tot:.1f} lanes on average")
print(f"# {'phase':62s} {'inst%':>6s} {'lanes':>6s} {'stall%':>6s}")
for k, a in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    if a[0] / tot >= 0.002:
        print(f"{k[:62]:62s} {100 * a[0] / tot:6.1f} {a[1] / max(a[0], 1):6.1f} {100 * a[2] / tots:6.1f}")
