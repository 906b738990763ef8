"""Per-source-line summary of one kernel from an ncu report (`--import-source on`, built with -lineinfo).

    python scripts/ncu_regions.py REPORT.ncu-rep KERNEL_REGEX [min_pct]

Prints, per CUDA source line, the share of executed warp instructions, the average active lanes and the share of
stall samples; used to write the region tables in profiles/."""
import csv, subprocess, sys, collections, io

rep, kern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
fname = None
lines = collections.OrderedDict()
hdr = None
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if row[0] in ("Function Name", "Kernel Name") or hdr is None:
        continue
    if row[0] == "":      # SASS row
        continue
    try:
        ln = int(row[0])
    except ValueError:
        continue
    def col(name):
        i = [k for k, h in enumerate(hdr) if h == name][0]
        v = row[i].replace(",", "")
        try:
            return float(v) if v not in ("-", "") else 0.0
        except ValueError:   # a source line whose own quotes broke the CSV row
            return 0.0
    key = (fname, ln)
    e = lines.setdefault(key, [row[1].strip(), 0.0, 0.0, 0.0, 0.0])
    e[1] += col("# Samples"); e[2] += col("Instructions Executed"); e[3] += col("Thread Instructions Executed")
    e[4] += col("L1 Tag Requests Global")
tot_s = sum(e[1] for e in lines.values()) or 1.0
tot_i = sum(e[2] for e in lines.values()) or 1.0
tot_t = sum(e[3] for e in lines.values())
print(f"# kernel {kern}: warp instructions {tot_i:.3e}, thread instructions {tot_t:.3e}, avg lanes {tot_t / tot_i:.2f}, samples {tot_s:.0f}")
print(f"# {'file:line':28s} {'inst%':>6s} {'lanes':>6s} {'stall%':>6s} {'L1tag%':>6s}  source")
tot_l1 = sum(e[4] for e in lines.values()) or 1.0
for (f, ln), e in lines.items():
    ip, sp = 100 * e[2] / tot_i, 100 * e[1] / tot_s
    if ip >= min_pct or sp >= min_pct:
        print(f"{f + ':' + str(ln):30s} {ip:6.2f} {e[3] / max(e[2], 1):6.1f} {sp:6.2f} {100 * e[4] / tot_l1:6.2f}  {e[0][:90]}")
