"""Warp-stall samples of an ncu capture by CUDA source line and by enclosing function:
   ncu -i X.ncu-rep --page source --csv --print-source=cuda,sass > src.csv ; python scripts/ncu_hot_lines.py src.csv out.txt "note" """
import bisect, collections, csv, re, sys
csv.field_size_limit(1 << 30)
raw, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(open(raw)))
hdr = next(r for r in rows if r and r[0] == "Line No")
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
byline = collections.defaultdict(lambda: collections.Counter())
srcline = {}
for r in rows:
    if len(r) != len(hdr) or r[0] == "Line No" or not r[0].isdigit():
        continue
    ln = int(r[0]); srcline[ln] = r[1]
    for h in stalls:
        v = r[col[h]]
        if v and v != "0":
            byline[ln][h] += int(float(v))
tot = sum(sum(c.values()) for c in byline.values())
src = open("statusswitchingqp.jl_b200/csrc/ssqp_kernel.cuh").read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.search(r"__device__.*?\b([A-Za-z_0-9]+)\s*\(", l)
    if (l.startswith("static __device__") or l.startswith("__device__")) and m:
        funcs.append((i, m.group(1)))
    if l.startswith("__global__"):
        funcs.append((i, "ssqp_solve_kernel"))
starts = [f[0] for f in funcs]
byf, byreason = collections.Counter(), collections.Counter()
for ln, c in byline.items():
    k = bisect.bisect_right(starts, ln) - 1
    byf[funcs[k][1] if k >= 0 else "(header / inline helpers)"] += sum(c.values())
    byreason.update(c)
with open(out, "w") as f:
    f.write("# %s\n# total samples %d\n" % (note, tot))
    f.write("# stall reasons, share of all samples: " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / tot) for k, v in byreason.most_common(10)) + "\n")
    f.write("# by enclosing function: " + "; ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in byf.most_common(14)) + "\n")
    f.write("line  share   top stall reasons (samples)                 | source\n")
    for ln, c in sorted(byline.items(), key=lambda kv: -sum(kv[1].values()))[:40]:
        top = " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in c.most_common(3))
        f.write("%5d %5.1f%%  %-44s | %s\n" % (ln, 100.0 * sum(c.values()) / tot, top, srcline.get(ln, "").strip()[:120]))
print(open(out).read()[:3500])
