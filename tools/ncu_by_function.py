"""Aggregate an ncu report's source page by device function of k1_cut.cu (warp instructions and stall samples per
source line, summed over the line range of each function).  Usage: python tools/ncu_by_function.py report.ncu-rep [--rev COMMIT] [--lines]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
src_path = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "sgufp_solver_b200/csrc/k1_cut.cu"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
per_line = {}
embedded = {}                      # the source as imported into the report (--import-source on): line ranges follow IT
cur_file = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or not r[0].isdigit():
        continue
    if not cur_file or not cur_file.endswith(src_path.split("/")[-1]):
        continue
    d = dict(zip(hdr, r))
    embedded[int(r[0])] = r[1]
    try:
        per_line[int(r[0])] = (int(d["Instructions Executed"]), int(d["# Samples"]), int(d["Thread Instructions Executed"]))
    except (ValueError, KeyError):
        pass

# function ranges: a line that starts a __device__/__global__ function or a struct begins a new range
starts = []
pat = re.compile(r"^(?:__device__|__global__|static __device__)[^;]*?\b(?!__launch_bounds__)(\w+)\s*\((?!NW)")
lines = open(src_path).read().split("\n")
if "--rev" in sys.argv:             # the source as of the commit the profiled library was built from
    rev = sys.argv[sys.argv.index("--rev") + 1]
    lines = subprocess.run(["git", "show", "%s:%s" % (rev, src_path)], capture_output=True, text=True, check=True).stdout.split("\n")
bad = sum(1 for k, v in embedded.items() if k > len(lines) or lines[k - 1].strip() != v.strip())
if bad:
    print("WARNING: %d profiled lines differ from the source read here; pass --rev <commit of the profiled build>" % bad)
for i, s in enumerate(lines, 1):
    m = pat.match(s)
    if m:
        starts.append((i, m.group(1)))
starts.append((len(lines) + 1, None))
tot_i = sum(v[0] for v in per_line.values()) or 1
tot_s = sum(v[1] for v in per_line.values()) or 1
print("total warp instructions attributed: %.3f G, samples %d" % (tot_i / 1e9, tot_s))
agg = []
first = starts[0][0]
pre = [v for k, v in per_line.items() if k < first]
if pre:
    agg.append(("(helpers before line %d)" % first, sum(v[0] for v in pre), sum(v[1] for v in pre), sum(v[2] for v in pre)))
for (a, name), (b, _) in zip(starts, starts[1:]):
    vs = [v for k, v in per_line.items() if a <= k < b]
    if vs:
        agg.append((name + " (%d-%d)" % (a, b - 1), sum(v[0] for v in vs), sum(v[1] for v in vs), sum(v[2] for v in vs)))
for name, ni, ns, nt in sorted(agg, key=lambda x: -x[1]):
    print("%-42s inst %6.2f %%  samples %6.2f %%  lanes %4.1f" % (name, 100.0 * ni / tot_i, 100.0 * ns / tot_s, nt / max(ni, 1)))
if "--lines" in sys.argv:
    for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%5d  inst %5.2f %%  samples %5.2f %%  %s" % (k, 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s, lines[k - 1].strip()[:110]))
