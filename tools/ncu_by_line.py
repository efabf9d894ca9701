"""Per-source-line totals of an ncu report's SASS page (instructions executed, stall samples), joined with the line
table of the cubin the library was built from (nvdisasm -g).  For reports whose embedded source ncu cannot resolve.
usage: ncu_by_line.py <report.ncu-rep> <object.o> <kernel substring> [top N]"""
import csv, re, subprocess, sys, tempfile, os, collections

rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# address -> (file, line) for the wanted function
line_of, cur, infn = {}, None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        infn = kern in l
    elif infn:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
iA, iS, iI, iSmp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) <= iSmp:
        continue
    a = int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA])
    if base is None:
        base = a
    off = a - base
    ln = line_of.get(off, ((None, 0), ""))[0] or ("?", 0)
    ins, smp = int(float(r[iI] or 0)), int(float(r[iSmp] or 0))
    agg[ln][0] += ins; agg[ln][1] += smp; agg[ln][2] += 1
    tot_i += ins; tot_s += smp
print(f"total instructions {tot_i:.4e}  samples {tot_s}")
print("  inst%  smpl%  sass  file:line")
for ln, (i, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f" {100 * i / tot_i:6.2f} {100 * s / max(tot_s, 1):6.2f} {n:5d}  {ln[0]}:{ln[1]}")
# optional: totals per line range of the main file, e.g. REGIONS="vote_grouped:90-232,heavy:283-402"
if os.environ.get("REGIONS"):
    for spec in os.environ["REGIONS"].split(","):
        name, rng = spec.split(":"); a, b = map(int, rng.split("-"))
        i = sum(v[0] for k, v in agg.items() if k[0].endswith("ppf_vote_grouped.cu") and a <= k[1] <= b)
        s = sum(v[1] for k, v in agg.items() if k[0].endswith("ppf_vote_grouped.cu") and a <= k[1] <= b)
        print(f"region {name:14s} inst {i:.4e} ({100 * i / tot_i:5.2f}%)  samples {100 * s / max(tot_s, 1):5.2f}%")
    other = collections.defaultdict(lambda: [0, 0])
    for k, v in agg.items():
        if not k[0].endswith("ppf_vote_grouped.cu"):
            other[k[0]][0] += v[0]; other[k[0]][1] += v[1]
    for k, v in sorted(other.items(), key=lambda kv: -kv[1][0]):
        print(f"file   {k:24s} inst {v[0]:.4e} ({100 * v[0] / tot_i:5.2f}%)  samples {100 * v[1] / max(tot_s, 1):5.2f}%")
