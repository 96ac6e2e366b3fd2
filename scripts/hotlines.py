"""Fold an ncu --set full report's SASS-level stall samples / executed instructions onto CUDA source lines.
usage: hotlines.py <report.ncu-rep> <kernel-substring> <cubin-name-substring> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kern, cub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "openwakeword_b200/csrc/libowwb200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if cub in f and f.endswith(".cubin") and "-" not in f.split(".")[0]][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
ins, cur, on = [], None, False
for ln in dis:
    if ln.startswith(".text."):
        on = kern in ln
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*;", ln):
        ins.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {h: i for i, h in enumerate(rows[hi])}
data = [r for r in rows[hi + 1:] if len(r) > 5]
byl, st = collections.Counter(), collections.Counter()
for k in range(min(len(ins), len(data))):
    byl[ins[k]] += int(data[k][col["Instructions Executed"]]); st[ins[k]] += int(data[k][col["Warp Stall Sampling (All Samples)"]])
tot, tots = sum(byl.values()), sum(st.values())
cache = {}
def line(f, l):
    if f not in cache:
        p = os.path.join(ROOT, "openwakeword_b200/csrc", f)
        cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return cache[f][l - 1].strip()[:100] if 0 < l <= len(cache[f]) else ""
print(f"{kern}: {len(ins)} SASS instructions matched, warp-inst {tot}, stall samples {tots}")
for key, v in sorted(st.items(), key=lambda kv: -kv[1])[:top]:
    f, l = key if key else ("?", 0)
    print(f"{100*v/max(tots,1):5.1f}% stall {100*byl[key]/max(tot,1):5.1f}% inst  {f}:{l}  {line(f, l)}")
