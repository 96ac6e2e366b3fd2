"""Write profiles/r2_sass_tcgen05.txt: tcgen05 / TMEM / bulk-copy mnemonics per kernel of the built library."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "openwakeword_b200/csrc/libowwb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
KEEP = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "ELECT", "UTCCP", "UTCATOMSWS")
cur, cnt, eg = None, collections.OrderedDict(), {}
for ln in sass:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); eg[cur] = []; continue
    if cur is None:
        continue
    for op in re.findall(r"\b([A-Z][A-Z0-9_]+(?:\.[A-Z0-9_x]+)*)\b", ln.split("*/")[1] if "*/" in ln else ""):
        base = op.split(".")[0]
        if base in KEEP:
            cnt[cur][op] += 1
            if len(eg[cur]) < 2 and base in ("UTCHMMA", "LDTM", "UBLKCP"):
                eg[cur].append(re.sub(r"\s+", " ", ln.strip())[:140])
            break
out = ["# SASS evidence (cuobjdump -sass openwakeword_b200/csrc/libowwb200.so, sm_100a): tcgen05 / TMEM / bulk-copy / mbarrier",
       "# mnemonics per kernel.  UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (the 1-D TMA form),",
       "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, ELECT = elect.sync.  No HMMA (legacy mma.sync) and, by design, no",
       "# UTMALDG (tensor-map TMA): every tile on this path is a set of contiguous runs (DESIGN.md section 3).", ""]
for k, c in cnt.items():
    if not c:
        continue
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    out.append(re.sub(r"\(anonymous namespace\)::", "", name)[:170])
    out.append("    " + "  ".join(f"{op} x{n}" for op, n in sorted(c.items())))
    for l in eg[k]:
        out.append("      e.g. " + l)
open(os.path.join(ROOT, "profiles/r2_sass_tcgen05.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
