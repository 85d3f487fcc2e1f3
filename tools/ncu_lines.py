"""Map an ncu SASS source-page CSV onto CUDA source lines using nvdisasm line info (dev tool).

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel mangled substring> [top]
"""
import csv, re, subprocess, sys, os, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyloo_b200", "lib", "libpsisloo_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
dis = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    if any(l.startswith(".text.") and kern in l for l in d):
        dis = d
        break
# locate kernel section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines = []  # per instruction: source line
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---------------------") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ci, cs, cst = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
inst = [r for r in rows[hi + 1:] if len(r) > ci and r[ci].isdigit()]
print("sass rows", len(inst), "disasm instrs", len(lines))
agg = collections.defaultdict(lambda: [0, 0])
tot = 0; tots = 0
for r, ln in zip(inst, lines):
    n = int(r[ci]); s = int(r[cs] or 0)
    agg[ln][0] += n; agg[ln][1] += s; tot += n; tots += s
src_cache = {}
def src(ln):
    if ln is None: return "?"
    f, n = ln
    for base in ("pyloo_b200/csrc", "/usr/local/cuda/include", "/usr/local/cuda/include/crt"):
        p = os.path.join(base if base.startswith("/") else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), base), f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p, errors="replace").read().splitlines()
            L = src_cache[p]
            return L[n - 1].strip()[:90] if n - 1 < len(L) else ""
    return ""
print(f"total warp-instructions {tot}, samples {tots}")
for ln, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:>11} {100*n/tot:5.1f}% | stall samples {100*s/max(tots,1):5.1f}% | {ln} | {src(ln)}")

# ---- aggregate by file and by coarse line ranges of b2l_row_kernel.cuh
byfile = collections.defaultdict(int)
for ln, (n, s) in agg.items():
    byfile[ln[0] if ln else "?"] += n
print("\nby file:")
for f, n in sorted(byfile.items(), key=lambda kv: -kv[1]): print(f"  {n:>11} {100*n/tot:5.1f}%  {f}")
if os.environ.get("RANGES"):
    ranges = [tuple(map(int, r.split("-"))) for r in os.environ["RANGES"].split(",")]
    print("\nby line range:")
    for a, b in ranges:
        n = sum(v[0] for ln, v in agg.items() if ln and ln[0] == os.environ.get("RANGE_FILE", "b2l_row_kernel.cuh") and a <= ln[1] <= b)
        s = sum(v[1] for ln, v in agg.items() if ln and ln[0] == os.environ.get("RANGE_FILE", "b2l_row_kernel.cuh") and a <= ln[1] <= b)
        print(f"  {a:>4}-{b:<4} {n:>11} {100*n/tot:5.1f}%  stall {100*s/max(tots,1):5.1f}%")
