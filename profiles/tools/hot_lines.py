#!/usr/bin/env python
"""Correlate an ncu report's per-instruction warp-stall samples with CUDA source lines.

    python profiles/tools/hot_lines.py <report.ncu-rep> <kernel-mangled-or-substring> [top]

ncu's CSV source page is SASS-only, so the line table comes from `nvdisasm --print-line-info`
on the cubin extracted from rawhash_b200/librawhash_b200.so (built with -lineinfo).  The n-th
instruction of the kernel in nvdisasm order is the n-th row of the ncu page.
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(root, "rawhash_b200", "librawhash_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "rh_host" not in f][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
for l in sass:
    if l.startswith(".text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + re.sub(r"^_Z\d+", "", kern)[:12]], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed")
body = rows[2:]
if len(body) != len(lines):
    print(f"warning: {len(body)} ncu rows vs {len(lines)} disassembled instructions", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, ln in zip(body, lines):
    a = agg[ln]; a[0] += int(r[si] or 0); a[1] += int(r[ii] or 0); a[2] += int(r[ti] or 0)
tot = sum(a[0] for a in agg.values()) or 1
src_cache = {}
def src(ln):
    if ln is None: return ""
    f = os.path.join(root, "rawhash_b200", "csrc", ln[0])
    if f not in src_cache:
        src_cache[f] = open(f).read().splitlines() if os.path.isfile(f) else []
    s = src_cache[f]
    return s[ln[1] - 1].strip()[:110] if 0 < ln[1] <= len(s) else ""
print(f"total samples {tot}")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot:5.1f}%  inst={a[1]:>10}  thr/inst={a[2]/max(a[1],1):5.1f}  {ln[0] if ln else '?'}:{ln[1] if ln else 0}  {src(ln)}")
