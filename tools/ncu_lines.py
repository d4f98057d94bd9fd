#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel: zips the SASS rows of an ncu report (source page)
with the line table of nvdisasm for the same cubin.  usage: ncu_lines.py <rep> <kernel regex> <mangled name> [top]"""
import csv, re, subprocess, sys, os, collections, tempfile
rep, kre, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "nano-kazen_b200/csrc/libkzgpu.so")], cwd=td, capture_output=True)
cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cubin)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(sass) if l.startswith(".text." + mangled + ":")][0]
lines = []; cur = ("?", 0)
for l in sass[start + 1:]:
    if l.startswith("//---") or l.startswith("\t.section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append((cur, l.split("*/", 1)[1].strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
# the report may hold several launches: take the LAST one (incoherent batch) unless KZ_LAUNCH is set
# kernel-name filters match the base name only; template instances are told apart by the substring KZ_NAME_HAS
names = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
has = os.environ.get("KZ_NAME_HAS", "")
blocks = [i + 1 for i in names if has in rows[i][1] and i + 1 < len(rows) and rows[i + 1] and rows[i + 1][0] == "Address"]
which = int(os.environ.get("KZ_LAUNCH", len(blocks) - 1))
b = blocks[which]; nxt = [i for i in names if i > b]; e = nxt[0] if nxt else len(rows)
hdr = rows[b]; body = [r for r in rows[b + 1:e] if len(r) == len(hdr)]
ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
assert len(body) == len(lines), (len(body), len(lines))
agg = collections.OrderedDict(); tot = 0; tott = 0; tots = 0
for (ln, txt), r in zip(lines, body):
    a = agg.setdefault(ln, [0, 0, 0]); a[0] += int(r[ci]); a[1] += int(r[ti]); a[2] += int(r[si]); tot += int(r[ci]); tott += int(r[ti]); tots += int(r[si])
print(f"launch {which}: {tot} warp instructions, {tott/tot:.1f} threads/inst, {tots} samples")
src = {}
skey = 2 if os.environ.get("KZ_SORT") == "smp" else 0
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][skey])[:top]:
    if f not in src:
        try: src[f] = open(os.path.join(root, "nano-kazen_b200/csrc", f)).read().splitlines()
        except Exception: src[f] = []
    text = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    print(f"{a[0]/tot:6.2%} inst  {a[2]/max(tots,1):6.2%} smp  thr/inst {a[1]/max(a[0],1):5.1f}  {f}:{n}: {text}")
