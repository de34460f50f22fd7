"""Summarise an ncu report of tb_render_kernel: executed SASS opcode mix per tile and hottest source lines.
usage: python tools/ncu_mix.py report.ncu-rep <voice-samples per launch> [cubin-from-current-so]"""
import csv, re, subprocess, sys, os, tempfile
from collections import Counter, defaultdict
rep, vs = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ix = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
data = rows[2:]
tiles = vs / 256
c = Counter(); tot = 0
for r in data:
    n = int(r[ix]); t = r[1].strip().split()
    op = t[1] if t[0].startswith('@') else t[0]
    parts = op.split('.')
    key = parts[0] + ('.' + parts[1] if len(parts) > 1 and parts[0] in ('IMAD', 'LDS', 'STS', 'SHFL', 'F2F', 'I2F', 'F2I', 'ISETP', 'SHF') else '')
    c[key] += n; tot += n
print(f"total warp-instructions per tile: {tot / tiles:.1f}  ({tot / vs:.2f} per voice-sample)")
for op, n in c.most_common(32):
    print(f"  {op:12s} {n / tot * 100:5.1f}%  {n / tiles:7.1f}/tile")
# line attribution
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tuun_b200", "libtuun_b200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.startswith("render.")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur = None; infunc = False; a2l = {}
for ln in dis.splitlines():
    if ln.startswith('.text.'): infunc = 'tb_render_kernel' in ln
    if not infunc: continue
    m = re.search(r'//## File ".*?render\.cu", line (\d+)', ln)
    if m: cur = int(m.group(1)); continue
    m2 = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m2: a2l[int(m2.group(1), 16)] = cur
base = int(data[0][0], 16)
by = defaultdict(lambda: [0, 0])
for r in data:
    l = a2l.get(int(r[0], 16) - base)
    by[l][0] += int(r[ix]); by[l][1] += int(r[isamp])
text = open(os.path.join(os.path.dirname(so), "csrc", "render.cu")).read().split('\n')
tots = sum(v[1] for v in by.values())
print("hottest source lines (inst%, stall-sample%):")
for l, (n, sm) in sorted(by.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {n / tot * 100:5.1f}% {sm / max(1, tots) * 100:5.1f}%  {n / tiles:6.1f}/tile  L{l}: {text[l - 1].strip()[:95] if l else ''}")
