"""Attribute ncu source-page samples / executed instructions to CUDA source lines.

usage: ncu_lines.py <src.csv from `ncu --page source --csv`> <lib.so> <kernel substring> [top]
The i-th SASS row of the ncu page is matched with the i-th instruction nvdisasm prints for
the kernel (same cubin), whose `//## File "..", line N` markers give the line."""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith('.cubin'):
        continue
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur = None; inside = False; fname = None
    for ln in out.splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            inside = kname in ln and '.text.' in ln
            cur = None
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), 'inlined' in m.group(3))
            continue
        if re.match(r'\s*/\*[0-9a-f]+\*/\s+', ln):
            lines.append((cur, ln.strip()))
    if lines:
        break
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); it = hdr.index('Thread Instructions Executed')
stall_cols = [(h, k) for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
sass = [r for r in rows[2:] if len(r) > isamp and r[0].startswith('0x')]
n = min(len(sass), len(lines))
print(f'ncu rows {len(sass)}, nvdisasm instrs {len(lines)}')
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot = 0
for k in range(n):
    key = lines[k][0]
    s = int(sass[k][isamp] or 0); e = int(sass[k][ie] or 0); t = int(sass[k][it] or 0)
    a = agg[key]; a[0] += s; a[1] += e; a[2] += t
    for h, c in stall_cols:
        v = int(sass[k][c] or 0)
        if v:
            a[3][h[6:]] += v
    tot += s
srcs = {}
def srcline(key):
    if not key: return ''
    f, l, _ = key
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ''
tote = sum(a[1] for a in agg.values())
print(f'total samples {tot}, executed warp-instrs {tote}')
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ' '.join(f'{h}:{v}' for h, v in a[3].most_common(3))
    print(f'{100*a[0]/max(tot,1):5.1f}% smp  {100*a[1]/max(tote,1):5.1f}% ins thr {a[2]/max(a[1],1):4.1f} {key[0] if key else None}:{key[1] if key else 0:4d} | {srcline(key)} | {st}')
