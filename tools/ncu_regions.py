"""Region x instruction-class breakdown (executed warp-instructions per warp) for k_step."""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, lib, kname, nwarps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith('.cubin'): continue
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur = None; inside = False; stack = None
    for ln in out.splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            inside = kname in ln; cur = None; continue
        if not inside: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            rest = m.group(3)
            # outermost inlined-at location in dem_step.cu tells the region
            cur = (os.path.basename(m.group(1)), int(m.group(2)), rest); continue
        if re.match(r'\s*/\*[0-9a-f]+\*/\s+', ln): lines.append((cur, ln.strip()))
    if lines: break
rows = list(csv.reader(open(src_csv))); hdr = rows[1]
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
sass = [r for r in rows[2:] if len(r) > isamp and r[0].startswith('0x')]
bounds = eval(open(os.path.join(os.path.dirname(os.path.abspath(lib)), 'dem_step.regions')).read()) if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(lib)), 'dem_step.regions')) else None
def klass(op):
    if op in ('DMUL', 'DADD', 'DFMA', 'DSETP', 'MUFU'): return 'fp64'
    if op in ('LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL', 'LDC', 'LDCU', 'CCTL', 'PREFETCH', 'LD', 'ST'): return 'mem'
    if op in ('BRA', 'BSSY', 'BSYNC', 'CALL', 'RET', 'EXIT', 'WARPSYNC', 'BREAK', 'NOP'): return 'ctl'
    return 'int'
# region by address order: we use markers = the dem_step.cu line of the instruction when file is dem_step.cu,
# else inherit the last seen dem_step.cu line (instructions are mostly contiguous per region)
tab = collections.defaultdict(lambda: collections.Counter()); samp = collections.Counter()
last = 0
import bisect
marks = [int(x) for x in sys.argv[5].split(',')]; names = sys.argv[6].split(',')
for k in range(min(len(sass), len(lines))):
    key = lines[k][0]
    if key and key[0] == 'dem_step.cu': last = key[1]
    reg = names[bisect.bisect_right(marks, last)]
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', sass[k][ia].strip()); op = m.group(2).split('.')[0]
    tab[reg][klass(op)] += int(sass[k][ie] or 0); samp[reg] += int(sass[k][isamp] or 0)
tots = sum(samp.values())
print(f'{"region":12s} {"fp64":>8s} {"int":>8s} {"mem":>8s} {"ctl":>8s} {"total":>8s}  samples%')
for reg in names:
    c = tab[reg]; t = sum(c.values())
    print(f'{reg:12s} {c["fp64"]/nwarps:8.0f} {c["int"]/nwarps:8.0f} {c["mem"]/nwarps:8.0f} {c["ctl"]/nwarps:8.0f} {t/nwarps:8.0f}  {100*samp[reg]/tots:5.1f}')
