"""Static SASS instruction count per CUDA source line of one kernel (nvdisasm -g)."""
import collections, os, re, subprocess, sys
obj, kname = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['nvdisasm', '-g', '-c', obj], capture_output=True, text=True).stdout
cnt = collections.Counter(); inside = False; cur = None; n = 0
for ln in out.splitlines():
    if ln.startswith('//---') and '.text.' in ln:
        inside = kname in ln; cur = None; continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]+\*/\s+', ln): cnt[cur] += 1; n += 1
print('static instrs', n)
for k, v in cnt.most_common(top): print(v, k)
