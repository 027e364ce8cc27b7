"""Aggregate ncu per-instruction 'Instructions Executed' by CUDA source line (needs -lineinfo build).
usage: ncu_lines.py <report.ncu-rep> <object.o> <kernel-mangled-substring> [top-N] [ncu -k regex, needed when the report holds
several kernels]"""
import csv, sys, subprocess, collections, re, os, tempfile
rep, obj, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur = None; infn = False
for l in dis:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        infn = ksub in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
kflt = ['-k', 'regex:' + sys.argv[5]] if len(sys.argv) > 5 else []
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + kflt, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
hdr = rows[hi]
ai, ci, sm = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = None
agg = collections.Counter(); smp = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    try: a = int(r[ai], 16); n = int(r[ci]); s = int(r[sm])
    except Exception: continue
    if base is None: base = a
    key = line_of.get(a - base)
    agg[key] += n; smp[key] += s; tot += n
srcs = {}
def text(key):
    if key is None: return ''
    f, ln = key
    if f not in srcs:
        for root in ('motion_planning_baselines_b200/csrc',):
            p = os.path.join(root, f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ''
print('total executed warp instr', tot, ' stall samples', sum(smp.values()))
for key, n in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 40):
    print(f'{n / tot * 100:5.1f}% inst {smp[key] / max(1, sum(smp.values())) * 100:5.1f}% smp  {str(key):28s} {text(key)}')
