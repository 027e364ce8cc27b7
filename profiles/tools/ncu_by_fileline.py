"""Executed instructions + stall samples per (source file, line) of one kernel, from an ncu report with --import-source.
usage: ncu_by_fileline.py report.ncu-rep object.o kernel-substring [top] [--opcodes]
Unlike ncu_by_line.py it keeps the FILE of every line (inlined headers) and can print the SASS opcode histogram."""
import csv, sys, subprocess, collections, re, os, tempfile
args = [a for a in sys.argv[1:] if not a.startswith('--')]
rep, obj, ksub = args[:3]
top = int(args[3]) if len(args) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of, op_of = {}, {}
cur = None; infn = False
for l in dis:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: infn = ksub in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m: line_of[int(m.group(1), 16)] = cur; op_of[int(m.group(1), 16)] = m.group(2)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
hdr = rows[hi]
ai, ci, sm = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = None
smp = collections.Counter(); ins = collections.Counter(); ops = collections.Counter(); fileins = collections.Counter()
for r in rows[hi + 1:]:
    try: a = int(r[ai], 16); n = int(r[ci]); s = int(r[sm])
    except Exception: continue
    if base is None: base = a
    k = line_of.get(a - base); smp[k] += s; ins[k] += n
    ops[op_of.get(a - base, '?').split('.')[0]] += n
    fileins[k[0] if k else None] += n
tot_i = sum(ins.values()); tot_s = sum(smp.values())
print(f'total warp instructions {tot_i}  samples {tot_s}')
for f, n in fileins.most_common(): print(f'  {f}: {100 * n / tot_i:.1f}%')
srcs = {}
def text(k):
    if not k: return ''
    for root in ('motion_planning_baselines_b200/csrc', '.'):
        p = os.path.join(root, k[0])
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][k[1] - 1].strip()[:90] if k[1] - 1 < len(srcs[p]) else ''
    return ''
for k, v in ins.most_common(top):
    print(f'{100 * v / tot_i:5.1f}% inst {v:9d}  samp {100 * smp[k] / tot_s:4.1f}%  {k[0] if k else None}:{k[1] if k else 0}: {text(k)}')
if '--opcodes' in sys.argv:
    print('--- opcodes')
    for o, n in ops.most_common(40): print(f'{100 * n / tot_i:5.1f}%  {o}')
