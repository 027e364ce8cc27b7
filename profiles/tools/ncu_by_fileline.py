"""Executed instructions + stall samples per (source file, line) of one kernel, from an ncu report with --import-source.
usage: ncu_by_fileline.py report.ncu-rep object.o kernel-substring [top] [--opcodes] [--outer] [--regions=name:lo-hi,...]
Unlike ncu_by_line.py it keeps the FILE of every line (inlined headers) and can print the SASS opcode histogram.
--outer attributes an instruction to the OUTERMOST frame of its inline chain (the line of the kernel body) and also
prints issue slots (FFMA2 / FADD2 / FMUL2 counted twice: measured, profiles/tools/ffma2_bench.cu); --regions sums
outer lines of the kernel body's file into named line ranges."""
import csv, sys, subprocess, collections, re, os, tempfile
args = [a for a in sys.argv[1:] if not a.startswith('--')]
rep, obj, ksub = args[:3]
top = int(args[3]) if len(args) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
outer = '--outer' in sys.argv
dis = subprocess.run(['nvdisasm', '-gi' if outer else '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of, op_of = {}, {}
cur = None; infn = False
for l in dis:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: infn = ksub in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue     # with -gi the last line of a chain is the outermost frame
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m: line_of[int(m.group(1), 16)] = cur; op_of[int(m.group(1), 16)] = m.group(2)
kern = [a[len('--kernel='):] for a in sys.argv if a.startswith('--kernel=')]      # ncu -k filter when the report holds several kernels
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + (['-k', kern[0]] if kern else []), capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
hdr = rows[hi]
ai, ci, sm = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = None
smp = collections.Counter(); ins = collections.Counter(); ops = collections.Counter(); fileins = collections.Counter(); slots = collections.Counter()
for r in rows[hi + 1:]:
    try: a = int(r[ai], 16); n = int(r[ci]); s = int(r[sm])
    except Exception: continue
    if base is None: base = a
    k = line_of.get(a - base); smp[k] += s; ins[k] += n
    o = op_of.get(a - base, '?').split('.')[0]
    ops[o] += n
    slots[k] += n * (2 if o in ('FFMA2', 'FADD2', 'FMUL2') else 1)
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
reg = [a for a in sys.argv if a.startswith('--regions=')]
if reg:
    tot_sl = sum(slots.values())
    print(f'--- regions (issue slots, total {tot_sl}; packed f32x2 ops count twice)')
    main_file = collections.Counter({k[0]: 0 for k in slots if k})
    for k, v in slots.items():
        if k: main_file[k[0]] += v
    mf = main_file.most_common(1)[0][0]
    used = 0
    for spec in reg[0].split('=', 1)[1].split(','):
        name, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))
        v = sum(c for k, c in slots.items() if k and k[0] == mf and lo <= k[1] <= hi)
        vi = sum(c for k, c in ins.items() if k and k[0] == mf and lo <= k[1] <= hi)
        used += v
        print(f'  {name:28s} {100 * v / tot_sl:5.1f}% of slots   {100 * vi / tot_i:5.1f}% of instructions')
    print(f'  {"(elsewhere)":28s} {100 * (tot_sl - used) / tot_sl:5.1f}%')
if '--opcodes' in sys.argv:
    print('--- opcodes')
    for o, n in ops.most_common(40): print(f'{100 * n / tot_i:5.1f}%  {o}')
