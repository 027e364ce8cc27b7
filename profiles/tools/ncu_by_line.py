"""Stall samples per CUDA source line. usage: ncu_by_line.py report.ncu-rep object.o kernel-substring source.cu [top]"""
import csv, sys, subprocess, collections, re, os, tempfile
rep, obj, ksub, srcf = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of = {}; cur=None; infn=False
for l in dis:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: infn = ksub in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = int(m.group(2)); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
hdr = rows[hi]
ai, ci, sm = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols=[i for i,c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
base=None; smp=collections.Counter(); ins=collections.Counter(); why=collections.defaultdict(collections.Counter)
for r in rows[hi+1:]:
    try: a=int(r[ai],16); n=int(r[ci]); s=int(r[sm])
    except: continue
    if base is None: base=a
    k=line_of.get(a-base); smp[k]+=s; ins[k]+=n
    for i in stall_cols: why[k][hdr[i]]+=int(r[i] or 0)
text=open(srcf).read().splitlines()
tot=sum(smp.values())
print('total samples',tot)
for k,v in smp.most_common(top):
    w=', '.join(f'{n[6:]} {c}' for n,c in why[k].most_common(2))
    print(f"{v:6d} {100*v/tot:5.1f}%  inst {ins[k]:9d}  L{k}: {text[k-1].strip()[:80] if k else ''}   [{w}]")
