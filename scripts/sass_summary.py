"""Per-kernel SASS opcode counts of libtc_b200.so (cuobjdump -sass): the instruction classes that show what each kernel
is built from -- DMMA (FP64 tensor pipe, mma.sync.m8n8k4.f64), DFMA/DADD/DMUL (FP64 pipe), UBLKCP (cp.async.bulk = TMA
bulk copies), SYNCS (mbarrier), BAR, LDS/STS, LDG/STG, SHFL, MUFU.RSQ64H, local-memory spills (LDL/STL).
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'time_crystal_tensor_network_b200', 'libtc_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cols = ['DMMA', 'DFMA', 'DADD', 'DMUL', 'MUFU.RSQ64H', 'UBLKCP', 'SYNCS', 'BAR', 'LDS', 'STS', 'LDG', 'STG', 'SHFL', 'ATOM', 'LDL', 'STL', 'UTMALDG', 'UTCHMMA', 'HMMA']
kern, counts = None, {}
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(TcDev.*', '', kern).replace('void ', '')
        counts[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m and kern:
        op = m.group(1)
        counts[kern]['total'] += 1
        for c in cols:
            if op == c or op.startswith(c + '.') or (c == 'ATOM' and op.startswith(('ATOM', 'RED'))):
                counts[kern][c] += 1
print('SASS opcode counts per kernel (static instruction counts, cuobjdump -sass %s)' % os.path.relpath(lib, ROOT))
print('UTMALDG / UTCHMMA / HMMA are 0 everywhere: the path is FP64 (tcgen05 has no f64 kind; DMMA is the Blackwell FP64 tensor path),')
print('bulk copies are the 1-D cp.async.bulk form (UBLKCP), not tensor-map TMA.\n')
print('%-58s' % 'kernel' + ''.join('%8s' % c[:7] for c in ['total'] + cols[:16]))
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]['total']):
    print('%-58s' % k[:57] + ''.join('%8d' % c[x] for x in ['total'] + cols[:16]))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print('%-58s' % 'whole library' + ''.join('%8d' % tot[x] for x in ['total'] + cols[:16]))
print('\nUTMALDG %d, UTCHMMA %d, HMMA %d' % (tot['UTMALDG'], tot['UTCHMMA'], tot['HMMA']))
