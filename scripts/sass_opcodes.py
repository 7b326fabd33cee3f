"""Per-kernel histogram of the SASS opcodes that prove the Blackwell-native path (tcgen05 -> UTC*MMA, TMEM ld/st -> LDTM/STTM,
TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR) from `cuobjdump -sass libopenglue_b200.so`.
usage: python scripts/sass_opcodes.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else 'openglue_b200/libopenglue_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
WANT = ('UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCOMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'UTCCP', 'HMMA', 'HGMMA',
        'SYNCS', 'LDGSTS', 'MUFU.EX2', 'F2FP', 'FFMA', 'BAR.SYNC')
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and kern:
        op = m.group(1)
        hist[kern]['total'] += 1
        for w in WANT:
            if op.startswith(w):
                hist[kern][w + ('.2CTA' if '.2CTA' in op and w.startswith('UTC') else '')] += 1
print(f'# SASS opcode histogram of {lib} (cuobjdump -sass; sm_100a)')
for k, c in hist.items():
    if c['total'] == 0:
        continue
    items = ', '.join(f'{o} {n}' for o, n in sorted(c.items()) if o != 'total')
    print(f'{k[:100]:100s} total {c["total"]:6d} | {items}')
