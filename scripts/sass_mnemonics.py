#!/usr/bin/env python
"""Counts of the SASS mnemonics that prove TMA / tcgen05 / TMEM use, per kernel of maxsim_tc.cu.
    cuobjdump -sass hybrid-rag-colbertv2_b200/csrc/build/maxsim_tc.o | python scripts/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections
import re
import sys

KEYS = ['UTMALDG', 'UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTCATOMSWS', 'SYNCS', 'UCGABAR', 'FMNMX3', 'FMNMX', 'FSEL', 'SHFL', 'CREDUX',
        'STG', 'LDG', 'ATOMG', 'RED', 'VOTE', 'BAR']
cur, counts = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        t = re.search(r'maxsim_tc_kernelILi(\d)ELi(\d)ELi(\d)ELb(\d)ELb(\d)E', m.group(1))
        u = re.search(r'maxsim_dm_kernelILb(\d)E', m.group(1))
        cur = (('maxsim_tc_kernel<MT=%s,ZP=%s,CG=%s,TK=%s,RR=%s>' % t.groups()) if t else
               (('maxsim_dm_kernel<TK=%s>' % u.groups()) if u else m.group(1)[:60]))
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + '.'):
                counts[cur][k] += 1
                break
print("SASS mnemonic counts per kernel of hybrid-rag-colbertv2_b200/csrc/maxsim_tc.cu (cuobjdump -sass, sm_100a).")
print("UTMALDG = TMA tensor load (.2CTA forms in the CTA-pair kernel), UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit,")
print("LDTM / STTM = tcgen05.ld / st, UTCATOMSWS = TMEM allocation, SYNCS = mbarrier, UCGABAR = cluster barrier,")
print("FMNMX3 = 3-input max (the epilogue's tree).\n")
for k, c in counts.items():
    print(k)
    print("   " + "  ".join(f"{n}={c[n]}" for n in KEYS if c[n]))
