"""SASS digest of libplaner_b200.so: per kernel, how many tcgen05 / TMA / TMEM instructions the sm_100a code contains.

    python tools/sass_digest.py > profiles/r02_sass_digest.txt

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor load (.IM2COL = im2col
mode), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit -> mbarrier, HMMA = legacy mma.sync, SYNCS = mbarrier ops.
Runs on a CPU-only box (cuobjdump reads the ELF)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'planer_b200', 'libplaner_b200.so')
PAT = ['UTCHMMA.2CTA', 'UTCHMMA', 'UTMALDG.4D.IM2COL', 'UTMALDG.4D', 'UTMALDG.2D', 'LDTM', 'UTCBAR', 'HMMA', 'UTMAPF', 'SYNCS', 'REDG', 'MEMBAR']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    sha = hashlib.sha256(open(LIB, 'rb').read()).hexdigest()
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        per[cur]['_total'] += 1
        for p in PAT:
            if op.startswith(p):
                per[cur][p] += 1
                break
    demangle = subprocess.run(['c++filt'] + list(per), capture_output=True, text=True).stdout.splitlines()
    print('# SASS digest of planer_b200/libplaner_b200.so (sm_100a), sha256 %s' % sha)
    print('# cuobjdump -sass | per-function instruction counts; only kernels with tensor-core / TMA / TMEM instructions listed in full')
    tot = collections.Counter()
    print('%-96s %7s  %s' % ('kernel', 'instrs', ' '.join('%s' % p for p in PAT)))
    for (name, c), dm in zip(per.items(), demangle):
        for p in PAT:
            tot[p] += c[p]
        short = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', dm)
        short = re.sub(r'\(.*', '', short)
        if any(c[p] for p in PAT[:9]):
            print('%-96s %7d  %s' % (short[:96], c['_total'], ' '.join('%*d' % (len(p), c[p]) for p in PAT)))
    print('%-96s %7s  %s' % ('TOTAL (all %d kernels)' % len(per), '', ' '.join('%*d' % (len(p), tot[p]) for p in PAT)))


if __name__ == '__main__':
    main()
