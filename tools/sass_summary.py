"""SASS opcode summary of the library's kernels (no GPU needed): `cuobjdump -sass` of libcontrack_b200.so, per kernel the count
of the instructions that characterise it -- bulk-copy engine (UBLKCP = cp.async.bulk), mbarrier (SYNCS), TMA tensor loads
(UTMALDG) and tcgen05 (UTC*MMA), none of which a 1-D row stream without a contraction needs -- plus atomics, votes, barriers.
usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'contrack_b200', 'lib', 'libcontrack_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pats = {'UBLKCP': r'\bUBLKCP', 'SYNCS(mbarrier)': r'\bSYNCS', 'UTMALDG': r'\bUTMALDG', 'UTC*MMA': r'\bUTC\w*MMA', 'VOTE': r'\bVOTE', 'SHFL': r'\bSHFL',
        'ATOM/RED.global': r'\b(ATOMG|ATOM|REDG|RED)\b|\bATOM\.|\bRED\.', 'ATOMS(shared)': r'\bATOMS', 'BAR': r'\bBAR\.', 'LDG': r'\bLDG', 'STG': r'\bSTG',
        'LDS': r'\bLDS', 'STS': r'\bSTS', 'DFMA/DMUL/DADD': r'\b(DFMA|DMUL|DADD)\b', 'MEMBAR/FENCE': r'\b(MEMBAR|FENCE)', 'STRONG.SYS (peer flags)': r'STRONG\.SYS', 'STRONG.GPU (chains)': r'STRONG\.GPU'}
arch = re.findall(r'arch = (sm_\w+)', out)
print('library:', os.path.relpath(so, ROOT), ' arch of every embedded cubin:', sorted(set(arch)))
kern = None
tab = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', kern)
        kern = re.sub(r'^void ', '', kern)
        kern = re.sub(r'\(.*', '', kern)
        n = 2
        base = kern
        while kern in tab:
            kern = '%s #%d' % (base, n); n += 1
        tab[kern] = collections.Counter()
        continue
    if kern and re.search(r'/\*[0-9a-f]{4,}\*/', line):
        tab[kern]['instructions'] += 1
        for k, p in pats.items():
            if re.search(p, line):
                tab[kern][k] += 1
cols = ['instructions'] + list(pats)
print('%-58s' % 'kernel' + ''.join('%9s' % c[:8] for c in cols))
tot = collections.Counter()
for k, c in tab.items():
    print('%-58s' % k[:57] + ''.join('%9d' % c[x] for x in cols))
    tot.update(c)
print('%-58s' % ('TOTAL (%d kernels)' % len(tab)) + ''.join('%9d' % tot[x] for x in cols))
