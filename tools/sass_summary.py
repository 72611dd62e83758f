"""Per-kernel SASS evidence of the Blackwell-native paths: counts of UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG / UBLKCP (TMA tensor / bulk copies), LDGSTS (cp.async) and legacy HMMA in libdqn_b200.so.
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'deepqnetwork_b200', 'libdqn_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pats = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'HMMA', 'SYNCS', 'UTCBAR']
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur:
        for p in pats:
            if re.search(r'\b' + p + r'\b|\b' + p + r'\.', line):
                counts[cur][p] += 1
dem = subprocess.run(['cu++filt'] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
names = dict(zip(counts, dem)) if len(dem) == len(counts) else {k: k for k in counts}
tot = collections.Counter()
print('# cuobjdump -sass deepqnetwork_b200/libdqn_b200.so (sm_100a), instruction counts per kernel')
print('# %-88s %s' % ('kernel', ' '.join('%8s' % p for p in pats)))
for k, c in counts.items():
    tot.update(c)
    if not any(c[p] for p in pats):
        continue
    n = names[k].replace('void ', '').replace('(int)', '').replace('(bool)', '')
    n = n[:n.rfind('>(') + 1] if '>(' in n else re.sub(r'\(.*', '', n)
    print('%-90s %s' % (n[:90], ' '.join('%8d' % c[p] for p in pats)))
print('%-90s %s' % ('TOTAL (%d kernels)' % len(counts), ' '.join('%8d' % tot[p] for p in pats)))
