"""Per-kernel counts of the Blackwell-specific SASS instructions in libssasr.so (cuobjdump -sass):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy
(cp.async.bulk, incl. multicast), UTCBAR = tcgen05.commit.  usage: python scripts/sass_summary.py > profiles/sass_summary_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'ss_asr_b200', 'libssasr.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pats = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'UCGABAR']
cnt = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r'\b' + p + r'\b|\b' + p + r'\.', line):
            cnt[cur][p] += 1
            if p == 'UBLKCP' and 'MULTICAST' in line:
                cnt[cur]['UBLKCP.MULTICAST'] += 1
demangle = subprocess.run(['c++filt'], input='\n'.join(cnt.keys()), capture_output=True, text=True).stdout.splitlines()
print('SASS summary of ss_asr_b200/libssasr.so (sm_100a), %d kernels; only kernels using tensor-core / TMA / bulk-copy instructions are listed' % len(cnt))
cols = pats[:9] + ['UBLKCP.MULTICAST']
print('%-72s ' % 'kernel' + ' '.join('%8s' % c[:8] for c in cols))
tot = collections.Counter()
for (k, c), name in zip(cnt.items(), demangle):
    if not any(c[p] for p in cols):
        continue
    short = re.sub(r'\(.*', '', name.replace('(anonymous namespace)::', '')).replace('ssasr::', '').replace('void ', '')
    print('%-72s ' % short[:72] + ' '.join('%8d' % c[p] for p in cols))
    tot.update(c)
print('%-72s ' % 'TOTAL' + ' '.join('%8d' % tot[p] for p in cols))
