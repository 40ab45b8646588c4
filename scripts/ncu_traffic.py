"""profiles/ncu_traffic.json from `ncu --set full` raw pages: per kernel family the measured DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum) per launch, averaged over the captured launches of that family.
usage: python scripts/ncu_traffic.py family=regex:raw.csv [...]   e.g.  rec_bwd_tc=rec_q_bwd:profiles/prof_rec_q_r02_raw.csv"""
import csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def read(path, pattern):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ik, ir, iw, it = (hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum'))
    out = []
    for r in rows[2:]:
        if re.search(pattern, r[ik]):
            out.append((float(r[ir]) * SCALE[units[ir]], float(r[iw]) * SCALE[units[iw]], r[it] + ' ' + units[it], r[ik][:60]))
    return out


def main():
    try:
        cur = json.load(open(OUT))
    except Exception:
        cur = {}
    cur['_comment'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures (profiles/prof_*_raw.csv), '
                       'written by scripts/ncu_traffic.py; bench.py copies the entry of the dominant kernel into roofline.traffic')
    for arg in sys.argv[1:]:
        fam, rest = arg.split('=', 1)
        pattern, path = rest.split(':', 1)
        ls = read(path, pattern)
        if not ls:
            print('no launches match', arg)
            continue
        tot = sum(a + b for a, b, _, _ in ls)
        cur[fam] = {'bytes_per_launch': tot / len(ls), 'launches_captured': len(ls), 'source': os.path.relpath(path, ROOT),
                    'per_launch': [{'kernel': k, 'read': a, 'written': b, 'duration': t} for a, b, t, k in ls]}
        print(fam, '%.1f MB per launch over %d launches' % (tot / len(ls) / 1e6, len(ls)))
    json.dump(cur, open(OUT, 'w'), indent=1)


if __name__ == '__main__':
    main()
