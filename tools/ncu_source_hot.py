#!/usr/bin/env python
"""Hot CUDA source lines of one `ncu --set full --import-source on` capture.
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X_cs.csv; python tools/ncu_source_hot.py X_cs.csv [N]"""
import csv
import sys

STALLS = ('stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_lg', 'stall_no_inst',
          'stall_selected', 'stall_not_selected', 'stall_branch_resolving', 'stall_dispatch', 'stall_membar', 'stall_sleep')


def main(path, top_n):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = {}
    tot = 0
    for r in rows:
        if len(r) > 10 and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or r[0] == '':
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            s = int(d['Warp Stall Sampling (All Samples)'] or 0)
        except ValueError:
            continue
        a = agg.setdefault((ln, r[1][:100]), {'s': 0})
        a['s'] += s
        for k in STALLS:
            try:
                a[k] = a.get(k, 0) + int(d.get(k, '0') or 0)
            except ValueError:
                pass
        tot += s
    print('total samples', tot)
    allst = {k: sum(a.get(k, 0) for a in agg.values()) for k in STALLS}
    print('stall mix: ' + ' '.join(f"{k[6:]}={100 * v / tot:.1f}%" for k, v in sorted(allst.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
    for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1]['s'])[:top_n]:
        st = {k: v for k, v in a.items() if k != 's' and v > 0.15 * max(1, a['s'])}
        print(f"{ln:5d} {100 * a['s'] / tot:5.2f}% {src[:92]:92s} {' '.join(k[6:] + '=' + str(round(100 * v / a['s'])) for k, v in st.items())}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50)
