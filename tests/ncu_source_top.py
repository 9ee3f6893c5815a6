"""Top stalled SASS instructions from an `ncu --page source --csv` export (helper for profiles/)."""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
    hdr = rows[hi]
    i_src, i_s, i_ex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    data = []
    for n, r in enumerate(rows[hi + 1:]):
        if len(r) <= i_ex:
            continue
        try:
            data.append((int(r[i_s] or 0), int(r[i_ex] or 0), n, r[i_src].strip()))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    print('kernel:', rows[0][1] if rows[0] else '?', '| stall samples', tot, '| warp instructions', sum(d[1] for d in data))
    for s, e, n, src in sorted(data, key=lambda d: -d[0])[:top]:
        print(f"{s:6d} {100 * s / tot:5.1f}%  exec={e:7d}  #{n:4d}  {src[:100]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
