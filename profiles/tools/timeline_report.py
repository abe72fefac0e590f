"""Summarise the per-rank timelines written by `bench.py --timeline PATH` (one RK4 step: start / end of every
kernel, peer copy and cross-GPU barrier on the device clock).

    python profiles/tools/timeline_report.py PATH          # reads PATH.rank*.json

Per rank: span of the step, busy time of the plan stream (union of kernel intervals), time inside barriers, busy
time of the copy streams, how much of the copy time lies underneath kernels, and the largest idle gaps of the plan
stream with what ran before / after them."""
import glob
import json
import sys


def union(iv):
    iv = sorted(iv)
    tot, cur0, cur1 = 0.0, None, None
    for a, b in iv:
        if cur1 is None or a > cur1:
            if cur1 is not None:
                tot += cur1 - cur0
            cur0, cur1 = a, b
        else:
            cur1 = max(cur1, b)
    if cur1 is not None:
        tot += cur1 - cur0
    return tot


def overlap(iv1, iv2):
    """length of (union iv1) intersected with (union iv2)"""
    return union(iv1) + union(iv2) - union(list(iv1) + list(iv2))


def main(path):
    for f in sorted(glob.glob(path + '.rank*.json')):
        rows = json.load(open(f))
        k = [(r['t0_ms'], r['t1_ms'], r['what']) for r in rows if not r['what'].startswith('copy') and r['what'] != 'barrier']
        b = [(r['t0_ms'], r['t1_ms']) for r in rows if r['what'] == 'barrier']
        c = [(r['t0_ms'], r['t1_ms']) for r in rows if r['what'].startswith('copy')]
        cb = sum(r['bytes'] for r in rows if r['what'].startswith('copy'))
        t0 = min(r['t0_ms'] for r in rows)
        t1 = max(r['t1_ms'] for r in rows)
        kiv = [(x, y) for x, y, _ in k]
        print('%s: span %.3f ms | kernels busy %.3f | barriers %.3f | copies busy %.3f (%.0f GB/s) of which under kernels %.3f'
              % (f, t1 - t0, union(kiv), union(b), union(c), cb*1e-6/max(union(c), 1e-9), overlap(kiv, c)))
        ev = sorted(k + [(x, y, 'barrier') for x, y in b])
        gaps = sorted(((ev[i+1][0] - ev[i][1], ev[i][2], ev[i+1][2], ev[i][1]) for i in range(len(ev)-1)), reverse=True)[:6]
        for g, before, after, at in gaps:
            if g > 0.005:
                print('    idle %.3f ms at t=%.3f between %s and %s' % (g, at - t0, before, after))


if __name__ == '__main__':
    main(sys.argv[1])
