"""Per-kernel shares of ONE frame from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python tools/summarize_launches.py gpurun_out/launches.csv [--title "..."] > profiles/rN_launches_summary.txt

A frame starts at a `stem_conv16x4_kernel` launch (the first kernel of the image branch); the window between two
consecutive ones in the middle of the capture is summarised.  ncu times are cold-cache and serialised: compare SHARES."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    title = sys.argv[sys.argv.index('--title') + 1] if '--title' in sys.argv else ''
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            ns = float(r['Metric Value'].replace(',', ''))
            if r.get('Metric Unit') in ('us', 'usecond'):
                ns *= 1e3
            rows.append((r['Kernel Name'], ns))
    starts = [i for i, (k, _) in enumerate(rows) if 'stem_conv' in k]
    if len(starts) < 2:
        raise SystemExit(f'need two frame starts in the capture, found {len(starts)} in {len(rows)} launches')
    m = len(starts) // 2
    a, b = (starts[m - 1], starts[m]) if m >= 1 else (starts[0], starts[1])
    win = rows[a:b]
    tot = sum(t for _, t in win)
    acc, cnt = defaultdict(float), defaultdict(int)
    for k, t in win:
        k = k.split('(')[0][:96]
        acc[k] += t
        cnt[k] += 1
    ours = sum(t for k, t in acc.items() if 'far3d::' in k)
    if title:
        print(title)
    print('(cold-cache, serialised launch times: compare SHARES, not absolutes)')
    print(f'one frame (window rows {a}..{b - 1} of {len(rows)}): {len(win)} launches, {tot / 1e6:.3f} ms summed kernel time')
    print(f'far3d:: kernels {ours / 1e6:.3f} ms ({100 * ours / tot:.1f}%), torch glue kernels {(tot - ours) / 1e6:.3f} ms')
    for k, t in sorted(acc.items(), key=lambda kv: -kv[1])[:40]:
        print(f'  {t / 1e6:7.3f} ms {100 * t / tot:5.1f}% n={cnt[k]:4d}  {k}')


if __name__ == '__main__':
    main()
