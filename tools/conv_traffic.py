"""DRAM traffic of the image-branch convolutions of ONE frame from an ncu capture of every conv_persistent launch of an eager run:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_persistent \
        --csv --log-file conv_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-adaptive --eager
    python tools/conv_traffic.py conv_traffic.csv [--mode 2] > profiles/rN_conv_traffic_one_frame.txt

The image branch runs MODE 2 (fp16mx) or MODE 1 (fp16x3) kernels; the decoder's linears use MODE 1 of the same kernel.  A frame's
image branch is a run of `--per-frame` (132) consecutive launches of the image-branch mode; the LAST complete run is summarised."""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    mode = sys.argv[sys.argv.index('--mode') + 1] if '--mode' in sys.argv else '2'
    per_frame = int(sys.argv[sys.argv.index('--per-frame') + 1]) if '--per-frame' in sys.argv else 132
    lines = [l for l in open(path, newline='') if l.startswith('"')]
    launches = OrderedDict()
    for r in csv.DictReader(lines):
        i = int(r['ID'])
        d = launches.setdefault(i, dict(name=r['Kernel Name'].split('(')[0], grid=r.get('Grid Size', '')))
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', '')
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3,
                 'msecond': 1e3}.get(unit, 1)
        d[r['Metric Name']] = v * scale
    rows = list(launches.values())
    tag = f'conv_persistent_kernel<{mode},'
    runs, cur = [], []
    for r in rows:
        if tag in r['name']:
            cur.append(r)
        else:
            if cur:
                runs.append(cur)
            cur = []
    if cur:
        runs.append(cur)
    # the image branch of a frame = the longest runs; the head's linears break them
    frames = [r for r in runs if len(r) >= per_frame - 30]
    if not frames:
        raise SystemExit(f'no run of ~{per_frame} consecutive {tag} launches (runs: {[len(r) for r in runs][:20]})')
    fr = frames[-1]
    rd = sum(r.get('dram__bytes_read.sum', 0) for r in fr)
    wr = sum(r.get('dram__bytes_write.sum', 0) for r in fr)
    t = sum(r.get('gpu__time_duration.sum', 0) for r in fr)
    print(f'source: {path}; mode {mode}; frames found in the capture: {len(frames)} (launches per run {[len(r) for r in frames]}); last one summarised')
    print(f'image-branch convolutions ({len(fr)} launches): DRAM read {rd / 1e9:.3f} GB, written {wr / 1e9:.3f} GB, total {(rd + wr) / 1e9:.3f} GB; '
          f'{(rd + wr) / len(fr) / 1e6:.2f} MB per launch; summed kernel time {t / 1e3:.2f} ms (serialised, cold caches: compare bytes, not times)')
    print()
    print('  #  kernel<mode,halo,cta_group>              grid        time us   read MB  write MB')
    for i, r in enumerate(fr):
        print(f'{i + 1:3d}  {r["name"][-36:]:38s} {r["grid"]:14s} {r.get("gpu__time_duration.sum", 0):8.1f} {r.get("dram__bytes_read.sum", 0) / 1e6:9.2f} '
              f'{r.get("dram__bytes_write.sum", 0) / 1e6:9.2f}')


if __name__ == '__main__':
    main()
