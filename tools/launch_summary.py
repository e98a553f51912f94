"""Summarise the ncu launch list of tools/gpu_profile.sh (gpu__time_duration.sum per launch) into profiles/:
    python tools/launch_summary.py gpurun_out/launches.csv r01b"""
import csv, sys
from collections import OrderedDict

src, tag = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith('==')]
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
launches = []
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    v = float(r[iv].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[iu], 1.0)
    launches.append((r[ik].split('(')[0].split('::')[-1], v))
agg = OrderedDict()
for k, v in launches:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(v for _, v in launches)
open('profiles/%s_launches.csv' % tag, 'w').write(open(src).read())
with open('profiles/%s_launches_summary.md' % tag, 'w') as f:
    f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none) -- %s\n\n' % tag)
    f.write('Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 '
            '--no-cpu-baseline --min-warm-sec 0` (tools/gpu_profile.sh).  ResNet-18 fp16 batch 128; the first 400 launches = the one-off '
            'weight cast / packing kernels of the executor build plus the first forwards.  Times are cold-cache and serialised: '
            'compare shares.\n\n| kernel | launches | total us | share | us / launch |\n|---|---|---|---|---|\n')
    for k, (n, v) in sorted(agg.items(), key=lambda t: -t[1][1]):
        f.write('| `%s` | %d | %.1f | %.1f%% | %.1f |\n' % (k, n, v, 100 * v / tot, v / n))
    # one forward = from a stem_pool launch to the next
    idx = [i for i, (k, _) in enumerate(launches) if k.startswith('stem_pool')]
    if len(idx) >= 3:
        # the LAST complete forward with no build kernels in between: the timed batch-128 step (earlier forwards are the
        # bench's parity checks on small batches and the executor's first eager pass)
        n_min = min(b - a for a, b in zip(idx[:-1], idx[1:]))
        fw = [launches[a:b] for a, b in zip(idx[:-1], idx[1:]) if b - a == n_min][-1]
        f.write('\nOne forward (%d launches), us: %s = %.1f us\n' % (len(fw), ', '.join('%.1f' % v for _, v in fw), sum(v for _, v in fw)))
        conv = sum(v for k, v in fw if k.startswith(('conv_', 'stem_pool', 'gap_dense', 'pooled_dense')))
        f.write('Share of the roofline kernel set (stem_pool + conv_stack + conv_shift + conv_igemm + gap_dense | pooled_dense: every conv / dense layer) in that forward: %.1f%% '
                '(bench.py `roofline.kernel_share_of_step`, timed live with CUDA events, must agree with this share).\n'
                % (100 * conv / sum(v for _, v in fw)))
print('wrote profiles/%s_launches_summary.md (%d launches)' % (tag, len(launches)))
