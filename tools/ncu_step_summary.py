"""Summarise one steady-state step captured by tools/gpu_profile.sh (ncu --set full) into profiles/:
    python tools/ncu_step_summary.py gpurun_out/prof_step.ncu-rep r01b
writes profiles/<tag>_step_full.md (per-kernel table) and profiles/<tag>_step_traffic.json (DRAM bytes per launch, read by bench.py)."""
import csv, hashlib, io, json, os, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_to=None):
    v = float(r[col[name]].replace(',', ''))
    u = units[col[name]]
    if scale_to == 'MB':
        v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[u]
    if scale_to == 'us':
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'msecond': 1e3, 'usecond': 1.0, 'nsecond': 1e-3}[u]
    return v


kern = []
for r in data:
    name = r[col['Kernel Name']]
    short = name.split('(')[0].split('::')[-1]
    kern.append({'kernel': short, 'grid': r[col['Grid Size']],
                 'us': val(r, 'gpu__time_duration.sum', 'us'),
                 'dram_read_mb': val(r, 'dram__bytes_read.sum', 'MB'), 'dram_write_mb': val(r, 'dram__bytes_write.sum', 'MB'),
                 'tensor_pct': val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')})
tot = sum(k['us'] for k in kern)
with open('profiles/%s_step_traffic.json' % tag, 'w') as f:
    # the capture is of the library that is in the tree NOW (gpurun ships the in-tree build): bench.py reports `traffic` only
    # when the library it runs has this hash
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'planer_b200', 'libplaner_b200.so')
    sha = hashlib.sha256(open(lib, 'rb').read()).hexdigest()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hs = hashlib.sha256()                      # same definition as bench.py:src_sha256
    csrc = os.path.join(root, 'planer_b200', 'csrc')
    for fn in sorted(os.path.join(csrc, f_) for f_ in os.listdir(csrc) if f_.endswith(('.cu', '.cuh'))) + [os.path.join(root, 'include', 'planer_b200.h')]:
        hs.update(os.path.basename(fn).encode())
        hs.update(open(fn, 'rb').read())
    json.dump({'source': rep, 'lib_sha256': sha, 'src_sha256': hs.hexdigest(), 'config': 'resnet18', 'kernels': kern}, f, indent=1)
with open('profiles/%s_step_full.md' % tag, 'w') as f:
    f.write('# One steady-state step under `ncu --set full` (ResNet-18 fp16, batch 128, B200) -- %s\n\n' % tag)
    f.write('Command: tools/gpu_profile.sh (`ncu --set full --clock-control none --import-source on -k regex:conv_shift|conv_stack|'
            'conv_igemm|stem_pool|gap_dense|pooled_dense -s 36 -c 18 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --min-warm-sec 0`).\n'
            'Durations are cold-cache and serialised: compare shares.  Metrics: `gpu__time_duration.sum`, `dram__bytes_read.sum`, '
            '`dram__bytes_write.sum`, `sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active`.\n\n')
    f.write('| # | kernel | grid | us | share | DRAM read MB | DRAM write MB | tensor pipe active % |\n|---|---|---|---|---|---|---|---|\n')
    for i, k in enumerate(kern):
        f.write('| %d | `%s` | %s | %.1f | %.1f%% | %.1f | %.1f | %.1f |\n' % (i, k['kernel'], k['grid'], k['us'], 100 * k['us'] / tot,
                                                                          k['dram_read_mb'], k['dram_write_mb'], k['tensor_pct']))
    f.write('| | **sum** | | %.1f | | %.1f | %.1f | |\n' % (tot, sum(k['dram_read_mb'] for k in kern), sum(k['dram_write_mb'] for k in kern)))
print('wrote profiles/%s_step_full.md, %d kernels, %.1f us' % (tag, len(kern), tot))
