"""Headline benchmark: ResNet-18 224x224 forward images/sec on N B200s (BASELINE.json metric), fp16,
batch 128 per GPU (configs[2]; configs[4] = 8 x 128 is the same workload weak-scaled to 8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  A "step" is one forward of the whole graph over one batch of synthetic
NCHW input.  `value` is device-timed with inputs resident in HBM; `e2e` goes through the public API from
HOST memory -- ``for y in net.map(batches)``, a pinned numpy batch uploaded and the logits read back for
every step inside the timed region, two batches in flight -- and `e2e.blocking_call` is the same through
the reference-shaped blocking ``y = net(x)`` per batch.
The reference arm times the numpy restatement of the reference (oracle/planer_oracle.py, "port") on the
host cores -- the only place outside tests/ and smoke() where oracle/ is executed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'resnet18_224_fwd_images_per_sec'
UNIT = 'images/s'
BATCH = 128                       # per GPU
IN_SHAPE = (3, 224, 224)
N_INPUT_BUFFERS = 8               # 8 x 38.5 MB = 308 MB of rotating inputs > 126 MB L2


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'tflops_burst': p['bf16_tflops'], 'tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'hbm_gbs': p['hbm_gbs'], 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'hbm_gbs': 6650.0,
            'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def cpu_baseline_run(steps, warmup, batch=8):
    """Reference numpy path (restated in oracle/planer_oracle.py) on the host cores: ResNet-18 fp32."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import planer_oracle as oracle
    from planer_b200 import zoo
    model, blob = zoo.resnet18(0)
    net = oracle.build_net(model, blob)
    x = np.random.default_rng(1).standard_normal((batch,) + IN_SHAPE).astype(np.float32)
    cores = len(os.sched_getaffinity(0))
    # torchrun exports OMP_NUM_THREADS=1, which OpenBLAS obeys at import: give the reference path every host core back
    import contextlib
    try:
        from threadpoolctl import threadpool_limits
        blas = threadpool_limits(limits=cores)
    except ImportError:
        blas = contextlib.nullcontext()
    with blas:
        for _ in range(warmup):
            net(x.copy())
        t0 = time.perf_counter()
        for _ in range(steps):
            net(x.copy())
        dt = time.perf_counter() - t0
    return {'value': batch * steps / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'ResNet-18 fp32 (numpy fp16 matmul has no BLAS: 18 s/img) batch %d x %d forwards, numpy %s '
                      'BLAS threads=all cores' % (batch, steps, np.__version__)}, dt / steps


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    cb, sec = cpu_baseline_run(steps, max(1, min(args.warmup, 2)))
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': steps, 'warmup': max(1, min(args.warmup, 2)), 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'ResNet-18 224x224 forward, reference numpy path on host cores, bounded sample: '
                                   'batch 8 per step'},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    _emit(line)


def per_kernel_profile(net, x_dev, flops_by_step, reps=3):
    """Un-graphed pass with a CUDA event pair around every launch: durations per step (ms)."""
    import ctypes as C
    from planer_b200 import _capi, backend as B
    ex = net.executor([x_dev.shape])
    lib, ctx = B.lib(), B.ctx()
    fns = [lambda: ex._load_inputs([x_dev])] + list(ex.launches)
    n = len(fns)
    best = [float('inf')] * n
    for _ in range(reps):
        evs = []
        for fn in fns:
            a, b = C.c_void_p(), C.c_void_p()
            lib.plnr_event_create(C.byref(a)); lib.plnr_event_create(C.byref(b))
            lib.plnr_event_record(ctx, a)
            fn()
            lib.plnr_event_record(ctx, b)
            evs.append((a, b))
        B.synchronize()
        for i, (a, b) in enumerate(evs):
            ms = C.c_float()
            lib.plnr_event_elapsed_ms(a, b, C.byref(ms))
            best[i] = min(best[i], ms.value)
            lib.plnr_event_destroy(a); lib.plnr_event_destroy(b)
    fused = [f for f in ex.fused_stems.values()]
    in_kind = 'conv' if fused else 'input'
    in_name = ('+'.join(fused[0]['conv'].fused + fused[0]['pool'].fused) + ' (one kernel, at input time)') if fused else 'input layout'
    return [{'kind': k, 'name': nm, 'ms': t} for k, nm, t in zip([in_kind] + list(ex.kinds), [in_name] + list(ex.names), best)]


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything a library prints on fd 1 while the bench runs (NCCL's version banner, for one) goes to stderr: stdout carries
    exactly ONE line, the JSON result."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH, help='images per GPU (the headline config is 128)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--min-warm-sec', type=float, default=1.5, help='keep warming up at least this long (clock sampling)')
    ap.add_argument('--dump', default=None, help='write the per-kernel table to this JSON file')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank)

    import torch
    import planer_b200 as planer
    from planer_b200 import zoo, dist, backend as B
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    warmup, steps = max(args.warmup, 3), args.steps
    planer.core(planer.b200)
    B.init(local)

    # ---- model: rank 0 builds the blob, ONE NCCL broadcast hands it to the other ranks ----
    model, blob = zoo.resnet18(0)
    net = planer.Net()
    net.load_json(model['input'], model['inits'], model['layers'], model['flow'])
    net.load_weights(blob if rank == 0 else None) if world > 1 else net.load_weights(blob)
    net.half()
    del blob

    rng = np.random.default_rng(100 + rank)
    shape = (args.batch,) + IN_SHAPE
    hosts = [rng.standard_normal(shape).astype(np.float16) for _ in range(2)]
    xs = [B.asarray(hosts[i % 2] if i < 2 else np.roll(hosts[i % 2], i, axis=0)) for i in range(N_INPUT_BUFFERS)]
    B.synchronize()

    ex = net.executor([shape])
    flops = ex.plan.flops
    # clocks are sampled from the start of the warm-up to the end of the timed region: the same load throughout.
    # The warm-up runs at least W steps AND at least ~1.5 s so that nvidia-smi (200 ms period) sees the loaded state.
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t_w, i = time.perf_counter(), 0
    while i < warmup or time.perf_counter() - t_w < args.min_warm_sec:
        net.forward(xs[i % N_INPUT_BUFFERS])
        i += 1
        if i % 50 == 0:
            B.synchronize()
    B.synchronize()
    warmup = i

    # ---- timed region: device events on the library stream, barrier + sync on both sides ----
    stream = B.stream()
    dist.barrier(); torch.cuda.synchronize()
    l0 = B.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(steps):
            net.forward(xs[i % N_INPUT_BUFFERS])
        e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    ms_total = dist.max_over_ranks(e0.elapsed_time(e1))
    launches = B.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    value = world * args.batch * steps / (ms_total / 1e3)

    # ---- e2e: public API with host buffers (pinned), H2D + forward + D2H of EVERY step inside the timed region ----
    # (a) Net.map: the call a user with a stream of batches makes -- upload of batch i+1 / forward of batch i /
    #     download of batch i-1 overlap; (b) the blocking Net.__call__ per batch (planer/net.py:94-101), reported next to it.
    pinned = []
    for h in hosts:
        p = B.pinned_empty(h.shape, h.dtype)
        p[...] = h
        pinned.append(p)
    e2e_steps = max(5, steps // 2)

    def feed(n):
        for i in range(n):
            yield pinned[i % 2]

    for y in net.map(feed(4)):
        pass
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    nout = 0
    for y in net.map(feed(e2e_steps)):
        nout += y.shape[0]
    torch.cuda.synchronize()
    e2e_sec = dist.max_over_ranks(time.perf_counter() - t0)
    assert nout == args.batch * e2e_steps
    e2e_value = world * args.batch * e2e_steps / e2e_sec
    for i in range(3):
        net(pinned[i % 2])
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        y = net(pinned[i % 2])
    torch.cuda.synchronize()
    blk_sec = dist.max_over_ranks(time.perf_counter() - t0)
    e2e_blocking = world * args.batch * e2e_steps / blk_sec
    h2d, d2h = int(hosts[0].nbytes), int(y.nbytes)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv), timed live per launch ----
    table = per_kernel_profile(net, xs[0], None)
    conv_ms = sum(r['ms'] for r in table if r['kind'] in ('conv', 'dense', 'gap'))
    all_ms = sum(r['ms'] for r in table)
    conv_nodes = [n for n in ex.plan.nodes if n.kind in ('conv', 'dense')]
    pk = peaks()
    achieved = flops / (conv_ms / 1e3) / 1e12
    traffic, traffic_src = None, None
    # newest committed ncu --set full capture of one step (tools/gpu_profile.sh + tools/ncu_step_summary.py)
    cands = sorted(f for f in os.listdir(os.path.join(ROOT, 'profiles')) if f.endswith('_step_traffic.json'))
    if cands and args.batch == BATCH:
        with open(os.path.join(ROOT, 'profiles', cands[-1])) as f:
            tj = json.load(f)
        traffic = sum(int((k['dram_read_mb'] + k['dram_write_mb']) * 1e6) for k in tj['kernels'])
        traffic_src = ('dram__bytes_read.sum + dram__bytes_write.sum summed over the same launches of one step, ncu --set full '
                       '(profiles/%s)' % cands[-1].replace('_traffic.json', '_full.md'))
    roofline = {'bound': 'tensor',
                'kernel': 'tcgen05 conv kernels of one step: stem_pool + conv_stack_f16 + conv_shift_f16 + conv_igemm_f16 + gap_dense (%d conv/dense layers)' % len(conv_nodes),
                'achieved': achieved, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['tflops_sustained'], 'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': pk['source'] + ', sustained figure (kernel timed inside a long step)',
                'frac_of_burst_peak': achieved / pk['tflops_burst'],
                'flops_per_step': flops, 'kernel_ms_per_step': conv_ms, 'kernel_share_of_step': conv_ms / all_ms,
                'algorithmic_bytes_per_step': int(args.batch * 9.34e6 + 23.4e6),
                'whole_step_frac': (flops * steps / (ms_total / 1e3) / 1e12) / pk['tflops_sustained']}
    if args.dump:
        by_name = {n.name: n for n in conv_nodes}
        for r in table:
            nd = by_name.get(r['name'].split('+')[0])
            if r['kind'] in ('conv', 'dense') and nd is not None:
                r.update(gflop=nd.flops / 1e9, tflops=nd.flops / (r['ms'] / 1e3) / 1e12)
        with open(args.dump, 'w') as f:
            json.dump({'batch': args.batch, 'table': table}, f, indent=1)

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_baseline_run(steps=8, warmup=1)

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
            'ms_per_step': ms_total / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16', 'data': 'synthetic',
            'config': {'workload': 'ResNet-18 224x224 fp16 forward, batch %d per GPU (BASELINE configs[2]; x8 = configs[4])'
                                   % args.batch, 'global_batch': world * args.batch, 'parallelism': 'dp%d batch split, no forward collective' % world,
                       'l2': 'inputs rotate over %d device buffers (%.0f MB > 126 MB L2)' % (N_INPUT_BUFFERS, N_INPUT_BUFFERS * h2d / 1e6),
                       'launch': '1 input-time kernel (fused first layer) + 1 CUDA graph (%d fused kernels) per step' % len(ex.launches)},
            'clocks': clk, 'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps, 'api': 'for y in net.map(batches): pinned numpy batch in, numpy logits out, 2 batches in flight',
                    'blocking_call': {'value': e2e_blocking, 'unit': UNIT, 'api': 'y = net(x) per batch (upload in two halves, one synchronisation)'}},
            'roofline': roofline, 'cpu_baseline': cb}
    _emit(line)


if __name__ == '__main__':
    main()
    try:
        import torch.distributed as _d
        if _d.is_available() and _d.is_initialized():
            _d.destroy_process_group()
    except Exception:
        pass
