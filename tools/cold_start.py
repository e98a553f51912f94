"""Cold start with and without the pre-packed weight cache: read_net -> half -> first forward, in fresh processes.

    python tools/cold_start.py            # writes a ResNet-18 model to /tmp, runs the two cases, prints a table
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CHILD = r'''
import sys, time, json
sys.path.insert(0, %r)
import numpy as np
t_imp = time.perf_counter()
import planer_b200 as planer
from planer_b200 import backend as B
planer.core(planer.b200)
B.init(); B.synchronize()
x = np.random.default_rng(0).standard_normal((%d, 3, 224, 224)).astype(np.float16)
t0 = time.perf_counter()
net = planer.read_net(%r); net.half()
t1 = time.perf_counter()
l0 = B.launch_count()
y = net(x)
t2 = time.perf_counter()
ex = net.executor([x.shape], [x.dtype])
if %d: planer.save_pack(net)
print(json.dumps({'read_net_s': t1 - t0, 'first_forward_s': t2 - t1, 'launches_first_forward': B.launch_count() - l0,
                  'pack_hits': ex.pack_hits, 'pack_misses': ex.pack_misses, 'checksum': float(np.abs(y.astype(np.float32)).sum())}))
'''


def run(path, batch, save):
    out = subprocess.run([sys.executable, '-c', CHILD % (ROOT, batch, path, save)], capture_output=True, text=True)
    if out.returncode:
        print(out.stderr[-2000:])
        raise SystemExit(1)
    return json.loads(out.stdout.strip().splitlines()[-1])


if __name__ == '__main__':
    from planer_b200 import zoo
    path = '/tmp/cold_r18'
    model, blob = zoo.resnet18(0)
    zoo.save_model(path, model, blob)
    if os.path.exists(path + '.b200pack.npz'):
        os.remove(path + '.b200pack.npz')
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    a = run(path, batch, 1)          # no cache: builds and writes it
    b = run(path, batch, 0)          # cache hit
    c = run(path, batch, 0)
    print('| case | read_net + half (s) | first forward incl. executor build (s) | launches | pack hits / misses |')
    print('|---|---|---|---|---|')
    for name, r in (('no cache', a), ('cache hit', b), ('cache hit (2nd run)', c)):
        print('| %s | %.3f | %.3f | %d | %d / %d |' % (name, r['read_net_s'], r['first_forward_s'], r['launches_first_forward'],
                                                       r['pack_hits'], r['pack_misses']))
    assert a['checksum'] == b['checksum'] == c['checksum'], 'outputs differ between the cached and the uncached run'
