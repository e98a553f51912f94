"""torchvision ResNet-18 -> ONNX (PyTorch's exporter) -> planer_b200.read_net -> B200, against PyTorch itself; run under gpurun.

    python tools/onnx_resnet18_check.py [batch]

Prints the range-relative error of the fp16 logits against torch's fp32 forward and the device-timed throughput of the
imported graph (BatchNorm already folded into the convolutions by the exporter: conv + bias + relu epilogues)."""
import ctypes as C, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, torchvision
import gen_onnx_fixtures as G
import planer_b200 as planer
from planer_b200 import backend as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gen = torch.Generator().manual_seed(0)
torch.manual_seed(0)
m = torchvision.models.resnet18(weights=None).eval()
G.randomize_bn(m, gen)
x = torch.randn((n, 3, 224, 224), generator=gen)
with torch.no_grad():
    ref = m(x[:8]).numpy()
path = os.path.join(tempfile.mkdtemp(), 'resnet18.onnx')
open(path, 'wb').write(G.export(m, x[:1]))
planer.core(planer.b200)
net = planer.read_net(path)
net.half()
xh = x.numpy().astype(np.float16)
y = net(xh)
err = float(np.abs(y[:8].astype(np.float64) - ref).max() / np.abs(ref).max())
xd = B.asarray(xh)
for _ in range(5):
    net.forward(xd)
lib, ctx = B.lib(), B.ctx()
a, b = C.c_void_p(), C.c_void_p()
lib.plnr_event_create(C.byref(a)); lib.plnr_event_create(C.byref(b))
lib.plnr_event_record(ctx, a)
for _ in range(50):
    net.forward(xd)
lib.plnr_event_record(ctx, b)
B.synchronize()
ms = C.c_float(); lib.plnr_event_elapsed_ms(a, b, C.byref(ms))
ex = net.executor([xh.shape])
print('torchvision resnet18 via ONNX: %d layers -> %d launches; fp16 logits vs torch fp32: %.2e range-relative; batch %d: %.3f ms, %.0f img/s'
      % (len(net.layer), len(ex.launches) + len(ex.fused_stems), err, n, ms.value / 50, n * 50 / ms.value * 1e3))
