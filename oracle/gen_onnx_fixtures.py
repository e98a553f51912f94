"""Write tests/golden/onnx/*.onnx with PyTorch's own (TorchScript) ONNX exporter, plus the torch outputs on a seeded input.

Run once in the authoring container:   python oracle/gen_onnx_fixtures.py [--resnet18]

The image has no ``onnx`` package, which torch needs only for a post-processing step that looks for custom onnx-script
functions (torch/onnx/_internal/torchscript_exporter/onnx_proto_utils.py:_add_onnxscript_fn); the protobuf itself is
serialised by torch's C++ exporter.  That one step is bypassed here, so the files are genuine exporter output and an
independent check of planer_b200/onnx_import.py (written against the ONNX schema, not against these files).
TEST INFRASTRUCTURE ONLY.  ``--resnet18`` additionally round-trips torchvision's ResNet-18 (46 MB, not committed) through the
importer and the numpy oracle and prints the error against torch.
"""
import io
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
OUT = os.path.join(ROOT, 'tests', 'golden', 'onnx')
os.makedirs(OUT, exist_ok=True)
warnings.filterwarnings('ignore')

from torch.onnx._internal.torchscript_exporter import onnx_proto_utils      # noqa: E402
onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes


def export(model, x, opset=13):
    f = io.BytesIO()
    torch.onnx.export(model, (x,), f, dynamo=False, opset_version=opset, input_names=['x'], output_names=['y'],
                      do_constant_folding=True)
    return f.getvalue()


def randomize_bn(model, gen):
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5, generator=gen)
            m.bias.data.normal_(0, 0.1, generator=gen)
            m.running_mean.normal_(0, 0.1, generator=gen)
            m.running_var.uniform_(0.5, 1.5, generator=gen)


class Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1, self.b1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm2d(cout)
        self.c2, self.b2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout)
        self.down = None if stride == 1 and cin == cout else nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                                                            nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.b2(self.c2(torch.relu(self.b1(self.c1(x)))))
        return torch.relu(y + (x if self.down is None else self.down(x)))


class MiniResNet(nn.Module):
    """conv7x7/s2 + bn + relu + maxpool, a plain and a down-sampling BasicBlock, gap + flatten + fc: ResNet-18's operators."""
    def __init__(self):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(3, 16, 7, 2, 3, bias=False), nn.BatchNorm2d(16), nn.ReLU(), nn.MaxPool2d(3, 2, 1))
        self.l1, self.l2 = Block(16, 16, 1), Block(16, 32, 2)
        self.pool, self.fc = nn.AdaptiveAvgPool2d(1), nn.Linear(32, 10)

    def forward(self, x):
        return self.fc(torch.flatten(self.pool(self.l2(self.l1(self.stem(x)))), 1))


class MiniDecoder(nn.Module):
    """The operators next to the path: LeakyReLU, AvgPool, ConvTranspose, nearest / bilinear Upsample (ONNX Resize), Concat,
    ReLU6 (ONNX Clip with input bounds), Hardsigmoid, Sigmoid, Softmax."""
    def __init__(self):
        super().__init__()
        self.c1, self.c2 = nn.Conv2d(3, 16, 3, 1, 1), nn.Conv2d(16, 32, 3, 1, 1)
        self.up = nn.ConvTranspose2d(32, 16, 4, 2, 1)
        self.c3, self.c4 = nn.Conv2d(32, 16, 3, 1, 1), nn.Conv2d(16, 6, 1)
        self.pool, self.bn = nn.AvgPool2d(2), nn.BatchNorm2d(16)      # a BatchNorm the exporter cannot fold into a conv

    def forward(self, x):
        e1 = nn.functional.leaky_relu(self.c1(x), 0.1)
        e2 = nn.functional.relu6(self.c2(self.bn(self.pool(e1))))
        d = torch.cat([e1, torch.relu(self.up(e2))], 1)
        d = nn.functional.hardsigmoid(self.c3(d))
        d = nn.functional.interpolate(d, scale_factor=2, mode='nearest')
        d = nn.functional.interpolate(self.c4(d), scale_factor=2, mode='bilinear', align_corners=False)
        return torch.softmax(d, 1), torch.sigmoid(d)


def main():
    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    for name, model, shape in (('mini_resnet', MiniResNet(), (2, 3, 64, 64)), ('mini_decoder', MiniDecoder(), (2, 3, 16, 24))):
        model.eval()
        randomize_bn(model, gen)
        x = torch.randn(shape, generator=gen)
        with torch.no_grad():
            y = model(x)
        ys = y if isinstance(y, tuple) else (y,)
        data = export(model, x)
        open(os.path.join(OUT, name + '.onnx'), 'wb').write(data)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), x=x.numpy(), **{'y%d' % i: t.numpy() for i, t in enumerate(ys)})
        print(name, len(data), 'bytes', [tuple(t.shape) for t in ys])
    if '--resnet18' in sys.argv:
        import torchvision
        import planer_oracle as oracle
        from planer_b200 import onnx_import
        m = torchvision.models.resnet18(weights=None).eval()
        randomize_bn(m, gen)
        x = torch.randn((2, 3, 224, 224), generator=gen)
        with torch.no_grad():
            y = m(x).numpy()
        model, blob = onnx_import.read_onnx(export(m, x))
        got = oracle.build_net(model, blob)(x.numpy())
        got = got[0] if isinstance(got, tuple) else got
        print('torchvision resnet18 -> ONNX -> importer -> numpy oracle vs torch: max abs err %.3g (range %.3g), %d layers, '
              'blob %d bytes' % (np.abs(got - y).max(), np.abs(y).max(), len(model['layers']), blob.size))


if __name__ == '__main__':
    main()
