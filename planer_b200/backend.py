"""The B200 backend module handed to ``planer_b200.core()`` (Level A of SURVEY 8b).

In the reference, ``planer.core(obj)`` rebinds the array module ``np`` used by util/layer/net/io
(planer/__init__.py:22-38) and the objects it hands around must offer ``asarray`` / ``asnumpy`` /
``zeros`` / ``load`` and arrays with ``shape, dtype, get(), astype(), reshape()``.  This module is that
object for the B200 path.  It does not emulate numpy slicing: the operators themselves are replaced one
level up (``planer_b200.layer.layer_map``), each calling hand-written sm_100a kernels through the C ABI.

PyTorch is used here for plumbing only: device memory (caching allocator), the CUDA stream the library
enqueues on, pinned host staging and ``torch.distributed``.  No torch operator computes anything.
"""
import ctypes as C

import numpy as np

from . import _capi

__name__ = 'planer_b200'          # what planer.core() prints / compares (planer/__init__.py:29-37)

float32, float16, uint8 = np.float32, np.float16, np.uint8

_state = {'ctx': None, 'device': None, 'stream': None, 'torch': None}


def _torch():
    if _state['torch'] is None:
        import torch
        _state['torch'] = torch
    return _state['torch']


def _torch_dtype(dt):
    torch = _torch()
    return {np.dtype('float32'): torch.float32, np.dtype('float16'): torch.float16,
            np.dtype('uint8'): torch.uint8, np.dtype('int32'): torch.int32,
            np.dtype('int64'): torch.int64, np.dtype('float64'): torch.float64,
            np.dtype('bool'): torch.bool, np.dtype('int8'): torch.int8,
            np.dtype('int16'): torch.int16}[np.dtype(dt)]


def init(device=None):
    """Create (once) the library context on ``device`` bound to a dedicated torch CUDA stream."""
    if _state['ctx'] is not None:
        return _state['ctx']
    lib = _capi.load()
    torch = _torch()
    if not torch.cuda.is_available():
        raise _capi.PlanerB200Error('no CUDA device visible: planer_b200 runs on B200 (sm_100a) only, no CPU fallback')
    if device is None:
        import os
        device = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(device)
    stream = torch.cuda.Stream(device=device)
    ctx = C.c_void_p()
    _capi.check(lib.plnr_create(device, C.c_void_p(stream.cuda_stream), C.byref(ctx)), 'plnr_create')
    _state.update(ctx=ctx, device=device, stream=stream)
    return ctx


def ctx():
    return init()


def lib():
    return _capi.load()


def stream():
    init()
    return _state['stream']


def device():
    init()
    return _state['device']


def synchronize():
    _capi.check(lib().plnr_stream_sync(ctx()), 'plnr_stream_sync')


def launch_count():
    n = C.c_int64()
    _capi.check(lib().plnr_launch_count(ctx(), C.byref(n)))
    return n.value


def last_kernel():
    """Family name of the kernel the most recent library call launched (diagnostics / bench tables)."""
    buf = C.create_string_buffer(64)
    _capi.check(lib().plnr_last_kernel(ctx(), buf, 64))
    return buf.value.decode()


def device_info():
    out = (C.c_int64 * 4)()
    _capi.check(lib().plnr_device_info(ctx(), out))
    return {'sm_count': out[0], 'cc': (out[1], out[2]), 'l2_bytes': out[3]}


class DeviceArray:
    """A device tensor.  ``layout`` is 'flat' (row-major of ``shape``: weights, NCHW graph inputs/outputs,
    2-D dense activations) or 'nhwc' (internal 4-D activations: pixel rows of ``ld`` elements, this tensor's
    channels at ``[coff, coff+C)``).  ``shape`` is always the logical NCHW / numpy shape."""

    __slots__ = ('buf', 'shape', 'dtype', 'layout', 'ld', 'coff', 'offset')

    def __init__(self, buf, shape, dtype, layout='flat', ld=None, coff=0, offset=0):
        self.buf, self.shape, self.dtype = buf, tuple(int(s) for s in shape), np.dtype(dtype)
        self.layout, self.coff, self.offset = layout, coff, offset
        self.ld = (self.shape[1] if ld is None else ld) if layout == 'nhwc' else None

    # -- numpy-like metadata ------------------------------------------------------------------
    @property
    def ndim(self): return len(self.shape)

    @property
    def size(self):
        n = 1
        for s in self.shape: n *= s
        return n

    @property
    def nbytes(self): return self.size * self.dtype.itemsize

    @property
    def ptr(self):
        """Device address of element 0 (for 'nhwc': pixel 0, channel 0 of the underlying rows)."""
        return self.buf.data_ptr() + self.offset

    def tensor(self):
        """plnr_tensor view.  2-D flat arrays (M, K) are M pixels of K channels."""
        if self.layout == 'nhwc':
            n, c, h, w = self.shape
            return _capi.Tensor(self.ptr, n, h, w, c, self.ld, self.coff)
        if self.ndim == 2:
            m, k = self.shape
            return _capi.Tensor(self.ptr, 1, 1, m, k, k, 0)
        raise _capi.PlanerB200Error('tensor(): need an nhwc or 2-D array, got %s %s' % (self.layout, self.shape))

    def reshape(self, *shape):
        """Metadata-only reshape of a flat array (planer/layer.py:59 Flatten, weight views)."""
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        if self.layout != 'flat':
            return to_flat(self).reshape(shape)
        shape = list(shape)
        if -1 in shape:
            known = 1
            for s in shape:
                if s != -1: known *= s
            shape[shape.index(-1)] = self.size // max(known, 1)
        assert int(np.prod(shape)) == self.size, (shape, self.shape)
        return DeviceArray(self.buf, shape, self.dtype, 'flat', offset=self.offset)

    # -- the handful of numpy idioms the reference's Net.load_weights uses on weights (planer/net.py:83-88):
    #    data.view(uint8); w.ravel().view(uint8)[:] = data[s:s+n]
    def ravel(self):
        assert self.layout == 'flat', 'ravel() of an internal nhwc array is not a view'
        return DeviceArray(self.buf, (self.size,), self.dtype, 'flat', offset=self.offset)

    def view(self, dtype=None):
        dtype = np.dtype(dtype)
        assert self.layout == 'flat'
        nbytes = self.size * self.dtype.itemsize
        assert nbytes % dtype.itemsize == 0
        shape = (nbytes // dtype.itemsize,) if self.ndim <= 1 else self.shape[:-1] + (
            self.shape[-1] * self.dtype.itemsize // dtype.itemsize,)
        return DeviceArray(self.buf, shape, dtype, 'flat', offset=self.offset)

    def __getitem__(self, key):
        if not (self.layout == 'flat' and self.ndim == 1 and isinstance(key, slice) and key.step in (None, 1)):
            raise NotImplementedError('DeviceArray supports only contiguous 1-D slices (weight-blob views); '
                                      'operators are replaced one level up, not emulated through numpy indexing')
        start, stop, _ = key.indices(self.shape[0])
        return DeviceArray(self.buf, (max(stop - start, 0),), self.dtype, 'flat',
                           offset=self.offset + start * self.dtype.itemsize)

    def __setitem__(self, key, value):
        dst = self[key] if not (isinstance(key, slice) and key == slice(None)) else self.ravel() if self.layout == 'flat' else None
        if dst is None:
            raise NotImplementedError('DeviceArray assignment supports whole-array / 1-D slice copies only')
        src = asarray(value) if not isinstance(value, DeviceArray) else to_flat(value)
        if src.size != dst.size or src.dtype.itemsize != dst.dtype.itemsize:
            raise ValueError('shape mismatch in device copy: %s <- %s' % (dst.shape, src.shape))
        with _torch().cuda.stream(stream()):
            dst._typed().copy_(src._typed().view(_torch_dtype(dst.dtype)), non_blocking=True)

    def astype(self, dtype):
        dtype = np.dtype(dtype)
        src = to_flat(self)
        if dtype == src.dtype:
            return src
        out = empty(src.shape, dtype)
        _capi.check(lib().plnr_cast(ctx(), src.ptr, _capi.dtype_code(src.dtype), out.ptr, _capi.dtype_code(dtype),
                                    src.size), 'plnr_cast')
        return out

    def get(self):
        """Device -> host numpy array of the logical shape (cupy convention; planer/net.py:100)."""
        torch = _torch()
        src = to_flat(self)
        with torch.cuda.stream(stream()):
            host = src._typed().cpu()
        return host.numpy().reshape(self.shape)

    def _typed(self):
        """torch view of the flat storage with the element dtype (plumbing for copies only)."""
        nbytes = self.size * self.dtype.itemsize
        raw = self.buf.view(_torch().uint8).reshape(-1)[self.offset:self.offset + nbytes]
        return raw.view(_torch_dtype(self.dtype))

    def __repr__(self):
        return 'DeviceArray(shape=%s, dtype=%s, layout=%s)' % (self.shape, self.dtype, self.layout)


def empty(shape, dtype=np.float32, layout='flat', ld=None):
    torch = _torch()
    init()
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    dtype = np.dtype(dtype)
    n = 1
    if layout == 'nhwc':
        nb, c, h, w = shape
        ld = c if ld is None else ld
        n = nb * h * w * ld
    else:
        for s in shape: n *= s
    with torch.cuda.stream(stream()):
        buf = torch.empty(max(n, 1) * dtype.itemsize + 16, dtype=torch.uint8, device='cuda:%d' % device())
    return DeviceArray(buf, shape, dtype, layout, ld)


def zeros(shape, dtype=np.float32):
    """planer/net.py:21 allocates every init with ``np.zeros``."""
    out = empty(shape, dtype)
    _capi.check(lib().plnr_memset(ctx(), out.ptr, 0, max(out.nbytes, 1)), 'plnr_memset')
    return out


def asarray(a, dtype=None, pinned=False):
    """Host -> device (planer/net.py:98).  DeviceArrays pass through."""
    if isinstance(a, DeviceArray):
        return a if dtype is None else a.astype(dtype)
    torch = _torch()
    init()
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype))
    out = empty(a.shape, a.dtype)
    if a.size:
        host = torch.from_numpy(a.reshape(-1).view(np.uint8))
        with torch.cuda.stream(stream()):
            out.buf[:a.nbytes].copy_(host, non_blocking=True)
    return out


def copy_stream():
    """The stream host<->device copies of pipelined calls run on (Net.map, chunked Net.__call__)."""
    init()
    if _state.get('copy_stream') is None:
        _state['copy_stream'] = _torch().cuda.Stream(device=device())
    return _state['copy_stream']


def pinned_empty(shape, dtype=np.float32):
    """Page-locked host array (numpy view of a pinned torch buffer): uploads from it run at the full PCIe rate and
    asynchronously; a pageable numpy array is staged by the driver at a fraction of that."""
    torch = _torch()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    return torch.empty(max(n, 1), dtype=torch.uint8).pin_memory().numpy()[:n].view(dtype).reshape(shape)


def asarray_async(a):
    """Host -> device on a separate COPY stream.  Returns (DeviceArray, event): work on the library stream that reads
    the array must first ``stream().wait_event(event)``.  Lets the upload of one chunk of a batch overlap the forward
    of the previous chunk (Net.__call__, planer/net.py:94-101 is one blocking asarray)."""
    torch = _torch()
    init()
    cs = copy_stream()
    a = np.ascontiguousarray(a)
    out = empty(a.shape, a.dtype)
    cs.wait_stream(stream())                    # the allocation above was made on the library stream
    host = torch.from_numpy(a.reshape(-1).view(np.uint8))
    with torch.cuda.stream(cs):
        out.buf[:a.nbytes].copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
    return out, ev


def clone(a):
    """Device copy of a flat array on the library stream (results handed to the caller must outlive the next forward)."""
    src = to_flat(a)
    out = empty(src.shape, src.dtype)
    if src.nbytes:
        with _torch().cuda.stream(stream()):
            out.buf[:src.nbytes].copy_(src.buf[src.offset:src.offset + src.nbytes], non_blocking=True)
    return out


def asnumpy(a):
    return a.get() if isinstance(a, DeviceArray) else np.asarray(a)


def load(path_or_file):
    """``np.load`` of the weight blob, uploaded (planer/io.py:18,24 call the *backend's* load)."""
    return asarray(np.load(path_or_file))


def to_flat(a):
    """Any DeviceArray -> dense row-major ('flat') array of its logical shape (NCHW for 4-D)."""
    if a.layout == 'flat':
        return a
    out = empty(a.shape, a.dtype)
    code = _capi.dtype_code(a.dtype)
    t = a.tensor()
    _capi.check(lib().plnr_nhwc_to_nchw(ctx(), C.byref(t), code, out.ptr, code), 'plnr_nhwc_to_nchw')
    return out


def to_nhwc(a, dtype=None, cpad=None):
    """Flat NCHW (4-D) -> internal pixel-major array, optionally casting and zero-padding channels."""
    dtype = np.dtype(a.dtype if dtype is None else dtype)
    if a.layout == 'nhwc':
        assert dtype == a.dtype
        return a
    n, c, h, w = a.shape
    cp = c if cpad is None else cpad
    out = empty((n, cp, h, w), dtype, 'nhwc')
    t = out.tensor()
    _capi.check(lib().plnr_nchw_to_nhwc(ctx(), a.ptr, _capi.src_dtype_code(a.dtype), c, C.byref(t),
                                        _capi.dtype_code(dtype)), 'plnr_nchw_to_nhwc')
    return out
