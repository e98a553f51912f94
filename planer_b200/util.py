"""Sliding-window ("tiled") inference over large images: the reference's ``tile`` decorator (planer/util.py:291-348) and the
two helpers it needs (``resize`` planer/util.py:256-273, ``grid_slice`` planer/util.py:245-251), SURVEY 8f rank 3.

Same decorator arguments and the same arithmetic (float32 throughout, uint16 blending weights, the same window grid), so a
function wrapped here returns what it returns under the reference's decorator.  One addition: ``batched=True`` hands ALL
windows to the wrapped function in one call as a stacked array ``(n_windows, h, w[, c])`` and expects the stacked results
back -- windows are independent, so they ride the batch dimension of one forward (``net(x)`` / ``net.map``) instead of one
forward per window, which is what the B200 path wants (the reference calls ``f`` once per window, planer/util.py:321,339).

Host-side numpy only: the images live on the host, the wrapped function decides what runs on the GPU.
"""
import math

import numpy as np


def _axis_samples(n_in, n_out):
    """Source coordinates of ``n_out`` samples along an axis of length ``n_in`` (pixel centres, planer/util.py:259-266):
    integer base index, base + 1 and the float32 weight of the latter."""
    k = n_out / n_in
    pos = np.linspace(-0.5 + 0.5 / k, n_in - 0.5 - 0.5 / k, n_out, dtype=np.float32)
    pos = np.clip(pos, 0, n_in - 1, out=pos)
    lo = np.floor(np.clip(pos, 0, n_in - 1.001)).astype(int)
    pos -= lo
    return lo, lo + 1, pos


def resize(img, size, backend=None):
    """Bilinear resize of an (H, W[, C]) image to ``size`` (planer/util.py:256-273): columns first, then rows."""
    h, w = img.shape[:2]
    ra, rb, fr = _axis_samples(h, size[0])
    ca, cb, fc = _axis_samples(w, size[1])
    fr = fr.reshape((-1, 1, 1)[:img.ndim])
    fc = fc.reshape((1, -1, 1)[:img.ndim])
    cols = img[:, ca] * (1 - fc) + img[:, cb] * fc
    return cols[ra, :] * (1 - fr) + cols[rb, :] * fr


def make_slice(length, window, margin):
    """Window starts spread evenly over ``length`` with at least ``margin`` overlap (planer/util.py:245-247)."""
    count = math.ceil((length - margin) / (window - margin))
    starts = np.linspace(0, length - window, count).astype(int).tolist()
    return [slice(s, s + window) for s in starts]


def grid_slice(H, W, h, w, margin):
    """Row-major list of (row slice, column slice) windows (planer/util.py:249-251)."""
    return [(r, c) for r in make_slice(H, h, margin) for c in make_slice(W, w, margin)]


def _blend_weights(shape2, ramp, ndim):
    """uint16 window weights (planer/util.py:327-331): ``ramp + 1`` inside, falling 1 per pixel to 1 at the border -- the
    reference's loop over border rows and columns is min(distance to the nearest edge, ramp + 1)."""
    rows = np.minimum(np.arange(shape2[0]), np.arange(shape2[0])[::-1]) + 1
    cols = np.minimum(np.arange(shape2[1]), np.arange(shape2[1])[::-1]) + 1
    wts = np.minimum(np.minimum(rows[:, None], cols[None, :]), ramp + 1).astype('uint16')
    return wts[:, :, None] if ndim == 3 else wts


_TILE_KEYS = ('sample', 'window', 'glob', 'margin', 'progress', 'batched')


def tile(sample=1, glob=1, window=1024, margin=0.1, astype='float32', progress=print, batched=False):
    """Decorator: run ``f(img, *args)`` window by window over a large image and blend the overlaps (planer/util.py:291-348).

    sample : float factor or (H, W) tuple the image is resized to before tiling;
    glob   : an image smaller than the window is resized up to a multiple of ``glob``;
    window : window edge in pixels (after sampling);  margin: overlap between windows, float = fraction of ``window``;
    batched: the wrapped function takes/returns stacked windows (see module docstring).
    Every one of these can also be overridden per call as a keyword argument, like in the reference.
    """
    def decorate(f):
        def run(*args, **kwargs):
            own = {k: kwargs.pop(k) for k in list(kwargs) if k in _TILE_KEYS}
            src = args[0]
            h, w = src.shape[:2]
            img = src.astype('float32')
            ssz = own.get('sample', sample)
            wsz = own.get('window', window)
            gsz = own.get('glob', glob)
            mar = own.get('margin', margin)
            info = own.get('progress', progress)
            stack = own.get('batched', batched)
            ssz = list(ssz) if isinstance(ssz, tuple) else [int(h * ssz), int(w * ssz)]
            wh = ww = wsz
            if wh > ssz[0]: wh = ssz[0] = math.ceil(ssz[0] / gsz) * gsz
            if ww > ssz[1]: ww = ssz[1] = math.ceil(ssz[1] / gsz) * gsz
            resized = ssz != [h, w]
            if resized:
                img = resize(img, ssz)
            if isinstance(mar, float):
                mar = int(wsz * mar)
            wins = grid_slice(ssz[0], ssz[1], wh, ww, mar)
            n = len(wins)
            if stack:
                outs = f(np.stack([img[s] for s in wins]), *args[1:], **kwargs)
                result = lambda i: outs[i]
            else:
                cache = {}

                def result(i):
                    if i not in cache:
                        if n > 1: info(i + 1, n)
                        cache.clear()
                        cache[i] = f(img[wins[i]], *args[1:], **kwargs)
                    return cache[i]
            first = result(0)
            k = first.shape[0] / (wins[0][0].stop - wins[0][0].start)          # output pixels per input pixel
            if n == 1:
                return resize(first, (int(h * k), int(w * k))) if resized else first

            def scaled(win):
                return (slice(int(win[0].start * k), int(win[0].stop * k)), slice(int(win[1].start * k), int(win[1].stop * k)))

            out_shape = (int(img.shape[0] * k), int(img.shape[1] * k)) + first.shape[2:]
            wts = _blend_weights(first.shape[:2], int(mar * k), first.ndim)
            acc = np.zeros(out_shape, dtype=np.float32)
            cnt = np.zeros(out_shape[:2], dtype='uint16')
            if first.ndim == 3:
                cnt = cnt[:, :, None]
            acc[scaled(wins[0])] = first * wts
            cnt[scaled(wins[0])] += wts
            for i in range(1, n):
                acc[scaled(wins[i])] += result(i) * wts
                cnt[scaled(wins[i])] += wts
            np.divide(acc, cnt, out=acc, casting='unsafe')
            if resized:
                acc = resize(acc, (int(h * k), int(w * k)))
            return acc.astype(first.dtype)
        return run
    return decorate
