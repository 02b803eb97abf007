"""Helpers shared by the `-m gpu` parity tests: same seeded inputs to the CUDA path (through
the C ABI) and to the oracle."""
import numpy as np

from rav1e_b200 import backend as B
from tests import oracle_lib as O

_CTX = None


def ctx():
    global _CTX
    if _CTX is None:
        # kernels go to torch's current stream: the tests fill torch tensors (torch.zeros, .cuda())
        # on that stream and the ctx's own stream is non-blocking, so sharing one stream is what
        # orders "fill, then launch" (the ordering contract stated in include/b200rdo.h)
        _CTX = B.Context(0, use_torch_stream=True)
    return _CTX


def to_dev(a):
    import torch
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()


def from_dev(t, dtype):
    return t.cpu().numpy().view(dtype)


def dev_empty(nbytes):
    import torch
    return torch.empty(max(nbytes, 1), dtype=torch.uint8, device="cuda")


def make_planes(width, height, pad, dtype=np.uint8, seed=0, bit_depth=8, smooth=True, shift=None):
    """`ref` = (optionally low-passed) noise, `cur` = ref displaced per-frame + noise, so that
    SAD surfaces have real minima.  Returns numpy images (visible area)."""
    rng = np.random.default_rng(seed)
    maxv = (1 << bit_depth) - 1
    base = rng.integers(0, maxv + 1, (height + 64, width + 64)).astype(np.float64)
    if smooth:
        k = np.ones(5) / 5.0
        base = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, base)
        base = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, base)
        base = (base - base.min()) / (base.max() - base.min()) * maxv
    dx, dy = shift if shift is not None else (3, -2)
    ref = base[32:32 + height, 32:32 + width]
    cur = base[32 + dy:32 + dy + height, 32 + dx:32 + dx + width] + rng.normal(0, 2, (height, width))
    # C-contiguous: callers hand .ctypes.data / .strides[0] to the C ABI
    ref = np.ascontiguousarray(np.clip(np.rint(ref), 0, maxv).astype(dtype))
    cur = np.ascontiguousarray(np.clip(np.rint(cur), 0, maxv).astype(dtype))
    return cur, ref


def both_planes(img, pad):
    """(oracle Plane, device b200_plane) of the same image, edges replicated into `pad`."""
    op = O.Plane(img.shape[1], img.shape[0], pad, dtype=img.dtype)
    op.fill_from(img)
    dp = ctx().plane_from_host(img, pad)
    return op, dp


def grid_blocks(width, height, bw, bh):
    xs = np.arange(0, width - bw + 1, bw)
    ys = np.arange(0, height - bh + 1, bh)
    b = np.zeros(len(xs) * len(ys), B.BLOCK_DTYPE)
    b["x"] = np.tile(xs, len(ys))
    b["y"] = np.repeat(ys, len(xs))
    return b


def random_cands(nblocks, per_block, max_px, seed=2, fullpel=True, jitter=True):
    """Grouped-by-block candidate list + CSR offsets.  per_block may vary +-50% per block."""
    rng = np.random.default_rng(seed)
    if jitter:
        counts = rng.integers(max(per_block // 2, 0), per_block * 3 // 2 + 1, nblocks)
    else:
        counts = np.full(nblocks, per_block)
    offs = np.zeros(nblocks + 1, np.uint32)
    offs[1:] = np.cumsum(counts)
    n = int(offs[-1])
    c = np.zeros(n, B.CAND_DTYPE)
    c["block"] = np.repeat(np.arange(nblocks, dtype=np.uint32), counts)
    mv = rng.integers(-max_px, max_px + 1, (n, 2))
    if fullpel:
        mv = mv * 8
    else:
        mv = mv * 8 + rng.integers(-7, 8, (n, 2))   # fullpel offset is still mv/8 (trunc)
    c["mv_row"] = mv[:, 0]
    c["mv_col"] = mv[:, 1]
    return c, offs
