#!/usr/bin/env python
"""tools/search_bench.py — times b200_me_search_dev / _multi_dev alone (kernel experiments).
Same inputs as the me_search legs of bench_kernels.py; prints one JSON line per configuration."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rav1e_b200 import backend as B  # noqa: E402
from bench_kernels import dev, grid_blocks, timed  # noqa: E402


def main():
    from scipy.ndimage import uniform_filter
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = B.Context(0, use_torch_stream=True)
    rng = np.random.default_rng(0)
    W, H, PAD, NP = 1920, 1080, 96, 8
    base = uniform_filter(rng.integers(0, 256, (H + 64, W + 64)).astype(np.float32), 9)
    base = (base - base.min()) / (base.max() - base.min()) * 255
    ref = np.rint(base[32:32 + H, 32:32 + W]).astype(np.uint8)
    cur = np.clip(np.rint(base[29:29 + H, 38:38 + W] + rng.normal(0, 1.5, (H, W))), 0, 255).astype(np.uint8)
    planes = [(ctx.plane_from_host(cur, PAD), ctx.plane_from_host(ref, PAD)) for _ in range(NP)]
    blocks = grid_blocks(W, H, 16, 16)
    nb = len(blocks)
    p = B.me_params(16, 16, W, H, 1600)
    for label, nsub, per, umh in (("non-extensive", 1, (10,), 0), ("extensive+umh24", 3, (1, 5, 4), 24)):
        counts = np.tile(np.array(per), nb * NP)
        offs = np.zeros(nb * NP * nsub + 1, np.uint32)
        offs[1:] = np.cumsum(counts)
        preds = np.zeros(int(offs[-1]), B.CAND_DTYPE)
        mv = (np.array([-3, 6]) + rng.integers(-8, 9, (len(preds), 2))) * 8
        preds["mv_row"], preds["mv_col"] = mv[:, 0], mv[:, 1]
        d_p, d_o = dev(preds), dev(offs)
        d_t = dev(np.zeros(nb * NP, np.uint32))
        d_b = dev(np.tile(blocks, NP))
        d_out = torch.empty(nb * NP * 16, dtype=torch.uint8, device="cuda")
        one = timed(lambda: ctx.me_search_dev(planes[0][0], planes[0][1], d_b, nb, d_p, d_o, nsub, p, d_out, None,
                                              d_t if nsub == 3 else None, umh), reps=10)
        pairs = B.PlanePairs([a for a, _ in planes], [b for _, b in planes], [(k + 1) * nb for k in range(NP)],
                             [(k + 1) * nb for k in range(NP)])
        multi = timed(lambda: ctx.me_search_multi_dev(pairs, d_b, nb * NP, d_p, d_o, nsub, p, d_out, None,
                                                      d_t if nsub == 3 else None, umh), reps=10)
        print(json.dumps({"config": label, "one_pair_ms": one, "blocks_per_s_one": nb / (one * 1e-3),
                          f"{NP}_pairs_ms": multi, "blocks_per_s_multi": nb * NP / (multi * 1e-3)}))


if __name__ == "__main__":
    main()
