"""BASELINE config 0 on the GPU: the speed-10 intra shape of the path over the luma of
tests/small_input.y4m (tests/config0.py) — CUDA through the C ABI == oracle, bit exact, at every
step: 13 predictions per 32x32 block, their SATD against the source (one launch over 13 plane
pairs), and the 32x32 DCT_DCT of the winner's residual."""
import numpy as np
import pytest

from rav1e_b200 import backend as B
from tests import config0 as C0
from tests import gpu_util as G
from tests.test_config0_oracle import oracle_pipeline

pytestmark = pytest.mark.gpu


def test_config0_intra_chain_matches_oracle():
    import torch
    luma = C0.load_luma()
    want = oracle_pipeline(luma)
    c = G.ctx()
    nm = len(C0.MODES13)
    k = 0
    for frame in luma:
        pf = C0.padded(frame)
        blks = C0.blocks_of(frame)
        nb = len(blks)
        # ---- 13 modes x 4 blocks in one predict launch
        edges = np.stack([C0.intra_edge(pf, x, y) for x, y in blks])
        items = np.zeros(nb * nm, B.INTRA_ITEM_DTYPE)
        for b, (x, y) in enumerate(blks):
            for m, (mode, variant, angle) in enumerate(C0.MODES13):
                items[b * nm + m] = (b, 0, x, y, angle, mode, variant, 0, 64, 64, 0)
        d_pred = torch.empty((nb * nm, C0.BS, C0.BS), dtype=torch.uint8, device="cuda")
        c.predict_intra_dev(G.to_dev(edges), G.to_dev(items), len(items), None, C0.BS, C0.BS, 8, 64, 64, d_pred)
        c.synchronize()
        preds = d_pred.cpu().numpy()
        for b in range(nb):
            np.testing.assert_array_equal(preds[b * nm:(b + 1) * nm], want[k + b][3])
        # ---- SATD of every prediction against the source: one launch over 13 (cur, ref) pairs,
        # ref plane m = mode m's predictions tiled at their block positions, zero motion
        dcur = c.plane_from_host(frame, C0.PAD)
        mode_planes = []
        for m in range(nm):
            img = np.zeros_like(frame)
            for b, (x, y) in enumerate(blks):
                img[y:y + C0.BS, x:x + C0.BS] = preds[b * nm + m]
            mode_planes.append(c.plane_from_host(img, C0.PAD))
        blocks = np.zeros(nb, B.BLOCK_DTYPE)
        blocks["x"], blocks["y"] = [x for x, _ in blks], [y for _, y in blks]
        all_blocks = np.tile(blocks, nm)
        cands = np.zeros(nb * nm, B.CAND_DTYPE)
        cands["block"] = np.arange(nb * nm)
        offs = np.arange(nb * nm + 1, dtype=np.uint32)
        ends = [(m + 1) * nb for m in range(nm)]
        d_satd = G.dev_empty(4 * nb * nm)
        p = B.me_params(C0.BS, C0.BS, 64, 64, 0, use_satd=True)
        c.me_candidates_multi_dev(B.PlanePairs([dcur] * nm, mode_planes, ends, ends), G.to_dev(all_blocks), nb * nm,
                                  G.to_dev(cands), nb * nm, p, G.to_dev(offs), None, d_satd, None, None)
        c.synchronize()
        satd = G.from_dev(d_satd, np.uint32)[:nb * nm].reshape(nm, nb).T          # [block][mode]
        best = []
        for b in range(nb):
            np.testing.assert_array_equal(satd[b], want[k + b][0])
            best.append(int(np.argmin(satd[b])))
            assert best[-1] == want[k + b][1]
        # ---- winner's residual -> TX_32X32 DCT_DCT (packed-prediction form)
        win = np.stack([preds[b * nm + best[b]] for b in range(nb)])
        d_coef = torch.empty((nb, C0.BS * C0.BS), dtype=torch.int16, device="cuda")
        c.fwd_txfm_pred_dev(dcur, G.to_dev(win), G.to_dev(blocks), nb, d_coef, 3, 0, 8)
        c.synchronize()
        coef = d_coef.cpu().numpy()
        for b in range(nb):
            np.testing.assert_array_equal(coef[b], want[k + b][2].reshape(-1))
        for pl in [dcur] + mode_planes:
            c.plane_free(pl)
        k += nb
