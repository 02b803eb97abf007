"""Config 0 (tests/small_input.y4m luma) through the oracle: fixture sanity and the invariants the
intra-pruning + transform chain must satisfy.  The CUDA side is compared in test_config0_gpu.py."""
import numpy as np

from tests import config0 as C0
from tests import oracle_lib as O


def oracle_pipeline(luma):
    """-> per (frame, block): satd[13], best mode index, coefficients of the winner's residual"""
    L = O.lib()
    out = []
    for frame in luma:
        pf = C0.padded(frame)
        for x, y in C0.blocks_of(frame):
            src = np.ascontiguousarray(frame[y:y + C0.BS, x:x + C0.BS])
            e = C0.intra_edge(pf, x, y)
            preds, satds = [], []
            for mode, variant, angle in C0.MODES13:
                p = O.predict_intra(mode, variant, e, C0.BS, C0.BS, 8, angle=angle, ief=0, left_len=64,
                                    above_len=64, plane_w=64, plane_h=64, dst_x=x, dst_y=y)
                preds.append(p)
                satds.append(L.orc_get_satd_u8(O.ptr(src), C0.BS, O.ptr(p), C0.BS, C0.BS, C0.BS))
            best = int(np.argmin(satds))
            resid = (src.astype(np.int16) - preds[best].astype(np.int16)).reshape(1, C0.BS, C0.BS)
            coef = O.forward_transform_batch(resid, 3, 0, 8)[0]
            out.append((np.array(satds, np.uint32), best, coef, np.stack(preds)))
    return out


def test_fixture_is_the_reference_clip():
    luma = C0.load_luma()
    assert luma.shape == (5, 64, 64) and luma.dtype == np.uint8
    assert int(luma.sum()) == 2747645                 # printed by make_golden.py when it was extracted
    assert luma.std() > 10                            # real picture content, not a flat test card


def test_oracle_chain_invariants():
    luma = C0.load_luma()
    res = oracle_pipeline(luma)
    assert len(res) == 5 * 4
    modes = set()
    for satds, best, coef, preds in res:
        assert satds[best] == satds.min()
        modes.add(best)
        # DC_PRED of a 32x32 block with both edges = rounded mean of the 64 edge pixels
        assert len(np.unique(preds[0])) == 1
    assert len(modes) >= 3                            # the content exercises several predictors
    # Parseval on the first block (shift triple of TX_32X32 sums to x4)
    satds, best, coef, preds = res[0]
    frame = luma[0]
    resid = frame[:32, :32].astype(np.float64) - preds[best].astype(np.float64)
    e_in, e_out = (resid ** 2).sum(), (coef.astype(np.float64) ** 2).sum()
    assert abs(e_out / max(e_in, 1.0) - 16.0) < 0.6
