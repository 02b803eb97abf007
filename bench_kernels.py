#!/usr/bin/env python
"""bench_kernels.py — per-kernel throughput and roofline for every row of SURVEY §8(a), on the
BASELINE configs that are not the headline one (configs[2..4] shapes).  Not the driver's
contract (that is bench.py); this fills BASELINE.md §5's per-kernel table.

Each leg: device-resident inputs, >=3 warm-up launches, CUDA events on the launching stream,
algorithmic bytes per unit from SURVEY §8(d).  Prints one JSON line per leg.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from rav1e_b200 import backend as B  # noqa: E402
from tests import oracle_lib as O  # noqa: E402  (CPU baseline column only)

PEAK = 6571.6
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
TX_SIZES = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 8), (8, 4), (8, 16), (16, 8), (16, 32),
            (32, 16), (32, 64), (64, 32), (4, 16), (16, 4), (8, 32), (32, 8), (16, 64), (64, 16)]


def timed(fn, reps=20, warm=3):
    """ms per launch, CUDA events.  The GPU idles (and clocks down) while a CPU-port leg runs between two GPU legs:
    warm up until at least 3 launches AND 3 ms of back-to-back work have gone by."""
    import time
    t0, k = time.perf_counter(), 0
    while k < warm or time.perf_counter() - t0 < 3e-3:
        fn()
        k += 1
        if k % 8 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_rate(fn, units):
    """units/s of the oracle (all usable host threads) on the same inputs: ~0.5 s of work."""
    import time
    fn()
    t0, reps = time.perf_counter(), 0
    while time.perf_counter() - t0 < 0.5:
        fn()
        reps += 1
    return units * reps / (time.perf_counter() - t0)


def emit(name, units, unit_name, ms, bytes_per_unit, note="", cpu=None):
    gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
    d = {"kernel": name, "units_per_launch": units, "unit": unit_name, "launch_ms": ms,
         "units_per_s": units / (ms * 1e-3), "algorithmic_bytes_per_unit": bytes_per_unit,
         "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / PEAK, "note": note}
    if cpu is not None:
        d["cpu_port_units_per_s"] = cpu
        d["cpu_threads"] = O.host_threads()
        d["gpu_over_cpu"] = d["units_per_s"] / cpu
    print(json.dumps(d))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()


def grid_blocks(w, h, bw, bh):
    xs, ys = np.arange(0, w - bw + 1, bw), np.arange(0, h - bh + 1, bh)
    b = np.zeros(len(xs) * len(ys), B.BLOCK_DTYPE)
    b["x"], b["y"] = np.tile(xs, len(ys)), np.repeat(ys, len(xs))
    return b


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = B.Context(0, use_torch_stream=True)
    rng = np.random.default_rng(0)
    W, H, PAD = 1920, 1080, 96

    # ---- config 3: forward-transform sweep over a 1080p residual frame, all 160 pairs, 8 bit
    resid = torch.from_numpy(rng.integers(-255, 256, (H * W,)).astype(np.int16)).cuda()
    out = torch.empty(H * W, dtype=torch.int32, device="cuda")
    sweep_ms = 0.0
    for ts, (w, h) in enumerate(TX_SIZES):
        n = (W // w) * (H // h)
        for tt in range(17):
            if not ctx.L.b200_valid_av1_transform(ts, tt):
                continue
            ms = timed(lambda: ctx.fwd_txfm_dev(resid, w * h, w, out, n, ts, tt, 8, False), reps=10)
            sweep_ms += ms
            if tt == 0:
                cpu = None
                if (w, h) in ((16, 16), (4, 4), (64, 64)):
                    hres = resid.cpu().numpy()[:n * w * h].reshape(n, h, w)
                    cpu = cpu_rate(lambda: O.forward_transform_batch(hres, ts, 0, 8, threads=O.host_threads()), n)
                emit(f"fwd_txfm {w}x{h} DCT_DCT 8-bit", n, "blocks", ms, w * h * 4,
                     "1080p residual frame tiled by this size; i16 in + i16 out", cpu)
    print(json.dumps({"kernel": "fwd_txfm sweep (160 size/type pairs x whole 1080p frame)",
                      "total_ms": sweep_ms, "pixels_per_s": 160 * W * H / (sweep_ms * 1e-3)}))

    # ---- planes
    oplanes = {}

    def mk(key, dtype, bd):
        img = rng.integers(0, 1 << bd, (H, W)).astype(dtype)
        op = O.Plane(W, H, PAD, dtype=dtype)
        op.fill_from(img)
        oplanes[key] = op
        return ctx.plane_from_host(img, PAD)
    cur8, ref8 = mk("cur8", np.uint8, 8), mk("ref8", np.uint8, 8)
    cur10, ref10 = mk("cur10", np.uint16, 10), mk("ref10", np.uint16, 10)
    TH = O.host_threads()
    blocks = grid_blocks(W, H, 16, 16)
    nb = len(blocks)
    d_blocks = dev(blocks)

    # ---- mc: put_8tap HV, 16x16, REGULAR, 8 sub-pel vectors per block (config 4 shape)
    pattern = np.array([(r * s, c * s) for s in (4, 2) for r, c in ((1, 0), (0, 1), (-1, 0), (0, -1))])
    centre = rng.integers(-16, 17, (nb, 2)) * 8 + 3
    mvs = (centre[:, None, :] + pattern[None]).reshape(-1, 2).astype(np.int16)
    mblocks = np.repeat(blocks, len(pattern))
    n = len(mblocks)
    d_mb, d_mv = dev(mblocks), dev(mvs)
    for name, ref, oref, bd, bpp in (("8-bit", ref8, "ref8", 8, 1), ("10-bit", ref10, "ref10", 10, 2)):
        d_pred = torch.empty(n * 256 * bpp, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: ctx.mc_blocks_dev(ref, d_mb, d_mv, n, 16, 16, 0, 0, bd, 0, 0, 0, d_pred), reps=10)
        cpu = cpu_rate(lambda: O.mc_blocks(oplanes[oref], mblocks, mvs, 16, 16, 0, 0, bd, threads=TH), n)
        emit(f"put_8tap HV 16x16 {name}", n, "blocks", ms, (23 * 23 + 256) * bpp, cpu=cpu)

    # ---- config 4: sub-pel candidates (MC + SATD + cost + argmin), 10-bit, 8 per block
    cands = np.zeros(n, B.CAND_DTYPE)
    cands["block"] = np.repeat(np.arange(nb, dtype=np.uint32), len(pattern))
    cands["mv_row"], cands["mv_col"] = mvs[:, 0], mvs[:, 1]
    offs = (np.arange(nb + 1) * len(pattern)).astype(np.uint32)
    d_c, d_o = dev(cands), dev(offs)
    d_best = torch.empty(nb * 16, dtype=torch.uint8, device="cuda")
    for name, cur, ref, ocur, oref, bd, bpp in (("8-bit", cur8, ref8, "cur8", "ref8", 8, 1),
                                                ("10-bit", cur10, ref10, "cur10", "ref10", 10, 2)):
        p = B.me_params(16, 16, W, H, 6400, allow_hp=True, use_satd=True, bit_depth=bd)
        ms = timed(lambda: ctx.me_subpel_candidates_dev(cur, ref, d_blocks, nb, d_c, n, p, 0, d_o, None,
                                                        None, None, d_best), reps=10)
        cpu = cpu_rate(lambda: O.subpel_candidates(oplanes[ocur], oplanes[oref], blocks, cands, 16, 16, True,
                                                   6400, None, True, 0, bd, threads=TH), n)
        emit(f"subpel candidates (put_8tap + SATD + cost + argmin) 16x16 {name}", n, "candidates", ms,
             23 * 23 * bpp + 256 * bpp + 4, "two launches + argmin; prediction round-trips through HBM", cpu)
    # chain tail: winners' prediction -> residual -> 16x16 DCT (10-bit)
    d_pred = torch.empty(nb * 256 * 2, dtype=torch.uint8, device="cuda")
    d_coef = torch.empty(nb * 256, dtype=torch.int32, device="cuda")
    d_wmv = dev(mvs[::len(pattern)].copy())

    def chain():
        ctx.mc_blocks_dev(ref10, d_blocks, d_wmv, nb, 16, 16, 0, 0, 10, 0, 0, 0, d_pred)
        ctx.fwd_txfm_pred_dev(cur10, d_pred, d_blocks, nb, d_coef, 2, 0, 10)
    ms = timed(chain, reps=10)
    emit("winner put_8tap + residual + fwd_txfm 16x16 10-bit", nb, "blocks", ms, 2598,
         "SURVEY §8d fused MC->SATD->txfm figure (2598 B/block)")

    # ---- config 4, fused: sub-pel refinement + winner's residual + 16x16 DCT in ONE kernel
    # (b200_subpel_rdo_dev): no prediction touches HBM
    for name, cur, ref, bd, bpp in (("8-bit", cur8, ref8, 8, 1), ("10-bit", cur10, ref10, 10, 2)):
        p = B.me_params(16, 16, W, H, 6400, allow_hp=True, use_satd=True, bit_depth=bd)
        d_fc = torch.empty(nb * 256, dtype=torch.int16 if bd == 8 else torch.int32, device="cuda")
        ms = timed(lambda: ctx.subpel_rdo_dev(cur, ref, d_blocks, nb, d_c, n, d_o, p, 0, 2, 0, None, None, None,
                                              d_best, d_fc), reps=10)
        emit(f"FUSED subpel RDO (8 x [put_8tap + SATD + cost] + argmin + residual + DCT 16x16) {name}", nb, "blocks",
             ms, 8 * 23 * 23 * bpp + 256 * bpp + 256 * (2 if bd == 8 else 4) + 16,
             "BASELINE configs[3]; bytes: 8 source footprints + org block + coefficients + winner "
             "(SURVEY 8d quotes 2598 B per single-candidate block at 10 bit)")

    # ---- lookahead (api/lookahead.rs): pyramid, intra / inter cost maps, importance-block difference
    hres = ctx.plane_alloc(W // 2, H // 2, PAD // 2, 1)
    qres = ctx.plane_alloc(W // 4, H // 4, PAD // 4, 1)

    def pyramid():
        ctx.plane_downsample_dev(cur8, hres, W // 2, H // 2)
        ctx.plane_downsample_dev(hres, qres, W // 4, H // 4)
    ms = timed(pyramid, reps=10)
    emit("Plane::downsampled x2 (half + quarter resolution ME pyramid) 1080p 8-bit", W * H, "source pixels", ms,
         1 + 0.25 + 0.25 + 0.0625, "reads the level above, writes the padded level below")
    nimp = (W // 8) * (H // 8)
    d_ic = torch.empty(nimp, dtype=torch.int32, device="cuda")
    ms = timed(lambda: ctx.estimate_intra_costs_dev(cur8, 8, d_ic), reps=10)
    hic = np.zeros(nimp, np.uint32)
    OLk = O.lib()
    OLk.orc_estimate_intra_costs.restype = None
    OLk.orc_estimate_intra_costs.argtypes = [C.c_void_p, C.c_ssize_t] + [C.c_int] * 4 + [C.c_void_p]
    oc = oplanes["cur8"]
    cpu = cpu_rate(lambda: OLk.orc_estimate_intra_costs(oc.at(0, 0), oc.stride, W, H, 1, 8, hic.ctypes.data), nimp)
    emit("estimate_intra_costs (edges + DC + SATD 8x8, fused) 1080p 8-bit", nimp, "8x8 blocks", ms, 64 + 16 + 4, cpu=cpu)
    d_imv = dev((rng.integers(-32, 33, (nimp, 2)) * 4).astype(np.int16))
    d_scr, d_mean = torch.zeros(1, dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.float64, device="cuda")
    ms = timed(lambda: ctx.estimate_inter_costs_dev(cur8, ref8, d_imv, d_ic, d_scr, d_mean), reps=10)
    emit("estimate_inter_costs cost part (SATD 8x8 vs displaced reference + f64 mean) 1080p 8-bit", nimp, "8x8 blocks",
         ms, 128 + 4 + 4)
    ms = timed(lambda: ctx.importance_block_difference_dev(cur8, ref8, d_scr, d_mean), reps=10)
    emit("estimate_importance_block_difference 1080p 8-bit", nimp, "8x8 blocks", ms, 128)

    # ---- get_intra_edges for every 16x16 block of a frame x 13 modes, then the predictions from them
    eitems = np.zeros(nb * 13, B.EDGE_ITEM_DTYPE)
    eitems["po_x"], eitems["po_y"] = np.repeat(blocks["x"], 13), np.repeat(blocks["y"], 13)
    eitems["part_x"], eitems["part_y"] = eitems["po_x"] // 4, eitems["po_y"] // 4
    eitems["bsize"], eitems["tx_size"] = 6, 2                                  # BLOCK_16X16, TX_16X16
    eitems["mode"] = np.tile([0, 2, 1, 9, 11, 10, 12, 3, 4, 5, 6, 7, 8], nb)   # RAV1E_INTRA_MODES
    eitems["enable_ief"] = 1
    d_ei = dev(eitems)
    d_eb = torch.empty(len(eitems) * 257, dtype=torch.uint8, device="cuda")
    d_el = torch.empty(len(eitems) * 2, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ctx.get_intra_edges_dev(cur8, (0, 0, W, H), 0, 0, 8, d_ei, len(eitems), d_eb, d_el), reps=10)
    emit("get_intra_edges 16x16 x 13 modes 1080p 8-bit", len(eitems), "edge buffers", ms, 257 + 16 + 2,
         "item + up to 65 plane pixels in, 257-pixel buffer + lengths out")

    # ---- compute_rd_cost: 16 candidates per block, cost + first minimum
    nrd = nb * 16
    d_rate = dev(rng.integers(0, 1 << 16, nrd).astype(np.uint32))
    d_dist = dev(rng.integers(0, 1 << 30, nrd).astype(np.uint64))
    d_roffs = dev((np.arange(nb + 1) * 16).astype(np.uint32))
    d_rc = torch.empty(nrd, dtype=torch.float64, device="cuda")
    d_rb = torch.empty(nb, dtype=torch.int32, device="cuda")
    ms = timed(lambda: ctx.compute_rd_cost_dev(117.25, d_rate, d_dist, nrd, d_rc, d_roffs, nb, d_rb), reps=10)
    emit("compute_rd_cost (f64 fma) + first minimum per block, 16 candidates per block", nrd, "candidates", ms, 4 + 8 + 8)

    # ---- device-side search stages of full_pixel_me (me.rs:692-856), 1080p, 16x16 blocks.
    # Structured content (low-passed noise displaced by (+6, -3) px + noise) so the walks are real;
    # predictors scatter within +-8 px of the true motion like neighbouring blocks' vectors do.
    from scipy.ndimage import uniform_filter
    base = uniform_filter(rng.integers(0, 256, (H + 64, W + 64)).astype(np.float32), 9)
    base = (base - base.min()) / (base.max() - base.min()) * 255
    s_ref = np.rint(base[32:32 + H, 32:32 + W]).astype(np.uint8)
    s_cur = np.clip(np.rint(base[32 - 3:32 - 3 + H, 32 + 6:32 + 6 + W] + rng.normal(0, 1.5, (H, W))), 0, 255).astype(np.uint8)
    os_cur, os_ref = O.Plane(W, H, PAD), O.Plane(W, H, PAD)
    os_cur.fill_from(s_cur)
    os_ref.fill_from(s_ref)
    ds_cur, ds_ref = ctx.plane_from_host(s_cur, PAD), ctx.plane_from_host(s_ref, PAD)
    p_srch = B.me_params(16, 16, W, H, 1600)
    d_sbest = torch.empty(nb * 16, dtype=torch.uint8, device="cuda")
    for label, nsub, per, umh_range in (("predictors + diamond (non-extensive)", 1, (10,), 0),
                                        ("extensive ladder through UMH-24 + hexagon", 3, (1, 5, 4), 24)):
        counts = np.tile(np.array(per), nb)
        soffs = np.zeros(nb * nsub + 1, np.uint32)
        soffs[1:] = np.cumsum(counts)
        spreds = np.zeros(int(soffs[-1]), B.CAND_DTYPE)
        smv = (np.array([-3, 6]) + rng.integers(-8, 9, (len(spreds), 2))) * 8
        spreds["mv_row"], spreds["mv_col"] = smv[:, 0], smv[:, 1]
        sthresh = np.zeros(nb, np.uint32)                     # never exit early: every stage runs
        d_sp, d_so, d_st = dev(spreds), dev(soffs), dev(sthresh)
        ms = timed(lambda: ctx.me_search_dev(ds_cur, ds_ref, d_blocks, nb, d_sp, d_so, nsub, p_srch, d_sbest,
                                             None, d_st if nsub == 3 else None, umh_range), reps=10)
        cpu = cpu_rate(lambda: O.full_pixel_me_blocks(os_cur, os_ref, blocks, spreds, soffs, nsub, 16, 16, 1600,
                                                      None, sthresh if nsub == 3 else None, umh_range, threads=TH), nb)
        emit(f"me_search 16x16 8-bit: {label}", nb, "blocks", ms, 0,
             "data-dependent number of SAD evaluations per block (no SURVEY byte figure); one warp per block", cpu)

    # ---- intra: 13 modes per 16x16 block
    edges = rng.integers(0, 256, (nb, 257)).astype(np.uint8)
    modes = [(0, 3, 0), (2, 3, 180), (1, 3, 90), (9, 3, 0), (11, 3, 0), (10, 3, 0), (12, 3, 0), (3, 3, 45),
             (4, 3, 135), (5, 3, 113), (6, 3, 157), (7, 3, 203), (8, 3, 67)]        # RAV1E_INTRA_MODES
    items = np.zeros(nb * len(modes), B.INTRA_ITEM_DTYPE)
    items["edge"] = np.repeat(np.arange(nb, dtype=np.uint32), len(modes))
    items["mode"] = np.tile([m[0] for m in modes], nb)
    items["variant"] = 3
    items["angle"] = np.tile([m[2] for m in modes], nb)
    items["ief"] = 0
    items["left_len"] = items["above_len"] = 32
    items["x"] = np.repeat(blocks["x"], len(modes))
    items["y"] = np.repeat(blocks["y"], len(modes))
    d_e, d_i = dev(edges), dev(items)
    d_p = torch.empty(len(items) * 256, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ctx.predict_intra_dev(d_e, d_i, len(items), None, 16, 16, 8, W, H, d_p), reps=10)
    hout = np.zeros((len(items), 256), np.uint8)
    OL = O.lib()
    OL.orc_predict_intra_batch.restype = None
    OL.orc_predict_intra_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
    cpu = cpu_rate(lambda: OL.orc_predict_intra_batch(edges.ctypes.data, 1, items.ctypes.data, len(items), None,
                                                      16, 16, 8, W, H, hout.ctypes.data), len(items))
    emit("predict_intra 16x16 x 13 modes 8-bit", len(items), "predictions", ms, 257 + 256, cpu=cpu)

    # ---- quantize -> dequantize -> tx-domain distortion on the 16x16 DCT coefficients of a frame
    d_coef16 = torch.empty(nb * 256, dtype=torch.int16, device="cuda")
    ctx.fwd_txfm_residual_dev(cur8, ref8, d_blocks, nb, None, d_coef16, 2, 0, 8)
    d_q16 = torch.empty(nb * 256, dtype=torch.int16, device="cuda")
    d_r16 = torch.empty(nb * 256, dtype=torch.int16, device="cuda")
    d_eob = torch.empty(nb, dtype=torch.int16, device="cuda")
    d_txd = torch.empty(nb, dtype=torch.int64, device="cuda")
    ms = timed(lambda: ctx.quantize_dev(d_coef16, nb, 2, 0, 120, 96, False, False, d_q16, d_r16, d_eob, d_txd), reps=10)
    hcoef = d_coef16.cpu().numpy().reshape(nb, 256)
    hq, hr = np.zeros_like(hcoef), np.zeros_like(hcoef)
    he, hd = np.zeros(nb, np.uint16), np.zeros(nb, np.uint64)
    OLq = O.lib()
    OLq.orc_quantize_chain_batch.restype = None
    OLq.orc_quantize_chain_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                             C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    cpu = cpu_rate(lambda: OLq.orc_quantize_chain_batch(hcoef.ctypes.data, nb, 2, 0, 120, 96, 0, 0, hq.ctypes.data,
                                                        hr.ctypes.data, he.ctypes.data, hd.ctypes.data, TH), nb)
    emit("quantize + dequantize + tx-dist 16x16 8-bit (noise residual: long scans)", nb, "blocks", ms,
         3 * 512 + 10, "coefficients in, qcoeffs + rcoeffs + eob + distortion out", cpu)

    # ---- inverse transform + reconstruct, and encode_tx_block's numeric core in one call
    rec = ctx.plane_from_host(np.zeros((H, W), np.uint8) + 128, 0)
    ms = timed(lambda: ctx.inverse_transform_add_dev(d_r16, rec, d_blocks, nb, 2, 0, 8), reps=10)
    emit("inverse_transform_add 16x16 DCT_DCT 8-bit", nb, "blocks", ms, 512 + 2 * 256, "rcoeffs in, plane read-modify-write")
    d_co2 = torch.empty(nb * 256, dtype=torch.int16, device="cuda")
    ms = timed(lambda: ctx.encode_tx_blocks_dev(cur8, ref8, rec, d_blocks, nb, None, 2, 0, 8, 120, 96, False, True,
                                                d_co2, d_q16, d_r16, d_eob, d_txd), reps=10)
    emit("encode_tx_blocks (diff + fwd txfm + quantize + dequantize + tx-dist + inverse add) 16x16 8-bit", nb, "blocks",
         ms, 2 * 256 + 3 * 512 + 10 + 2 * 256, "")

    # ---- RDO distortion kernels on the 1080p 8-bit planes: every 16x16 block / every 8x8 block
    d_scale = dev(rng.integers(1 << 12, 1 << 16, ((H + 3) // 4, W // 4)).astype(np.uint32))
    d_wsse = torch.empty(nb * 8, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ctx.weighted_sse_dev(cur8, ref8, d_blocks, nb, 16, 16, d_scale, W // 4, d_wsse), reps=10)
    emit("get_weighted_sse 16x16 8-bit", nb, "blocks", ms, 2 * 256 + 16 * 4 + 8,
         "two 16x16 blocks + 16 chunk scales in, one u64 out")
    b8 = grid_blocks(W, H, 8, 8)
    d_b8 = dev(b8)
    d_cd = torch.empty(len(b8) * 4, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ctx.cdef_dist_dev(cur8, ref8, d_b8, len(b8), 8, 8, 8, d_cd, None), reps=10)
    emit("cdef_dist_kernel 8x8 8-bit (+ ssim boost)", len(b8), "8x8 blocks", ms, 2 * 64 + 4)

    # ---- config 5: CDEF on a 4K frame (luma + one 4:2:0 chroma plane)
    W4, H4 = 3840, 2160
    luma = rng.integers(0, 256, (H4, W4)).astype(np.uint8)
    chroma = rng.integers(0, 256, (H4 // 2, W4 // 2)).astype(np.uint8)
    pl, plo = ctx.plane_from_host(luma, 0), ctx.plane_from_host(np.zeros_like(luma), 0)
    pc, pco = ctx.plane_from_host(chroma, 0), ctx.plane_from_host(np.zeros_like(chroma), 0)
    n8 = (W4 // 8) * (H4 // 8)
    d_dir = torch.empty(n8, dtype=torch.uint8, device="cuda")
    d_var = torch.empty(n8, dtype=torch.int32, device="cuda")
    d_str = dev(np.full(((H4 + 63) // 64) * ((W4 + 63) // 64), 4 * 7 + 2, np.uint8))
    ms = timed(lambda: ctx.cdef_find_dir_dev(pl, 8, None, d_dir, d_var), reps=10)
    cpu = cpu_rate(lambda: O.cdef_analyze_frame(luma, 8), n8)
    emit("cdef_find_dir 4K luma", n8, "8x8 blocks", ms, 64 + 8, cpu=cpu)
    hdirs, hvar = O.cdef_analyze_frame(luma, 8)
    hstr = np.full(((H4 + 63) // 64, (W4 + 63) // 64), 4 * 7 + 2, np.uint8)
    ms = timed(lambda: ctx.cdef_filter_plane_dev(pl, plo, 0, 0, 0, W4, H4, 8, 5, None, d_dir, d_var, d_str), reps=10)
    cpu = cpu_rate(lambda: O.cdef_filter_plane(luma, 0, 0, 0, W4, H4, 8, 5, None, hdirs, hvar, hstr), n8)
    emit("cdef_filter 4K luma", n8, "8x8 blocks", ms, 208, cpu=cpu)
    ms = timed(lambda: ctx.cdef_filter_plane_dev(pc, pco, 1, 1, 1, W4, H4, 8, 5, None, d_dir, d_var, d_str), reps=10)
    emit("cdef_filter 4K chroma 4:2:0", n8, "4x4 blocks", ms, 8 * 8 + 16)


if __name__ == "__main__":
    main()
