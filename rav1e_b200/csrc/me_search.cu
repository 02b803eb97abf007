// me_search.cu — the move-to-best search stages of rav1e's full_pixel_me on the device.
//
// Replaces, for every block of a frame (or of several (cur, ref) plane pairs) in one launch, the
// serial per-block chain of src/me.rs:
//   full_pixel_me            :692-856   (the `try_cands` closure, the extensive-search ladder)
//   get_best_predictor       :884-909
//   fullpel_diamond_search   :955-998   (4-point diamond, radius 2 then 1, move until no gain)
//   hexagon_search           :1055-1135 (6 / 3-point hexagon, then the 8-point square)
//   uneven_multi_hex_search  :1170-1303 (cross 24 + 12, 5x5, 16-point hexagons x 6 scales)
// The exhaustive grid that may follow (:822-846) is b200_me_full_search_dev.
//
// One WARP per block.  A stage's candidate set is evaluated four candidates at a time (8 lanes
// each, lane = row mod 8), straight from the planes through L1 (neighbouring candidates overlap
// almost entirely); every "if rd.cost < best.cost" chain of the reference is a first-minimum
// argmin over the set, so the winner is found with shuffles and the data-dependent control flow
// (move, shrink the radius, early exits) stays warp-uniform.  Predictor subsets come from the
// caller: get_subset_predictors (me.rs:381-533) reads neighbouring blocks' results of the same
// pass, a wavefront dependency that stays with the encoder's block order.
#include <cuda_runtime.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr unsigned long long kEmptyCost = ~0ull;  // MVCandidateRD::empty(), me.rs:139-146
constexpr uint32_t kEmptySad = ~0u;
constexpr int kMaxPairs = 32;

struct PlaneRef {
  const void *data;  // pixel (0,0)
  int stride;        // elements
};

struct SearchArgs {
  int npairs;
  uint32_t block_end[kMaxPairs];  // pair k owns blocks [block_end[k-1], block_end[k])
  PlaneRef cur[kMaxPairs], ref[kMaxPairs];
  const b200_block *blocks;
  const b200_cand *preds;
  const uint32_t *subset_offsets;  // nblocks * nsubsets + 1
  const short *pmv;                // 4 shorts per block or null
  const uint32_t *thresh;          // per block (extensive ladder) or null
  b200_me_result *out;
  size_t nblocks;
  int nsubsets;  // 1: all_mvs (non-extensive); 3: median | subset_b | subset_c
  int w, h, w_in_b, h_in_b;
  uint32_t lambda;
  int allow_hp;
  int umh_range;
};

struct Best {  // MotionSearchResult (+ the index of the winner inside its candidate set)
  unsigned long long cost;
  uint32_t sad;
  int row, col;  // i16 values
  int idx;
};

__device__ __forceinline__ Best best_empty() {
  Best b;
  b.cost = kEmptyCost;
  b.sad = kEmptySad;
  b.row = 0;
  b.col = 0;
  b.idx = 0x7fffffff;
  return b;
}

__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// search_pattern! tables (me.rs:944-947, :1022-1025, :1035-1038, :1153-1156), full pixels
__constant__ signed char kDiamond[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};  // {row, col}
__constant__ signed char kHexagon[6][2] = {{-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}};
__constant__ signed char kSquare[8][2] = {{1, -1}, {1, 0}, {1, 1}, {0, -1}, {0, 1}, {-1, -1}, {-1, 0}, {-1, 1}};
__constant__ signed char kUmh[16][2] = {{4, -2}, {4, -1}, {4, 0},  {4, 1},   {4, 2},   {2, 3},  {0, 4},  {-2, 3},
                                        {-4, 2}, {-4, 1}, {-4, 0}, {-4, -1}, {-4, -2}, {-2, 3}, {0, -4}, {2, -3}};

struct BlockCtx {
  const uint8_t *org;  // block pixel (0,0), bytes
  long long org_pitch;  // bytes
  const uint8_t *ref0;
  long long ref_pitch;  // bytes
  int po_x, po_y;
  MvRange rng;
  int p0r, p0c, p1r, p1c;
  int w, h;
  uint32_t lambda;
  int allow_hp;
};

// sum |org - ref| over rows sub, sub+8, ... of the block displaced by (dx, dy) full pixels
template <typename T>
__device__ __forceinline__ uint32_t partial_sad(const BlockCtx &c, int dx, int dy, int sub) {
  uint32_t acc = 0;
  if (sizeof(T) == 1) {
    const uint8_t *r0 = c.ref0 + (long long)(c.po_y + dy) * c.ref_pitch + (c.po_x + dx);
    const int rsh = (int)((uintptr_t)r0 & 3), osh = (int)((uintptr_t)c.org & 3);
    // pitches are multiples of 4 bytes (checked on the host), so the shifts hold for every row
    const uint32_t *rw = (const uint32_t *)(r0 - rsh) + (long long)sub * (c.ref_pitch >> 2);
    const uint32_t *ow = (const uint32_t *)(c.org - osh) + (long long)sub * (c.org_pitch >> 2);
    const int nw = c.w >> 2;
    for (int y = sub; y < c.h; y += 8) {
      uint32_t rlo = __ldg(rw), olo = __ldg(ow);
      for (int k = 0; k < nw; k++) {
        const uint32_t rhi = __ldg(rw + k + 1), ohi = __ldg(ow + k + 1);
        acc = sad4_acc(__funnelshift_r(rlo, rhi, rsh * 8), __funnelshift_r(olo, ohi, osh * 8), acc);
        rlo = rhi;
        olo = ohi;
      }
      rw += 2 * c.ref_pitch;  // 8 rows, in words
      ow += 2 * c.org_pitch;
    }
  } else {
    const uint16_t *r0 = (const uint16_t *)(c.ref0 + (long long)(c.po_y + dy) * c.ref_pitch) + (c.po_x + dx);
    const uint16_t *o0 = (const uint16_t *)c.org;
    for (int y = sub; y < c.h; y += 8) {
      const uint16_t *rr = (const uint16_t *)((const uint8_t *)r0 + (long long)y * c.ref_pitch);
      const uint16_t *oo = (const uint16_t *)((const uint8_t *)o0 + (long long)y * c.org_pitch);
      for (int x = 0; x < c.w; x++) {
        const int d = (int)oo[x] - (int)rr[x];
        acc += (uint32_t)(d < 0 ? -d : d);
      }
    }
  }
  return acc;
}

// First-minimum argmin of get_fullpel_mv_rd (me.rs:1386-1409) over candidates gen(0..n): what
// `for cand { if rd.cost < best.cost { best = rd } }` starting from empty() leaves in `best`.
// Every lane returns the same result.
template <typename T, typename Gen>
__device__ __forceinline__ Best set_best(const BlockCtx &c, int n, Gen gen) {
  const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
  Best b = best_empty();
  for (int k0 = 0; k0 < n; k0 += 4) {
    const int k = k0 + g;
    int row = 0, col = 0;
    bool inr = false;
    if (k < n) {
      gen(k, row, col);
      inr = !(col < c.rng.x_min || col > c.rng.x_max || row < c.rng.y_min || row > c.rng.y_max);
    }
    uint32_t acc = 0;
    if (inr) acc = partial_sad<T>(c, col / 8, row / 8, sub);  // trunc toward zero, me.rs:1402-1403
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (inr) {
      const unsigned long long cost =
          b200_mv_cost(acc, row, col, c.p0r, c.p0c, c.p1r, c.p1c, c.lambda, c.allow_hp);
      if (cost < b.cost) {
        b.cost = cost;
        b.sad = acc;
        b.row = row;
        b.col = col;
        b.idx = k;
      }
    }
  }
  // the four groups: lowest cost, ties to the lowest index (the serial scan's strict `<`)
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    Best q;
    q.cost = __shfl_xor_sync(0xffffffffu, b.cost, o);
    q.sad = __shfl_xor_sync(0xffffffffu, b.sad, o);
    q.row = __shfl_xor_sync(0xffffffffu, b.row, o);
    q.col = __shfl_xor_sync(0xffffffffu, b.col, o);
    q.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
    if (q.cost < b.cost || (q.cost == b.cost && q.idx < b.idx)) b = q;
  }
  return b;
}

__device__ __forceinline__ int s16(int v) { return (int)(short)v; }  // i16 wrap (release builds)

// me.rs:955-998
template <typename T>
__device__ __forceinline__ void fullpel_diamond(const BlockCtx &c, Best &cur) {
  int radius_log2 = 1;
  for (;;) {
    const int cr = cur.row, cc = cur.col, sh = radius_log2 + 3;
    const Best b = set_best<T>(c, 4, [&](int k, int &row, int &col) {
      row = s16(cr + s16(kDiamond[k][0] * (1 << sh)));
      col = s16(cc + s16(kDiamond[k][1] * (1 << sh)));
    });
    if (cur.cost <= b.cost) {
      if (radius_log2 == 0) break;
      radius_log2--;
    } else {
      cur = b;
    }
  }
}

// me.rs:1055-1135
template <typename T>
__device__ __forceinline__ void hexagon(const BlockCtx &c, Best &cur) {
  Best b = set_best<T>(c, 6, [&](int k, int &row, int &col) {
    row = s16(cur.row + kHexagon[k][0] * 8);
    col = s16(cur.col + kHexagon[k][1] * 8);
  });
  int best_idx = b.cost != kEmptyCost ? b.idx : 0;
  while (b.cost < cur.cost) {
    cur = b;
    const int center_idx = best_idx;
    b = set_best<T>(c, 3, [&](int k, int &row, int &col) {
      const int i = (center_idx + 5 + k) % 6;
      row = s16(cur.row + kHexagon[i][0] * 8);
      col = s16(cur.col + kHexagon[i][1] * 8);
    });
    if (b.cost != kEmptyCost) best_idx = (center_idx + 5 + b.idx) % 6;
  }
  b = set_best<T>(c, 8, [&](int k, int &row, int &col) {
    row = s16(cur.row + kSquare[k][0] * 8);
    col = s16(cur.col + kSquare[k][1] * 8);
  });
  if (b.cost < cur.cost) cur = b;
}

// me.rs:1170-1303 (quirks kept: the first line of the cross steps the ROW component, :1195-1198;
// the 5x5 stage adds raw {row, col} eighth-pels, :1240-1246)
template <typename T>
__device__ __forceinline__ void umh(const BlockCtx &c, Best &cur, int me_range) {
  int cr = cur.row, cc = cur.col;
  const int n_row = (me_range + 1) >> 1;         // i = 1, 3, ... <= me_range
  const int n_col = ((me_range >> 1) + 1) >> 1;  // i = 1, 3, ... <= me_range / 2
  Best b = set_best<T>(c, 2 * (n_row + n_col), [&](int k, int &row, int &col) {
    row = cr;
    col = cc;
    if (k < 2 * n_row) {
      const int i = 1 + 2 * (k >> 1);
      row = s16(cr + s16(((k & 1) ? 8 : -8) * i));
    } else {
      const int kk = k - 2 * n_row, i = 1 + 2 * (kk >> 1);
      col = s16(cc + s16(((kk & 1) ? 8 : -8) * i));
    }
  });
  if (b.cost < cur.cost) cur = b;
  cr = cur.row, cc = cur.col;
  b = set_best<T>(c, 24, [&](int k, int &row, int &col) {
    const int idx = k < 12 ? k : k + 1;  // row-major 5x5 without its centre
    row = s16(cr + idx / 5 - 2);
    col = s16(cc + idx % 5 - 2);
  });
  if (b.cost < cur.cost) cur = b;
  cr = cur.row, cc = cur.col;
  b = set_best<T>(c, 16 * (me_range >> 2), [&](int k, int &row, int &col) {
    const int i = 1 + (k >> 4), p = k & 15;
    row = s16(cr + s16(kUmh[p][0] * 8 * i));
    col = s16(cc + s16(kUmh[p][1] * 8 * i));
  });
  if (b.cost < cur.cost) cur = b;
  hexagon<T>(c, cur);
}

template <typename T>
__global__ void __launch_bounds__(128) me_search_kernel(const __grid_constant__ SearchArgs a) {
  const size_t warp0 = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
  const int lane = threadIdx.x & 31;
  for (size_t blk = warp0; blk < a.nblocks; blk += nwarps) {
    int pi = 0;
    if (a.npairs > 1) {  // first pair with blk < block_end
      int lo = 0, hi = a.npairs - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)blk < a.block_end[mid])
          hi = mid;
        else
          lo = mid + 1;
      }
      pi = lo;
    }
    const b200_block bk = a.blocks[blk];
    BlockCtx c;
    c.org_pitch = (long long)a.cur[pi].stride * (long long)sizeof(T);
    c.ref_pitch = (long long)a.ref[pi].stride * (long long)sizeof(T);
    c.org = (const uint8_t *)a.cur[pi].data + (long long)bk.y * c.org_pitch + (long long)bk.x * (long long)sizeof(T);
    c.ref0 = (const uint8_t *)a.ref[pi].data;
    c.po_x = bk.x;
    c.po_y = bk.y;
    c.w = a.w;
    c.h = a.h;
    c.rng = b200_mv_range(a.w_in_b, a.h_in_b, bk.x / MI_SIZE, bk.y / MI_SIZE, a.w, a.h);
    c.p0r = c.p0c = c.p1r = c.p1c = 0;
    if (a.pmv) {
      const short *p = a.pmv + 4 * blk;
      c.p0r = p[0], c.p0c = p[1], c.p1r = p[2], c.p1c = p[3];
    }
    c.lambda = a.lambda;
    c.allow_hp = a.allow_hp;

    Best best = best_empty();
    bool done = false;
    for (int k = 0; k < a.nsubsets && !done; k++) {
      const uint32_t lo = a.subset_offsets[blk * a.nsubsets + k];
      const uint32_t hi = a.subset_offsets[blk * a.nsubsets + k + 1];
      if (a.nsubsets == 3 && k == 0 && lo == hi) continue;  // `if let Some(median)`, me.rs:774
      const b200_cand *pp = a.preds + lo;
      Best r = set_best<T>(c, (int)(hi - lo), [&](int j, int &row, int &col) {
        const b200_cand q = pp[j];
        row = q.mv_row;
        col = q.mv_col;
      });
      fullpel_diamond<T>(c, r);
      if (r.cost < best.cost) best = r;
      if (a.nsubsets == 3 && best.sad < a.thresh[blk]) done = true;  // me.rs:777-790
    }
    if (a.nsubsets == 3 && !done && a.umh_range > 0 && best.cost != kEmptyCost) umh<T>(c, best, a.umh_range);
    if (lane == 0) {
      b200_me_result res;
      res.cost = best.cost;
      res.sad = best.sad;
      res.mv_row = (int16_t)best.row;
      res.mv_col = (int16_t)best.col;
      a.out[blk] = res;
    }
  }
}

int search_pairs(b200_ctx *ctx, size_t npairs, const b200_plane *curs, const b200_plane *refs,
                 const uint32_t *block_end, const b200_block *d_blocks, size_t nblocks,
                 const b200_cand *d_preds, const uint32_t *d_subset_offsets, int nsubsets,
                 const int16_t *d_pmv, const uint32_t *d_thresh, const b200_me_params *p, int umh_range,
                 b200_me_result *d_best) {
  B200_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
  B200_REQUIRE(ctx, p != nullptr && npairs >= 1 && curs && refs && block_end, "NULL params / plane pair table");
  B200_REQUIRE(ctx, nsubsets == 1 || nsubsets == 3, "nsubsets must be 1 (all_mvs) or 3 (median|b|c), got %d", nsubsets);
  B200_REQUIRE(ctx, nsubsets == 1 || d_thresh != nullptr, "the extensive ladder needs per-block thresholds");
  B200_REQUIRE(ctx, p->w > 0 && p->h > 0 && p->w <= 128 && p->h <= 128 && (p->w & 3) == 0,
               "block size %dx%d out of range (<= 128, width multiple of 4)", p->w, p->h);
  B200_REQUIRE(ctx, umh_range >= 0 && umh_range <= 64, "umh_range %d out of range", umh_range);
  B200_REQUIRE(ctx, nblocks < (1ull << 32), "nblocks must fit 32 bits");
  for (size_t k = 0; k < npairs; k++) {
    B200_REQUIRE(ctx, curs[k].data && refs[k].data, "pair %zu: plane has no device memory", k);
    B200_REQUIRE(ctx, curs[k].bpp == refs[k].bpp && curs[k].bpp == curs[0].bpp && (curs[0].bpp == 1 || curs[0].bpp == 2),
                 "plane pairs must share bpp (1 or 2)");
    B200_REQUIRE(ctx, curs[0].bpp == 2 || ((curs[k].stride & 3) == 0 && (refs[k].stride & 3) == 0),
                 "8-bit planes need a row pitch that is a multiple of 4");
    B200_REQUIRE(ctx, block_end[k] >= (k ? block_end[k - 1] : 0), "pair %zu: block ends must be non-decreasing", k);
  }
  B200_REQUIRE(ctx, block_end[npairs - 1] == nblocks, "last pair must end at nblocks");
  if (nblocks == 0) return B200_OK;
  B200_REQUIRE(ctx, d_blocks && d_subset_offsets && d_best, "NULL blocks / offsets / output");
  B200_CUDA(ctx, cudaSetDevice(ctx->device));
  for (size_t k0 = 0; k0 < npairs; k0 += kMaxPairs) {
    const int n = (int)std::min<size_t>(kMaxPairs, npairs - k0);
    const uint32_t b0 = k0 ? block_end[k0 - 1] : 0;
    SearchArgs a{};
    a.npairs = n;
    for (int k = 0; k < n; k++) {
      a.block_end[k] = block_end[k0 + k] - b0;
      a.cur[k] = {curs[k0 + k].data, curs[k0 + k].stride};
      a.ref[k] = {refs[k0 + k].data, refs[k0 + k].stride};
    }
    a.nblocks = a.block_end[n - 1];
    if (a.nblocks == 0) continue;
    a.blocks = d_blocks + b0;
    a.preds = d_preds;                                          // offsets are global
    a.subset_offsets = d_subset_offsets + (size_t)b0 * nsubsets;
    a.pmv = d_pmv ? d_pmv + 4 * (size_t)b0 : nullptr;
    a.thresh = d_thresh ? d_thresh + b0 : nullptr;
    a.out = d_best + b0;
    a.nsubsets = nsubsets;
    a.w = p->w;
    a.h = p->h;
    a.w_in_b = p->frame_w_in_b;
    a.h_in_b = p->frame_h_in_b;
    a.lambda = p->lambda;
    a.allow_hp = p->allow_high_precision_mv;
    a.umh_range = umh_range;
    const int wpc = 4;
    const int grid = (int)std::min<size_t>((a.nblocks + wpc - 1) / wpc, (size_t)ctx->num_sms * 64);
    if (curs[0].bpp == 1)
      me_search_kernel<uint8_t><<<grid, wpc * 32, 0, ctx->stream>>>(a);
    else
      me_search_kernel<uint16_t><<<grid, wpc * 32, 0, ctx->stream>>>(a);
    B200_LAUNCH_CHECK(ctx);
  }
  return B200_OK;
}

}  // namespace

extern "C" int b200_me_search_dev(b200_ctx *ctx, const b200_plane *cur, const b200_plane *ref,
                                  const b200_block *d_blocks, size_t nblocks, const b200_cand *d_preds,
                                  const uint32_t *d_subset_offsets, int nsubsets, const int16_t *d_pmv,
                                  const uint32_t *d_thresh, const b200_me_params *params, int umh_range,
                                  b200_me_result *d_best) {
  B200_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
  B200_REQUIRE(ctx, cur && ref && nblocks < (1ull << 32), "NULL planes / nblocks must fit 32 bits");
  const uint32_t be = (uint32_t)nblocks;
  return search_pairs(ctx, 1, cur, ref, &be, d_blocks, nblocks, d_preds, d_subset_offsets, nsubsets, d_pmv,
                      d_thresh, params, umh_range, d_best);
}

extern "C" int b200_me_search_multi_dev(b200_ctx *ctx, size_t npairs, const b200_plane *curs,
                                        const b200_plane *refs, const uint32_t *pair_block_end,
                                        const b200_block *d_blocks, size_t nblocks,
                                        const b200_cand *d_preds, const uint32_t *d_subset_offsets,
                                        int nsubsets, const int16_t *d_pmv, const uint32_t *d_thresh,
                                        const b200_me_params *params, int umh_range,
                                        b200_me_result *d_best) {
  return search_pairs(ctx, npairs, curs, refs, pair_block_end, d_blocks, nblocks, d_preds, d_subset_offsets,
                      nsubsets, d_pmv, d_thresh, params, umh_range, d_best);
}
