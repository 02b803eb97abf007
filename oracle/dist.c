/*
 * oracle/dist.c — restatement of rav1e src/dist.rs `rust::get_sad` / `rust::get_satd`.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Pinned by the reference KATs
 * src/dist.rs:418-441 (SAD) and :477-500 (SATD) — tests/test_oracle_dist.py.
 */
#include "oracle.h"

#include <stdlib.h>

/* v_frame::math::msb — 31 - clz(x) for x > 0 (pinned via the SATD KATs). */
static inline int orc_msb(int32_t x) { return 31 - __builtin_clz((uint32_t)x); }

/* dist.rs:55-57 */
#define BUTTERFLY(a, b, s, d) \
  do {                        \
    int32_t _a = (a), _b = (b); \
    (s) = _a + _b;            \
    (d) = _a - _b;            \
  } while (0)

/* dist.rs:60-82 hadamard4_1d<LEN,N,STRIDE0,STRIDE1> */
static void hadamard4_1d(int32_t *data, int n, int stride0, int stride1) {
  for (int i = 0; i < n; i++) {
    int32_t *sub = data + i * stride0;
    int32_t a0, a1, a2, a3, b0, b1, b2, b3;
    BUTTERFLY(sub[0 * stride1], sub[1 * stride1], a0, a1);
    BUTTERFLY(sub[2 * stride1], sub[3 * stride1], a2, a3);
    BUTTERFLY(a0, a2, b0, b2);
    BUTTERFLY(a1, a3, b1, b3);
    sub[0 * stride1] = b0;
    sub[1 * stride1] = b1;
    sub[2 * stride1] = b2;
    sub[3 * stride1] = b3;
  }
}

/* dist.rs:84-123 hadamard8_1d */
static void hadamard8_1d(int32_t *data, int n, int stride0, int stride1) {
  for (int i = 0; i < n; i++) {
    int32_t *sub = data + i * stride0;
    int32_t a[8], b[8], c[8];
    BUTTERFLY(sub[0 * stride1], sub[1 * stride1], a[0], a[1]);
    BUTTERFLY(sub[2 * stride1], sub[3 * stride1], a[2], a[3]);
    BUTTERFLY(sub[4 * stride1], sub[5 * stride1], a[4], a[5]);
    BUTTERFLY(sub[6 * stride1], sub[7 * stride1], a[6], a[7]);

    BUTTERFLY(a[0], a[2], b[0], b[2]);
    BUTTERFLY(a[1], a[3], b[1], b[3]);
    BUTTERFLY(a[4], a[6], b[4], b[6]);
    BUTTERFLY(a[5], a[7], b[5], b[7]);

    BUTTERFLY(b[0], b[4], c[0], c[4]);
    BUTTERFLY(b[1], b[5], c[1], c[5]);
    BUTTERFLY(b[2], b[6], c[2], c[6]);
    BUTTERFLY(b[3], b[7], c[3], c[7]);
    for (int k = 0; k < 8; k++) sub[k * stride1] = c[k];
  }
}

/* dist.rs:125-143 hadamard2d<LEN,W,H>: vertical pass then horizontal pass. */
static void hadamard2d(int32_t *data, int w, int h) {
  if (h == 4)
    hadamard4_1d(data, w, 1, h);
  else
    hadamard8_1d(data, w, 1, h);
  if (w == 4)
    hadamard4_1d(data, h, w, 1);
  else
    hadamard8_1d(data, h, w, 1);
}

#define PIXEL uint8_t
#define SFX(name) name##_u8
#include "dist_impl.h"
#undef PIXEL
#undef SFX

#define PIXEL uint16_t
#define SFX(name) name##_u16
#include "dist_impl.h"
#undef PIXEL
#undef SFX
