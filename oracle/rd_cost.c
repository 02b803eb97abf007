/* rd_cost.c — CPU restatement of rav1e's RDO cost (TEST INFRASTRUCTURE ONLY, see oracle.h).
 *
 *   compute_rd_cost   src/rdo.rs:718-723
 *       rate_in_bits = (rate as f64) / ((1 << OD_BITRES) as f64)        OD_BITRES = 3
 *       fi.lambda.mul_add(rate_in_bits, distortion.0 as f64)
 *
 * The only floating-point value on the path.  `f64::mul_add` is the IEEE-754 fused multiply-add
 * (one rounding); `rate as f64` and the division by 8 are exact, `u64 as f64` rounds to nearest
 * even.  C's fma() has the same contract, so this restatement is exact, not within 1 ULP.
 * Pinned by tests/test_oracle_rd_cost.py against exact rational arithmetic.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "oracle.h"

double orc_compute_rd_cost(double lambda, uint32_t rate, uint64_t distortion) {
  const double rate_in_bits = (double)rate / 8.0; /* rdo.rs:721, OD_BITRES = 3 (ec.rs) */
  return fma(lambda, rate_in_bits, (double)distortion); /* rdo.rs:722 */
}

void orc_compute_rd_cost_batch(double lambda, const uint32_t *rate, const uint64_t *distortion,
                               size_t n, double *out) {
  for (size_t i = 0; i < n; i++) out[i] = orc_compute_rd_cost(lambda, rate[i], distortion[i]);
}
