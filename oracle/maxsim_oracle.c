/* maxsim_oracle.c -- plain-C restatement of the reference's masked MaxSim (TEST INFRASTRUCTURE).
 *
 * Follows evaluator/retrieval.py:166-213 of /root/reference one scalar at a time:
 *   out[q,c] = has[c] * sum_n qmask[q,n] * max_m ( pmask[c,m] ? <Q[q,n,:],P[c,m,:]> : -1e4 )
 * fp32 arithmetic; dot products accumulate in index order (the reference's SGEMM uses another
 * order, differences are ~1e-7 relative and covered by the test tolerance).
 * Built by oracle/build_oracle.py into oracle/_build/liboracle.so; only tests/, smoke() and
 * bench.py's cpu_baseline may load it.  Parity: pinned by tests/golden (outputs of the reference).
 */
#include <stdint.h>
#include <stddef.h>

#define NEG_FILL (-1e4f) /* evaluator/retrieval.py:185 */

/* returns 0; argmax (optional, [nq, np, lq] int32) gets the first index attaining the max */
int oracle_maxsim_masked(const float* Q, const float* P, const uint8_t* qmask, const uint8_t* pmask,
                         int nq, int np, int lq, int lp, int d, float* out, int32_t* argmax) {
  for (int c = 0; c < np; ++c) {
    int has = 0;
    for (int m = 0; m < lp; ++m) has |= pmask[(size_t)c * lp + m] != 0; /* :192 */
    for (int q = 0; q < nq; ++q) {
      float sum = 0.0f;
      for (int n = 0; n < lq; ++n) {
        const float* qv = Q + ((size_t)q * lq + n) * d;
        float best = 0.0f;
        int bi = 0;
        for (int m = 0; m < lp; ++m) {
          float s = NEG_FILL; /* :198 */
          if (pmask[(size_t)c * lp + m]) {
            const float* pv = P + ((size_t)c * lp + m) * d;
            s = 0.0f;
            for (int k = 0; k < d; ++k) s += qv[k] * pv[k]; /* :195 */
          }
          if (m == 0 || s > best) { best = s; bi = m; } /* :201, first index on ties */
        }
        if (argmax) argmax[((size_t)q * np + c) * lq + n] = bi;
        float v = best * (has ? 1.0f : 0.0f);                   /* :204 */
        v = v * (qmask[(size_t)q * lq + n] ? 1.0f : 0.0f);      /* :207 */
        sum += v;                                               /* :210 */
      }
      out[(size_t)q * np + c] = sum;
    }
  }
  return 0;
}
