/* CPU oracle for the NNCodec uniform quantizer (TEST INFRASTRUCTURE, not product code).
 *
 * PARITY UNPINNED: the arithmetic of quantLayer/dequantLayer lives in the third-party pybind11
 * module `deepCABAC` (upstream fraunhoferhhi/nncodec, extensions/deepCABAC; NOT vendored in
 * /root/reference, no version pinned there, not installable here).  The reference has no tests
 * or golden vectors for it.  What IS pinned in-tree:
 *   - the step size  delta(qp, qp_density)            nnc_core/common.py:28-46   (exact)
 *   - reconstruction  value = level * delta           nnc_core/approximator/codebook.py:346-356
 *   - call shape, in/out dtypes, qp clip contract     nnc_core/approximator/baseline.py:39-62,98
 * The rounding rule below (nearest integer of |w|/delta, ties away from zero, computed in
 * float32 as (int)(|w|/delta + 0.5f)) restates ISO/IEC 15938-17 uniform reconstruction
 * quantization as published; it cannot be checked against the real module in this container.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include <math.h>
#include <stdint.h>

/* nnc_core/common.py:28-46.  Python semantics: `qp & (k-1)` on a two's-complement int and an
 * arithmetic right shift; C's >> on negative int32 is arithmetic with gcc. */
float nncq_stepsize(int qp, int qp_density)
{
    int k = 1 << qp_density;
    int mul = k + (qp & (k - 1));
    int shift = qp >> qp_density;
    return (float)ldexp((double)mul, shift - qp_density);
}

/* baseline.py:60-62 contract: the returned qp may be larger than requested "to avoid int32_t
 * overflow".  Restated as: smallest qp' >= qp whose largest level fits in int32. */
int nncq_clip_qp(float max_abs, int qp, int qp_density)
{
    if (!(max_abs <= 3.402823466e+38f)) return qp;   /* NaN / Inf maximum: qp as requested (our convention, no reference) */
    for (int it = 0; it < 1024; ++it) {
        float d = nncq_stepsize(qp, qp_density);
        float q = max_abs / d + 0.5f;
        if (q < 2147483648.0f) return qp;
        ++qp;
    }
    return qp;
}

/* baseline.py:48-57 (dq_flag = 0): w float32 -> int32 levels, row-major scan (scan_order 0).
 * Returns the qp actually used. */
int nncq_quant_urq(const float *w, int32_t *lvl, int64_t n, int qp, int qp_density)
{
    float max_abs = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        float a = fabsf(w[i]);
        if (a > max_abs || a != a) max_abs = a;      /* NaN sticks, like the bit-pattern maximum of the GPU kernel */
        if (max_abs != max_abs) break;
    }
    int qp_used = nncq_clip_qp(max_abs, qp, qp_density);
    float d = nncq_stepsize(qp_used, qp_density);
    for (int64_t i = 0; i < n; ++i) {
        float a = fabsf(w[i]);
        if (!(a <= 3.402823466e+38f)) { lvl[i] = 0; continue; }     /* NaN / Inf -> level 0 (our convention) */
        volatile float q = a / d;        /* keep the fp32 rounding of the quotient */
        volatile float r = q + 0.5f;
        int32_t m = r >= 2147483648.0f ? 2147483647 : (int32_t)r;
        lvl[i] = (w[i] < 0.0f) ? -m : m;
    }
    return qp_used;
}

/* baseline.py:98: int32 levels -> float32 values, value = level * delta. */
void nncq_dequant(const int32_t *lvl, float *w, int64_t n, int qp, int qp_density)
{
    float d = nncq_stepsize(qp, qp_density);
    for (int64_t i = 0; i < n; ++i) w[i] = (float)lvl[i] * d;
}
