/* Plain-C restatement of the counting / normalisation arithmetic of the reference's
 * BruteForce (cbn/parameter_learning/brute_force.py:17-53, :228-241) on integer codes.
 * TEST INFRASTRUCTURE ONLY: used by tests/ and bench.py's cpu_baseline leg as a fast checker
 * for sample counts too large for the numpy oracle.  Never linked into the product library.
 *
 * The reference finds the distinct rows [pa_1..pa_P, x] with a lexicographic sort
 * (torch.unique(dim=0), :42) and counts them; on codes k = position in the sorted domain the
 * same table is a dense histogram over the row-major index, node fastest.
 */
#include <stdint.h>
#include <string.h>

/* codes: uint8 [n_cols][ld]; family f = vars[f*max_vars .. +n_vars[f]) with cards alongside;
 * counts: int64, family f at offset[f].  Samples whose index is out of range are skipped. */
void oracle_count_families(const uint8_t* codes, int64_t ld, int64_t n, const int32_t* n_vars, const int32_t* vars,
                           const int32_t* cards, int32_t max_vars, const int64_t* offset, int32_t n_fams,
                           int64_t* counts) {
  for (int32_t f = 0; f < n_fams; ++f) {
    const int32_t nv = n_vars[f];
    const int32_t* v = vars + (int64_t)f * max_vars;
    const int32_t* c = cards + (int64_t)f * max_vars;
    int64_t cells = 1;
    for (int j = 0; j < nv; ++j) cells *= c[j];
    int64_t* t = counts + offset[f];
    for (int64_t s = 0; s < n; ++s) {
      int64_t idx = 0;
      int ok = 1;
      for (int j = 0; j < nv; ++j) {
        int code = codes[(int64_t)v[j] * ld + s];
        if (code >= c[j]) ok = 0;
        idx = idx * c[j] + code;
      }
      if (ok && idx < cells) t[idx] += 1;
    }
  }
}

/* joint = fp32(c)/fp32(n) (:43); cond = joint / (sum_x joint + 1e-10) (:228-241), sequential fp32 sum. */
void oracle_cpt_from_counts(const int64_t* counts, int64_t n_rows, int32_t card, int64_t n_total, float* joint,
                            float* cond) {
  const float nt = (float)n_total;
  for (int64_t r = 0; r < n_rows; ++r) {
    float parent = 0.0f;
    for (int x = 0; x < card; ++x) {
      float j = (float)counts[r * card + x] / nt;
      joint[r * card + x] = j;
      parent += j;
    }
    const float den = parent + 1e-10f;
    for (int x = 0; x < card; ++x) cond[r * card + x] = joint[r * card + x] / den;
  }
}
