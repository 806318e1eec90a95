/* sd_csr.c -- plain-C restatement of the reference's sparse hot loop.  TEST INFRASTRUCTURE, NOT
 * PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.
 *
 * PARITY UNPINNED: the reference (gcnmodel.py:39,130,153) calls theano.sparse.structured_dot,
 * whose CPU kernel (Theano 1.0.x `StructuredDotCSR`, "sd_csr"; Theano is an un-vendored
 * dependency, requirements.txt:5) is the loop below: for every row m, for every stored nonzero
 * of that row in CSR order, z[m, :] += val * b[col, :], accumulating in the output dtype
 * (float32), single-threaded.  tests/test_oracle.py checks this file against SciPy's
 * csr_matvecs (what Theano's Python `perform` path calls) bit for bit.
 *
 * Build: make -C oracle   ->  oracle/_build/libsdcsr.so
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* z[n_rows x k] = A[n_rows x n_cols] . b[n_cols x k]; ldb/ldz in floats */
void sd_csr_f32(int32_t n_rows, const int32_t* rowptr, const int32_t* colidx, const float* val,
                const float* b, int64_t ldb, float* z, int64_t ldz, int32_t k) {
  for (int32_t m = 0; m < n_rows; ++m) {
    float* zr = z + (size_t)m * ldz;
    memset(zr, 0, (size_t)k * sizeof(float));
    for (int32_t p = rowptr[m]; p < rowptr[m + 1]; ++p) {
      const float a = val[p];
      const float* br = b + (size_t)colidx[p] * ldb;
      for (int32_t n = 0; n < k; ++n) zr[n] += a * br[n];
    }
  }
}

/* The gradient wrt the dense operand, structured_dot(a.T, g) -> Theano's `sd_csc` on the CSR
 * arrays of a read as the CSC of a.T: scatter form, same accumulation dtype. */
void sd_csc_f32(int32_t n_rows_a, int32_t n_cols_a, const int32_t* rowptr, const int32_t* colidx,
                const float* val, const float* g, int64_t ldg, float* z, int64_t ldz, int32_t k) {
  for (int32_t c = 0; c < n_cols_a; ++c) memset(z + (size_t)c * ldz, 0, (size_t)k * sizeof(float));
  for (int32_t m = 0; m < n_rows_a; ++m) {
    const float* gr = g + (size_t)m * ldg;
    for (int32_t p = rowptr[m]; p < rowptr[m + 1]; ++p) {
      const float a = val[p];
      float* zr = z + (size_t)colidx[p] * ldz;
      for (int32_t n = 0; n < k; ++n) zr[n] += a * gr[n];
    }
  }
}
