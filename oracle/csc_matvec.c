/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under pygrank_b200/ may link,
 * import or execute this file; it exists so tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time the
 * CUDA path against a CPU restatement of the reference's arithmetic.
 *
 * What it restates.  pygrank's default backend computes `conv(signal, M)` as
 * `signal @ M` (/root/reference/pygrank/core/backend/numpy.py:64-65).  With M a
 * scipy CSR matrix that expression is evaluated by scipy (third-party, unpinned
 * in /root/reference/setup.py:28-30; 1.18.1 in this image) as `(M.T @ x)`, and
 * the transpose of a CSR matrix is the same three arrays read as CSC, so the
 * arithmetic that finally runs is scipy.sparse._sparsetools.csc_matvec: a
 * serial scatter, columns of M.T (= rows of M) visited in ascending order,
 * entries of each column in storage order,   y[Ai[k]] += Ax[k] * x[j].
 * The functions below restate that published algorithm in plain C so the
 * summation order (and therefore every rounding) is the reference's.
 *
 * Pinned by tests/test_oracle.py: bit-equal to scipy's own `x @ M` on the
 * golden graphs, and the filters built on it are bit-equal to the reference's
 * outputs stored under tests/golden/.
 */
#include <stdint.h>
#include <stddef.h>

/* y += x @ M  for M = CSR(indptr[n_row+1], indices[nnz], data[nnz]); y has n_col entries. */
void oracle_rowvec_times_csr_f64(int64_t n_row,
                                 const int32_t *indptr,
                                 const int32_t *indices,
                                 const double *data,
                                 const double *x,
                                 double *y)
{
    for (int64_t j = 0; j < n_row; ++j) {
        const double xj = x[j];
        const int32_t end = indptr[j + 1];
        for (int32_t k = indptr[j]; k < end; ++k)
            y[indices[k]] += data[k] * xj;
    }
}

/* Same scatter with 64-bit index arrays (scipy switches when nnz >= 2^31). */
void oracle_rowvec_times_csr_f64_i64(int64_t n_row,
                                     const int64_t *indptr,
                                     const int64_t *indices,
                                     const double *data,
                                     const double *x,
                                     double *y)
{
    for (int64_t j = 0; j < n_row; ++j) {
        const double xj = x[j];
        const int64_t end = indptr[j + 1];
        for (int64_t k = indptr[j]; k < end; ++k)
            y[indices[k]] += data[k] * xj;
    }
}

/* fp32 flavour used only to time an fp32 CPU leg; parity is always taken against f64. */
void oracle_rowvec_times_csr_f32(int64_t n_row,
                                 const int32_t *indptr,
                                 const int32_t *indices,
                                 const float *data,
                                 const float *x,
                                 float *y)
{
    for (int64_t j = 0; j < n_row; ++j) {
        const float xj = x[j];
        const int32_t end = indptr[j + 1];
        for (int32_t k = indptr[j]; k < end; ++k)
            y[indices[k]] += data[k] * xj;
    }
}

/* Row sums in storage order (scipy's csr @ ones): used to cross-check degrees(). */
void oracle_csr_row_sums_f64(int64_t n_row,
                             const int32_t *indptr,
                             const double *data,
                             double *out)
{
    for (int64_t i = 0; i < n_row; ++i) {
        double s = 0.0;
        for (int32_t k = indptr[i]; k < indptr[i + 1]; ++k)
            s += data[k];
        out[i] = s;
    }
}
