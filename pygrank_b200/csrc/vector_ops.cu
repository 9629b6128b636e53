// K6 — the O(n) passes either side of the fused steps: start of GraphFilter.rank
// (/root/reference/pygrank/algorithms/filters/abstract_filters.py:52-56), conversion between the
// user's node order / value domain and the engine's (relabelled, pre-scaled) one, and the three
// reductions the backend surface needs (sum, sum|.|, dot — core/backend/numpy.py:2).
#include "common.cuh"

namespace pgb {

template <typename T>
__global__ void scale_kernel(int64_t n, const T *__restrict__ a, const T *__restrict__ b, double scale,
                             const int32_t *__restrict__ perm, T *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t src = perm ? (int64_t)perm[i] : i;
        T v = a[src];
        if (b) v *= b[i];
        out[i] = (scale == 1.0) ? v : (T)(v * (T)scale);
    }
}

template <typename T>
__global__ void unscale_kernel(int64_t n, const T *__restrict__ a, const T *__restrict__ b,
                               const double *__restrict__ dev_scale, double scale,
                               const int32_t *__restrict__ perm, T *__restrict__ out) {
    const double s = (dev_scale ? *dev_scale : 1.0) * scale;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T v = a[i];
        if (b) v *= b[i];
        if (s != 1.0) v = (T)(v * (T)s);
        out[perm ? (int64_t)perm[i] : i] = v;
    }
}

template <typename T>
__global__ void reduce3_kernel(int64_t n, const T *__restrict__ x, const T *__restrict__ y, double *sums) {
    __shared__ double scratch[32];
    double sa = 0.0, ss = 0.0, sd = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        sa += fabs(v);
        ss += v;
        if (y) sd += v * (double)y[i];
    }
    sa = block_sum(sa, scratch);
    ss = block_sum(ss, scratch);
    sd = block_sum(sd, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[0], sa);
        atomicAdd(&sums[1], ss);
        if (y) atomicAdd(&sums[2], sd);
    }
}

struct PeerOut {
    int n;
    void *z[PGB_MAX_PEERS];
    const uint32_t *mask;
};

template <typename T>
__global__ void affine_init_kernel(int64_t n, const T *__restrict__ p, const T *__restrict__ warm,
                                   const T *__restrict__ sq, const T *__restrict__ c, double coef,
                                   const T *__restrict__ coefvec, const int32_t *__restrict__ perm,
                                   int64_t out_offset, T *__restrict__ z0, T *__restrict__ q, double *sf,
                                   const PeerOut peers) {
    __shared__ double scratch[32];
    const double norm = sf[PGB_SF_NORM];
    double tacc = 0.0, bacc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t src = perm ? (int64_t)perm[i] : i;
        const T pn = (T)((double)p[src] / norm);                 // abstract_filters.py:55
        const T start = warm ? warm[src] : pn;                   // abstract_filters.py:56
        const T sqi = sq[i];
        const T zi = start / sqi;
        const T qi = (T)((coefvec ? (double)coefvec[i] : coef) * (double)pn) / sqi;
        if (peers.n == 0) {
            z0[out_offset + i] = zi;
        } else {   // row-partitioned multi-GPU: the start vector goes straight into every rank's buffer
            const uint32_t m = peers.mask ? peers.mask[i] : 0xffffffffu;
            for (int r = 0; r < peers.n; ++r)
                if ((m >> r) & 1u) ((T *)peers.z[r])[out_offset + i] = zi;
        }
        if (q) q[i] = qi;
        if (c) tacc += (double)zi * (double)c[i];
        bacc += (double)qi * (double)sqi;
    }
    tacc = block_sum(tacc, scratch);
    bacc = block_sum(bacc, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(&sf[PGB_SF_TACC], tacc);
        atomicAdd(&sf[PGB_SF_BIAS], bacc);
    }
    if (peers.n > 0) __threadfence_system();
}

__global__ void affine_init_finish_kernel(double *sf, int32_t *si) {
    const double t = sf[PGB_SF_TACC];
    sf[PGB_SF_TACC] = 0.0;
    sf[PGB_SF_EACC] = 0.0;
    sf[PGB_SF_INVS] = si[PGB_SI_QUOTIENT] ? 1.0 / (sf[PGB_SF_ALPHA] * t + sf[PGB_SF_BIAS]) : 1.0;
}

// ---- staging of a panel job: features [n][B] in user order <-> column-major blocks in the engine's row order ----------
// A slot of the panel loads / stores ONE column at a time; from a row-major [n][B] matrix in user order that is one
// 32-byte sector per 4-byte value at random rows.  The job therefore runs on a staged copy, stage[j][i] =
// cols[perm[i]][j0 + j], written by one tiled transpose (rows gathered through perm in full lines), and its results
// go back the same way.
template <typename T>
__global__ void panel_stage_in_kernel(int64_t n, const T *__restrict__ cols, int64_t row_stride, int64_t col_stride,
                                      const int32_t *__restrict__ perm, int64_t j0, int G, T *__restrict__ stage) {
    __shared__ T tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int jj0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t i = i0 + r;
        const int j = jj0 + threadIdx.x;
        T v = (T)0;
        if (i < n && j < G) {
            const int64_t src = perm ? (int64_t)perm[i] : i;
            v = cols[src * row_stride + (j0 + j) * col_stride];
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = jj0 + r;
        const int64_t i = i0 + threadIdx.x;
        if (j < G && i < n) stage[(int64_t)j * n + i] = tile[threadIdx.x][r];
    }
}

template <typename T>
__global__ void panel_stage_out_kernel(int64_t n, const T *__restrict__ stage, const int32_t *__restrict__ perm,
                                       int64_t j0, int G, T *__restrict__ out, int64_t out_row_stride) {
    __shared__ T tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int jj0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = jj0 + r;
        const int64_t i = i0 + threadIdx.x;
        tile[r][threadIdx.x] = (j < G && i < n) ? stage[(int64_t)j * n + i] : (T)0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t i = i0 + r;
        const int j = jj0 + threadIdx.x;
        if (i < n && j < G) {
            const int64_t dst = perm ? (int64_t)perm[i] : i;
            out[dst * out_row_stride + j0 + j] = tile[threadIdx.x][r];
        }
    }
}

}  // namespace pgb

using namespace pgb;

extern "C" {

int pgb_scale(int64_t n, int dtype, const void *a, const void *b, double scale, const int32_t *perm, void *out,
              void *stream) {
    if (n <= 0) return 0;
    const int grid = stride_grid(n, 256);
    if (dtype == PGB_F32)
        scale_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, (const float *)a, (const float *)b, scale, perm,
                                                                 (float *)out);
    else if (dtype == PGB_F64)
        scale_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n, (const double *)a, (const double *)b, scale, perm,
                                                                  (double *)out);
    else
        return fail("pgb_scale: unknown dtype %d", dtype);
    PGB_LAUNCH_OK("scale_kernel");
    return 0;
}

int pgb_unscale(int64_t n, int dtype, const void *a, const void *b, const double *dev_scale, double scale,
                const int32_t *perm, void *out, void *stream) {
    if (n <= 0) return 0;
    const int grid = stride_grid(n, 256);
    if (dtype == PGB_F32)
        unscale_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, (const float *)a, (const float *)b, dev_scale,
                                                                   scale, perm, (float *)out);
    else if (dtype == PGB_F64)
        unscale_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n, (const double *)a, (const double *)b, dev_scale,
                                                                    scale, perm, (double *)out);
    else
        return fail("pgb_unscale: unknown dtype %d", dtype);
    PGB_LAUNCH_OK("unscale_kernel");
    return 0;
}

int pgb_reduce3(int64_t n, int dtype, const void *x, const void *y, double *sums, void *stream) {
    if (n <= 0) return 0;
    const int grid = stride_grid(n, 256) < 592 ? stride_grid(n, 256) : 592;
    if (dtype == PGB_F32)
        reduce3_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, (const float *)x, (const float *)y, sums);
    else if (dtype == PGB_F64)
        reduce3_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n, (const double *)x, (const double *)y, sums);
    else
        return fail("pgb_reduce3: unknown dtype %d", dtype);
    PGB_LAUNCH_OK("reduce3_kernel");
    return 0;
}

static int affine_init_impl(int64_t n, int dtype, const void *p, const void *warm, const void *sq, const void *c,
                            double coef, const void *coefvec, const int32_t *perm, int64_t out_offset, void *z0,
                            void *q, double *state_f64, const pgb_peers *peers, void *stream) {
    if (n <= 0) return 0;
    if (!sq) return fail("pgb_affine_init: the sq vector is required");
    PeerOut po;
    memset(&po, 0, sizeof(po));
    if (peers) {
        if (peers->n < 1 || peers->n > PGB_MAX_PEERS) return fail("pgb_affine_init_peer: bad peer description");
        po.n = peers->n;
        for (int r = 0; r < peers->n; ++r) po.z[r] = peers->zbuf0[r];
        po.mask = peers->row_mask;
    }
    const int grid = stride_grid(n, 256) < 592 ? stride_grid(n, 256) : 592;
    if (dtype == PGB_F32)
        affine_init_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
            n, (const float *)p, (const float *)warm, (const float *)sq, (const float *)c, coef,
            (const float *)coefvec, perm, out_offset, (float *)z0, (float *)q, state_f64, po);
    else if (dtype == PGB_F64)
        affine_init_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
            n, (const double *)p, (const double *)warm, (const double *)sq, (const double *)c, coef,
            (const double *)coefvec, perm, out_offset, (double *)z0, (double *)q, state_f64, po);
    else
        return fail("pgb_affine_init: unknown dtype %d", dtype);
    PGB_LAUNCH_OK("affine_init_kernel");
    return 0;
}

int pgb_affine_init(int64_t n, int dtype, const void *p, const void *warm, const void *sq, const void *c,
                    double coef, const void *coefvec, const int32_t *perm, int64_t out_offset, void *z0, void *q,
                    double *state_f64, void *stream) {
    return affine_init_impl(n, dtype, p, warm, sq, c, coef, coefvec, perm, out_offset, z0, q, state_f64, nullptr,
                            stream);
}

int pgb_affine_init_peer(int64_t n, int dtype, const void *p, const void *warm, const void *sq, const void *c,
                         double coef, const void *coefvec, const int32_t *perm, int64_t out_offset, void *q,
                         double *state_f64, const pgb_peers *peers, void *stream) {
    if (!peers) return fail("pgb_affine_init_peer: no peers");
    return affine_init_impl(n, dtype, p, warm, sq, c, coef, coefvec, perm, out_offset, nullptr, q, state_f64, peers,
                            stream);
}

int pgb_affine_init_finish(double *state_f64, int32_t *state_i32, void *stream) {
    affine_init_finish_kernel<<<1, 1, 0, as_stream(stream)>>>(state_f64, state_i32);
    PGB_LAUNCH_OK("affine_init_finish_kernel");
    return 0;
}

int pgb_panel_stage(int64_t n, int dtype, int direction, void *matrix, int64_t row_stride, int64_t col_stride,
                    const int32_t *perm, int64_t j0, int32_t n_cols, void *stage, void *stream) {
    if (n <= 0 || n_cols <= 0) return 0;
    if (!matrix || !stage) return fail("pgb_panel_stage: null matrix / stage");
    if (n > 32ll * 2147483647ll || (n_cols + 31) / 32 > 65535) return fail("pgb_panel_stage: too many rows / columns");
    const dim3 grid((unsigned)ceil_div(n, 32), (unsigned)((n_cols + 31) / 32)), block(32, 8);
    cudaStream_t st = as_stream(stream);
    if (direction == 0) {   // matrix (user order, any strides) -> stage [n_cols][n] (engine order)
        if (dtype == PGB_F32)
            panel_stage_in_kernel<float><<<grid, block, 0, st>>>(n, (const float *)matrix, row_stride, col_stride, perm, j0,
                                                                 n_cols, (float *)stage);
        else if (dtype == PGB_F64)
            panel_stage_in_kernel<double><<<grid, block, 0, st>>>(n, (const double *)matrix, row_stride, col_stride, perm,
                                                                  j0, n_cols, (double *)stage);
        else
            return fail("pgb_panel_stage: unknown dtype %d", dtype);
    } else {                // stage -> matrix columns j0.. (row-major: col_stride must be 1)
        if (col_stride != 1) return fail("pgb_panel_stage: the output matrix must have unit column stride");
        if (dtype == PGB_F32)
            panel_stage_out_kernel<float><<<grid, block, 0, st>>>(n, (const float *)stage, perm, j0, n_cols, (float *)matrix,
                                                                  row_stride);
        else if (dtype == PGB_F64)
            panel_stage_out_kernel<double><<<grid, block, 0, st>>>(n, (const double *)stage, perm, j0, n_cols,
                                                                   (double *)matrix, row_stride);
        else
            return fail("pgb_panel_stage: unknown dtype %d", dtype);
    }
    PGB_LAUNCH_OK("panel_stage_kernel");
    return 0;
}

}  // extern "C"
