// K2 / K4 / K5 — the per-iteration kernel of the hot path: one merge-path CSR row-gather
// fused with the filter's elementwise update and the convergence reduction.
//
// Reference op sequence replaced per iteration (paths under /root/reference/pygrank):
//   conv(ranks, M)                      core/backend/numpy.py:64-65 (scipy csc_matvec)
//   * alpha + personalization*(1-alpha) algorithms/filters/adhoc.py:36       (3 vector passes)
//   safe_div(ranks, sum(ranks))         algorithms/filters/abstract_filters.py:133-134 (2 passes)
//   Mabs(prev)(cur) <= tol              algorithms/convergence.py:96-101, measures/supervised.py:101-106
//
// Work decomposition.  The CSR is cut by merge path: the sequence "entries of row 0, end-marker of
// row 0, entries of row 1, end-marker of row 1, ..." (n + nnz items) is split into tiles of
// TILE_ITEMS consecutive items, so every tile costs the same no matter how skewed the degrees
// are (power-law hubs, 38 % empty rows on RMAT).  A persistent grid (CTAs-per-SM x 148 SMs) walks
// the tiles.  Per tile: (A) the column indices are streamed with coalesced evict-first loads and
// the gather vector z is read through the read-only path into shared memory; (B) every thread
// consumes IPT consecutive merge items from shared memory (IPT odd -> conflict-free strides) and
// deposits per-row sums; (C) one thread per finished row applies the fused update with coalesced
// reads/writes of the row-aligned vectors.  Rows cut by a tile boundary are completed by the LAST
// tile that reaches them (fp64 atomic partial + ticket per completing tile; no spinning, no
// ordering assumption).  Grid-level sums (error numerator, next normaliser) are one fp64 atomic
// per CTA; the last CTA to finish plays ConvergenceManager on the device.
#include "common.cuh"

namespace pgb {

constexpr int BLOCK = 256;
constexpr int IPT = 9;  // odd: thread-blocked reads of shared memory hit distinct banks
constexpr int TILE_ITEMS = BLOCK * IPT;

enum { MODE_CONV = 0, MODE_AFFINE = 1, MODE_POLY = 2 };

struct StepParams {
    int64_t n, nnz;
    const int32_t *indptr, *indices;
    const void *values;
    const int32_t *tile_row;
    int32_t n_tiles;
    const void *zin;
    void *zout;
    int64_t out_offset;  // index of local row 0 inside the (full-length) z vectors
    const void *w, *sq, *c, *q;
    void *ranks;
    const double *coef;
    const void *rscale, *xlap;
    const int32_t *out_perm;
    double alpha;
    double *sf;
    int32_t *si;
    double *err_hist;
    double *span_acc;
    uint32_t *span_cnt;
    int finalize;
};

__device__ __forceinline__ void finalize_state(double *sf, int32_t *si, double *err_hist) {
    volatile double *vsf = sf;
    volatile int32_t *vsi = si;
    const double tacc = vsf[PGB_SF_TACC], eacc = vsf[PGB_SF_EACC];
    vsf[PGB_SF_TACC] = 0.0;
    vsf[PGB_SF_EACC] = 0.0;
    vsi[PGB_SI_TICKET] = 0;
    const int k = vsi[PGB_SI_STEPS] + 1;  // _step calls done
    vsi[PGB_SI_STEPS] = k;
    const int it = k + 1;                 // ConvergenceManager.iteration at the next has_converged()
    const double errv = eacc / vsf[PGB_SF_MEAN];
    vsf[PGB_SF_LASTERR] = errv;
    if (err_hist) err_hist[k] = errv;
    int stop = PGB_RUNNING;
    if (it >= vsi[PGB_SI_MAX_ITERS])                                    // convergence.py:86-90
        stop = PGB_MAX_ITERS;
    else if (vsi[PGB_SI_ERR_MODE] != PGB_ERR_ITERS && (it % vsi[PGB_SI_END_MODULO]) == 0 &&
             errv <= vsf[PGB_SF_TOL])                                   // convergence.py:97-101
        stop = PGB_CONVERGED;
    if (stop != PGB_RUNNING) {
        vsi[PGB_SI_ITERATION] = it;
        vsi[PGB_SI_STOP] = stop;
    } else if (vsi[PGB_SI_QUOTIENT]) {
        // sum(next ranks) is linear in the current ranks: alpha * sum_i ranks_i*rowsum_i(M) + sum(bias)
        vsf[PGB_SF_INVS] = 1.0 / (vsf[PGB_SF_ALPHA] * tacc + vsf[PGB_SF_BIAS]);
    }
    __threadfence();
}

__global__ void state_finalize_kernel(double *sf, int32_t *si, double *err_hist) {
    if (si[PGB_SI_STOP] != PGB_RUNNING) return;
    finalize_state(sf, si, err_hist);
}

template <typename T>
struct RowMath;
template <>
struct RowMath<float> {
    static __device__ __forceinline__ float inv(float d) { return 1.0f / d; }
    static __device__ __forceinline__ float root(float d) { return sqrtf(d); }
};
template <>
struct RowMath<double> {
    static __device__ __forceinline__ double inv(double d) { return 1.0 / d; }
    static __device__ __forceinline__ double root(double d) { return sqrt(d); }
};

template <typename T, int MODE, bool SYMDEG>
struct RowUpdate {
    const StepParams &P;
    T alpha, invS, coef;
    int err_mode;
    double err, tsum;

    __device__ __forceinline__ RowUpdate(const StepParams &p) : P(p), err(0.0), tsum(0.0) {
        alpha = (T)p.alpha;
        invS = (T)1;
        coef = (T)0;
        err_mode = PGB_ERR_MABS;
        if (MODE != MODE_CONV) {
            err_mode = p.si[PGB_SI_ERR_MODE];
            invS = (T)p.sf[PGB_SF_INVS];
            if (MODE == MODE_POLY) coef = (T)p.coef[p.si[PGB_SI_STEPS] + 1];
        }
    }

    __device__ __forceinline__ void operator()(int64_t row, T acc, int deg) {
        const int64_t own = P.out_offset + row;
        if (MODE == MODE_CONV) {
            T y = P.rscale ? ((const T *)P.rscale)[row] * acc : acc;
            if (P.xlap) y = ((const T *)P.xlap)[row] - y;
            ((T *)P.zout)[P.out_perm ? (int64_t)P.out_perm[row] : own] = y;
            return;
        }
        T wi, sqi;
        if (SYMDEG) {
            wi = deg > 0 ? RowMath<T>::inv((T)deg) : (T)0;
            sqi = deg > 0 ? RowMath<T>::root((T)deg) : (T)1;
        } else {
            wi = ((const T *)P.w)[row];
            sqi = ((const T *)P.sq)[row];
        }
        const T zi = __ldg((const T *)P.zin + own);
        if (MODE == MODE_AFFINE) {
            const T znew = (alpha * wi * acc + ((const T *)P.q)[row]) * invS;
            ((T *)P.zout)[own] = znew;
            double d = (double)sqi * fabs((double)znew - (double)zi);
            err += (err_mode == PGB_ERR_MSQ) ? d * d : d;
            tsum += (double)znew * (double)((const T *)P.c)[row];
        } else {  // MODE_POLY
            const T pw = sqi * zi;
            if (coef != (T)0) {  // abstract_filters.py:226-228
                const T prev = ((T *)P.ranks)[row];
                const T cur = prev + pw * coef;
                ((T *)P.ranks)[row] = cur;
                // fp64: the literal |prev - cur| the reference's Mabs sees; fp32: the exact increment
                double d = (sizeof(T) == 8) ? fabs((double)prev - (double)cur) : fabs((double)coef * (double)pw);
                err += (err_mode == PGB_ERR_MSQ) ? d * d : d;
            }
            ((T *)P.zout)[own] = wi * acc;
        }
    }
};

template <typename T, bool WEIGHTED, int MODE, bool SYMDEG>
__global__ void __launch_bounds__(BLOCK, 4) tile_kernel(const StepParams P) {
    __shared__ T s_val[TILE_ITEMS];
    __shared__ T s_rowsum[TILE_ITEMS + 1];
    __shared__ int32_t s_end[TILE_ITEMS];
    __shared__ double s_red[32];

    if (MODE != MODE_CONV) {
        if (P.si[PGB_SI_STOP] != PGB_RUNNING) return;  // run-ahead launches after convergence are no-ops
    }
    const int tid = threadIdx.x;
    const T *__restrict__ zin = (const T *)P.zin;
    const int32_t *__restrict__ indices = P.indices;
    const int32_t *__restrict__ indptr = P.indptr;
    const T *__restrict__ values = (const T *)P.values;
    RowUpdate<T, MODE, SYMDEG> update(P);
    const int64_t total_items = P.n + P.nnz;

    for (int32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const int64_t item_lo = (int64_t)tile * TILE_ITEMS;
        const int64_t item_hi = (item_lo + TILE_ITEMS < total_items) ? item_lo + TILE_ITEMS : total_items;
        const int32_t r_lo = P.tile_row[tile], r_hi = P.tile_row[tile + 1];
        const int64_t e_lo = item_lo - r_lo, e_hi = item_hi - r_hi;
        const int nrows = r_hi - r_lo;           // rows whose end marker lies in this tile
        const int nedges = (int)(e_hi - e_lo);   // entries in this tile
        const int64_t row0_begin = indptr[r_lo];
        const int start0 = (int)((row0_begin > e_lo ? row0_begin : e_lo) - e_lo);

        // ---- phase A: stream indices, gather z, stage in shared memory ----------------------
        for (int k = tid; k < nrows; k += BLOCK) s_end[k] = (int32_t)((int64_t)indptr[r_lo + 1 + k] - e_lo);
        for (int k = tid; k <= nrows; k += BLOCK) s_rowsum[k] = (T)0;
        {
            int32_t cols[IPT];
            T vals[IPT];
#pragma unroll
            for (int s = 0; s < IPT; ++s) {
                const int i = s * BLOCK + tid;
                cols[s] = (i < nedges) ? ld_stream(indices + e_lo + i) : -1;
            }
#pragma unroll
            for (int s = 0; s < IPT; ++s) vals[s] = (cols[s] >= 0) ? __ldg(zin + cols[s]) : (T)0;
            if (WEIGHTED) {
#pragma unroll
                for (int s = 0; s < IPT; ++s) {
                    const int i = s * BLOCK + tid;
                    if (i < nedges) vals[s] *= ld_stream(values + e_lo + i);
                }
            }
#pragma unroll
            for (int s = 0; s < IPT; ++s) {
                const int i = s * BLOCK + tid;
                if (i < nedges) s_val[i] = vals[s];
            }
        }
        __syncthreads();

        // ---- phase B: each thread consumes IPT consecutive merge items -----------------------
        const int nitems = nrows + nedges;
        const int d = tid * IPT;
        if (d < nitems) {
            int lo = 0, hi = nrows;  // rows whose marker precedes item d
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_end[mid] + mid < d)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            int k = lo, ec = d - lo;
            T run = (T)0;
            bool first = true;
            const int stop = (d + IPT < nitems) ? IPT : nitems - d;
            for (int s = 0; s < stop; ++s) {
                if (k < nrows && ec == s_end[k]) {
                    if (first)
                        atomicAdd(&s_rowsum[k], run);  // row may have started in an earlier thread
                    else
                        s_rowsum[k] = run;             // row lies entirely inside this thread
                    first = false;
                    run = (T)0;
                    ++k;
                } else {
                    run += s_val[ec];
                    ++ec;
                }
            }
            atomicAdd(&s_rowsum[k], run);  // open row continues in the next thread / tile
        }
        __syncthreads();

        // ---- phase C: fused row update ---------------------------------------------------------
        const bool lead_span = (nrows > 0) && (row0_begin < e_lo);  // first finished row began earlier
        for (int k = tid; k < nrows; k += BLOCK) {
            if (k == 0 && lead_span) continue;
            const int deg = s_end[k] - (k ? s_end[k - 1] : start0);
            update((int64_t)r_lo + k, s_rowsum[k], deg);
        }
        // rows cut by a tile boundary: partial -> slot of the completing tile; last arrival finishes the row
        const bool has_trail = (r_hi < P.n) && (nrows > 0 ? (int)(s_end[nrows - 1]) < nedges : nedges > 0);
        if ((tid == 0 && lead_span) || (tid == 32 && has_trail)) {
            const int64_t r = (tid == 0) ? r_lo : r_hi;
            const double partial = (double)((tid == 0) ? s_rowsum[0] : s_rowsum[nrows]);
            const int64_t b = indptr[r], e = indptr[r + 1];
            const int64_t t_a = (b + r) / TILE_ITEMS, t_b = (e + r) / TILE_ITEMS;
            const uint32_t expected = (uint32_t)(t_b - t_a + 1);
            atomicAdd(&P.span_acc[t_b], partial);
            __threadfence();
            const uint32_t arrived = atomicAdd(&P.span_cnt[t_b], 1u);
            if (arrived == expected - 1) {
                __threadfence();
                const unsigned long long bits = atomicExch((unsigned long long *)&P.span_acc[t_b], 0ull);
                P.span_cnt[t_b] = 0;
                update(r, (T)__longlong_as_double((long long)bits), (int)(e - b));
            }
        }
        __syncthreads();  // shared memory is reused by the next tile
    }

    if (MODE != MODE_CONV) {
        const double err = block_sum(update.err, s_red);
        const double tsum = block_sum(update.tsum, s_red);
        if (tid == 0) {
            atomicAdd(&P.sf[PGB_SF_EACC], err);
            atomicAdd(&P.sf[PGB_SF_TACC], tsum);
            __threadfence();
            const int ticket = atomicAdd(&P.si[PGB_SI_TICKET], 1);
            if (ticket == (int)gridDim.x - 1) {
                __threadfence();
                if (P.finalize)
                    finalize_state(P.sf, P.si, P.err_hist);
                else
                    P.si[PGB_SI_TICKET] = 0;
            }
        }
    }
}

template <typename T, bool WEIGHTED, int MODE, bool SYMDEG>
static int launch_tiles(const StepParams &P, cudaStream_t st) {
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, tile_kernel<T, WEIGHTED, MODE, SYMDEG>, BLOCK, 0) !=
                cudaSuccess || v < 1)
            v = 2;
        ctas_per_sm = v;
    }
    int grid = sm_count() * ctas_per_sm;
    if (grid > P.n_tiles) grid = P.n_tiles;
    if (grid < 1) return 0;
    tile_kernel<T, WEIGHTED, MODE, SYMDEG><<<grid, BLOCK, 0, st>>>(P);
    PGB_LAUNCH_OK("tile_kernel");
    return 0;
}

template <int MODE>
static int dispatch(const StepParams &P, int dtype, bool symdeg, cudaStream_t st) {
    const bool weighted = P.values != nullptr;
    if (dtype == PGB_F32) {
        if (weighted) return launch_tiles<float, true, MODE, false>(P, st);
        if (symdeg) return launch_tiles<float, false, MODE, true>(P, st);
        return launch_tiles<float, false, MODE, false>(P, st);
    } else if (dtype == PGB_F64) {
        if (weighted) return launch_tiles<double, true, MODE, false>(P, st);
        if (symdeg) return launch_tiles<double, false, MODE, true>(P, st);
        return launch_tiles<double, false, MODE, false>(P, st);
    }
    return fail("unknown dtype %d", dtype);
}

static int fill_graph(StepParams &P, const pgb_csr *g) {
    if (!g) return fail("null graph");
    if (g->tile_items != TILE_ITEMS)
        return fail("graph partitioned with tile_items=%d, library uses %d", g->tile_items, TILE_ITEMS);
    if (g->nnz >= (1ll << 31) || g->n >= (1ll << 31)) return fail("graph exceeds the int32 limits of one device");
    if ((int64_t)g->n_tiles != ceil_div(g->n + g->nnz, TILE_ITEMS)) return fail("graph n_tiles inconsistent");
    P.n = g->n;
    P.nnz = g->nnz;
    P.indptr = g->indptr;
    P.indices = g->indices;
    P.values = g->values;
    P.tile_row = g->tile_row;
    P.n_tiles = g->n_tiles;
    return 0;
}

}  // namespace pgb

using namespace pgb;

extern "C" {

int pgb_tile_items(void) { return TILE_ITEMS; }

int pgb_spmv(const pgb_csr *g, int dtype, const void *z, const void *rscale, const void *x_for_laplacian,
             const int32_t *out_perm, void *out, pgb_span_ws ws, void *stream) {
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (fill_graph(P, g)) return 1;
    if (P.n == 0) return 0;
    P.zin = z;
    P.zout = out;
    P.rscale = rscale;
    P.xlap = x_for_laplacian;
    P.out_perm = out_perm;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    return dispatch<MODE_CONV>(P, dtype, false, as_stream(stream));
}

int pgb_affine_steps(const pgb_csr *g, int dtype, double alpha, const void *w, const void *sq, const void *c,
                     const void *q, void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64,
                     int32_t *state_i32, double *err_hist, pgb_span_ws ws, int first_step, int num_launches,
                     int finalize, void *stream) {
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (fill_graph(P, g)) return 1;
    if (P.n == 0) return 0;
    const bool symdeg = (w == nullptr && sq == nullptr);
    if (!symdeg && (!w || !sq)) return fail("pgb_affine_steps: w and sq must both be given or both be NULL");
    if (symdeg && g->values) return fail("pgb_affine_steps: degree-derived scales need an unweighted graph");
    if (first_step < 1) return fail("pgb_affine_steps: first_step must be >= 1");
    P.alpha = alpha;
    P.w = w;
    P.sq = sq;
    P.c = c;
    P.q = q;
    P.out_offset = out_offset;
    P.sf = state_f64;
    P.si = state_i32;
    P.err_hist = err_hist;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    P.finalize = finalize;
    void *buf[2] = {zbuf0, zbuf1};
    for (int j = 0; j < num_launches; ++j) {
        const int k = first_step + j;
        P.zin = buf[(k - 1) & 1];
        P.zout = buf[k & 1];
        if (dispatch<MODE_AFFINE>(P, dtype, symdeg, as_stream(stream))) return 1;
    }
    return 0;
}

int pgb_poly_steps(const pgb_csr *g, int dtype, const void *w, const void *sq, const double *coef, void *ranks,
                   void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64, int32_t *state_i32,
                   double *err_hist, pgb_span_ws ws, int first_step, int num_launches, int finalize, void *stream) {
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (fill_graph(P, g)) return 1;
    if (P.n == 0) return 0;
    const bool symdeg = (w == nullptr && sq == nullptr);
    if (!symdeg && (!w || !sq)) return fail("pgb_poly_steps: w and sq must both be given or both be NULL");
    if (symdeg && g->values) return fail("pgb_poly_steps: degree-derived scales need an unweighted graph");
    if (first_step < 1) return fail("pgb_poly_steps: first_step must be >= 1");
    P.w = w;
    P.sq = sq;
    P.coef = coef;
    P.ranks = ranks;
    P.out_offset = out_offset;
    P.sf = state_f64;
    P.si = state_i32;
    P.err_hist = err_hist;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    P.finalize = finalize;
    void *buf[2] = {zbuf0, zbuf1};
    for (int j = 0; j < num_launches; ++j) {
        const int k = first_step + j;
        P.zin = buf[(k - 1) & 1];
        P.zout = buf[k & 1];
        if (dispatch<MODE_POLY>(P, dtype, symdeg, as_stream(stream))) return 1;
    }
    return 0;
}

int pgb_state_finalize(double *state_f64, int32_t *state_i32, double *err_hist, void *stream) {
    state_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(state_f64, state_i32, err_hist);
    PGB_LAUNCH_OK("state_finalize_kernel");
    return 0;
}

}  // extern "C"
