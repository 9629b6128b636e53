// K3 — batched propagation: PB seed columns (one 32-byte sector per node: 8 x fp32 or 4 x fp64) advance
// together through the SAME fused iteration as spmv_fused.cu, each column with its own normaliser,
// error sum and stop decision.
//
// Replaces `NodeRanking.propagate` (/root/reference/pygrank/core/signals.py:225-226), which runs
// `rank` once per feature column: B sequential PageRank solves re-stream the CSR B times and pay one
// 32-byte sector for every 4-byte gather.  Here the index stream is read once per panel and every
// gathered sector carries PB useful values.
//
// Mapping.  A warp owns a merge tile of the item stream (same partition as the single-vector
// kernel) and walks it in passes of G*K items, G = 32/PB lane groups, K = 8 items per group.  Lane
// (g, c) accumulates column c over group g's K consecutive items: the PB lanes of a group read one
// node's sector together (coalesced), the 8 gathers of a lane are independent.  Row ends close
// segments; pieces open across groups are stitched with a log2(G)-step segmented scan over groups;
// finished rows are parked in shared memory and updated four (fp32) at a time so the row-aligned
// streams (z, q, z') are touched in full 128-byte lines.  Columns that have converged are frozen
// (z' = z) exactly where the reference's per-column loop would have stopped.
#include <stdlib.h>

#include "step_common.cuh"

namespace pgb {

constexpr int BBLOCK = 256;
constexpr int BWARPS = BBLOCK / 32;
constexpr int BK = 8;  // items per lane group per pass

struct BatchParams {
    int64_t n, nnz;
    const int32_t *indptr;
    const int32_t *tile_row;
    int32_t n_tiles, tile_items;
    const int32_t *istream;
    const void *zin;   // [n_cols_total][PB]
    void *zout;
    int64_t out_offset;
    const void *w, *sq, *c;  // per-row (NULL w/sq: degree-derived)
    const void *q;           // [n][PB]
    double alpha;
    double *sf;        // [PB][PGB_STATE_F64_LEN]
    int32_t *si;       // [PB][PGB_STATE_I32_LEN], then one shared ticket at si[PB*LEN]
    double *err_hist;  // [PB][hist_stride]
    int32_t hist_stride;
    double *span_acc;  // [n_tiles][PB]
    uint32_t *span_cnt;
};

template <typename T, int PB, bool SYMDEG>
__global__ void __launch_bounds__(BBLOCK, 3) panel_kernel(const BatchParams P) {
    constexpr int G = 32 / PB;            // lane groups per warp
    constexpr int PASS = G * BK;          // items per pass
    constexpr int NL = PASS / 32;         // item loads per lane per pass (1 for fp32, 2 for fp64)
    static_assert(PASS % 32 == 0, "a pass is a whole number of warp-wide loads");
    __shared__ T s_rows[BWARPS][PASS + 1][PB];     // finished-row sums of the pass
    __shared__ int32_t s_deg[BWARPS][PASS + 1];
    __shared__ double s_red[2][BWARPS][PB];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane % PB, g = lane / PB;
    const unsigned FULL = 0xffffffffu;
    const int32_t PADV = (int32_t)0x80000000;
    const T *__restrict__ zin = (const T *)P.zin;
    T *__restrict__ zout = (T *)P.zout;
    const T *__restrict__ qv = (const T *)P.q;
    const int32_t *__restrict__ istream = P.istream;

    // column state
    const bool active = P.si[c * PGB_STATE_I32_LEN + PGB_SI_STOP] == PGB_RUNNING;
    if (__ballot_sync(FULL, active) == 0u) return;  // every column has stopped: run-ahead launch is a no-op
    const T invS = (T)P.sf[c * PGB_STATE_F64_LEN + PGB_SF_INVS];
    const int err_mode = P.si[c * PGB_STATE_I32_LEN + PGB_SI_ERR_MODE];
    const T alpha = (T)P.alpha;
    double err = 0.0, tsum = 0.0;
    const int64_t total_items = P.n + P.nnz;
    const int TILE = P.tile_items;

    auto update_row = [&](int64_t row, T acc, int deg) {
        T wi, sqi;
        if (SYMDEG) {
            wi = deg > 0 ? (T)1 / (T)deg : (T)0;
            sqi = deg > 0 ? (T)sqrt((double)deg) : (T)1;
        } else {
            wi = ld_stream((const T *)P.w + row);
            sqi = ld_stream((const T *)P.sq + row);
        }
        const int64_t own = (P.out_offset + row) * PB + c;
        const T zi = __ldg(zin + own);
        T znew = zi;
        if (active) {
            znew = (alpha * wi * acc + ld_stream(qv + row * PB + c)) * invS;
            const double d = (double)sqi * fabs((double)znew - (double)zi);
            err += (err_mode == PGB_ERR_MSQ) ? d * d : d;
            tsum += (double)znew * (double)ld_stream((const T *)P.c + row);
        }
        zout[own] = znew;
    };

    // cross-tile completion of a row, one column per lane of group 0 (fence-free, see spmv_fused.cu)
    auto span_commit = [&](int32_t slot, T partial, uint32_t expected, int64_t row, int deg) {
        // called by the PB lanes of ONE group (mask below) with converged control flow
        const unsigned gmask = ((PB == 32) ? 0xffffffffu : ((1u << PB) - 1u)) << (g * PB);
        const double old = atomicAdd(&P.span_acc[(int64_t)slot * PB + c], (double)partial);
        unsigned dep = (__double_as_longlong(old) == 0x7ff8dead00000001ll) ? 1u : 0u;
#pragma unroll
        for (int o = PB / 2; o > 0; o >>= 1) dep |= __shfl_xor_sync(gmask, dep, o);  // all PB adds have returned
        uint32_t arrived = 0;
        if (c == 0) arrived = atomicAdd(&P.span_cnt[slot], 1u + dep);
        arrived = __shfl_sync(gmask, arrived, g * PB);
        if (arrived == expected - 1) {
            const unsigned long long raw = atomicExch((unsigned long long *)&P.span_acc[(int64_t)slot * PB + c], 0ull);
            unsigned done = 1u;
#pragma unroll
            for (int o = PB / 2; o > 0; o >>= 1) done &= __shfl_xor_sync(gmask, done, o);
            if (c == 0) atomicExch(&P.span_cnt[slot], 0u);
            update_row(row, (T)__longlong_as_double((long long)raw), deg);
        }
    };

    for (int32_t tile = blockIdx.x * BWARPS + warp; tile < P.n_tiles; tile += gridDim.x * BWARPS) {
        const int64_t item_lo = (int64_t)tile * TILE;
        const int64_t item_hi = (item_lo + TILE < total_items) ? item_lo + TILE : total_items;
        const int32_t r_lo = P.tile_row[tile], r_hi = P.tile_row[tile + 1];
        const bool lead_span = (r_hi > r_lo) && ((int64_t)P.indptr[r_lo] + r_lo < item_lo);
        int64_t r_cur = r_lo;
        T carry = (T)0;  // column c of the row still open before this pass (replicated over groups)

        for (int64_t I0 = item_lo; I0 < item_hi; I0 += PASS) {
            // ---- items of the pass: group g owns items [g*BK, (g+1)*BK) ----------------------------
            int32_t cols[BK];
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const int64_t gi = I0 + g * BK + k;
                cols[k] = (gi < item_hi) ? ld_stream(istream + gi) : PADV;   // PB lanes share the address
            }
            T x[BK];
#pragma unroll
            for (int k = 0; k < BK; ++k) x[k] = (cols[k] >= 0) ? __ldg(zin + (int64_t)cols[k] * PB + c) : (T)0;

            // marker bookkeeping: bit k of `mk` = item k of my group is a row end
            unsigned mk = 0u;
#pragma unroll
            for (int k = 0; k < BK; ++k) mk |= (cols[k] < 0 && cols[k] != PADV) ? (1u << k) : 0u;
            const int my_cnt = __popc(mk);
            // rows finished by earlier groups in this pass (group-uniform), via a scan over groups
            int before = 0, nrows = 0;
#pragma unroll
            for (int gg = 0; gg < G; ++gg) {
                const int cnt = __shfl_sync(FULL, my_cnt, gg * PB);
                if (gg < g) before += cnt;
                nrows += cnt;
            }
            // ---- merge my group's items -------------------------------------------------------------
            T run = (T)0, head = (T)0;
            int k_row = before, first_k = -1, first_deg = 0;
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                if ((mk >> k) & 1u) {
                    const int deg = -1 - cols[k];
                    if (first_k < 0) {
                        head = run;
                        first_k = k_row;
                        first_deg = deg;
                    } else {
                        s_rows[warp][k_row][c] = run;
                        if (c == 0) s_deg[warp][k_row] = deg;
                    }
                    run = (T)0;
                    ++k_row;
                } else {
                    run += x[k];
                }
            }
            // stitch across groups: segments restart at every group that closed a row
            const bool closed = first_k >= 0;
            T seg = run;
            bool f = closed;
#pragma unroll
            for (int d = PB; d < 32; d <<= 1) {
                const T t = __shfl_up_sync(FULL, seg, d);
                const bool tf = __shfl_up_sync(FULL, (int)f, d);
                if (lane >= d && !f) {
                    seg += t;
                    f = tf;
                }
            }
            T carry_in = __shfl_up_sync(FULL, seg, PB);
            if (g == 0) carry_in = (T)0;
            const unsigned closed_mask = __ballot_sync(FULL, closed);     // PB bits per group
            const unsigned below = (g == 0) ? 0u : (closed_mask & ((1u << (g * PB)) - 1u));
            if (below == 0u) carry_in += carry;
            if (closed) {
                s_rows[warp][first_k][c] = head + carry_in;
                if (c == 0) s_deg[warp][first_k] = first_deg;
            }
            const T last = __shfl_sync(FULL, seg, (G - 1) * PB + c);
            carry = (closed_mask == 0u) ? carry + last : last;
            __syncwarp();

            // ---- fused update of the finished rows, G rows per step ----------------------------------
            for (int k0 = 0; k0 < nrows; k0 += G) {
                const int k = k0 + g;
                if (k < nrows) {
                    const int64_t row = r_cur + k;
                    const T acc = s_rows[warp][k][c];
                    const int deg = s_deg[warp][k];
                    if (lead_span && row == r_lo) {
                        const int64_t t_a = ((int64_t)P.indptr[row] + row) / TILE;
                        span_commit(tile, acc, (uint32_t)(tile - t_a + 1), row, deg);
                    } else {
                        update_row(row, acc, deg);
                    }
                }
            }
            __syncwarp();
            r_cur += nrows;
        }

        // row still open at the end of the tile
        if (r_hi < P.n) {
            const int64_t b = P.indptr[r_hi], e = P.indptr[r_hi + 1];
            const int64_t e_lo = item_lo - r_lo, e_hi = item_hi - r_hi;
            const int64_t first_here = b > e_lo ? b : e_lo;
            if (e_hi > first_here && g == 0) {
                const int64_t t_a = (b + r_hi) / TILE, t_b = (e + r_hi) / TILE;
                span_commit((int32_t)t_b, carry, (uint32_t)(t_b - t_a + 1), r_hi, (int)(e - b));
            }
        }
        __syncwarp();
    }

    // ---- per-column grid reduction: lanes with equal c across groups, then warps, then one atomic/CTA
#pragma unroll
    for (int d = PB; d < 32; d <<= 1) {
        err += __shfl_xor_sync(FULL, err, d);
        tsum += __shfl_xor_sync(FULL, tsum, d);
    }
    if (g == 0) {
        s_red[0][warp][c] = err;
        s_red[1][warp][c] = tsum;
    }
    __syncthreads();
    if (threadIdx.x < PB) {
        double e = 0.0, t = 0.0;
        for (int wv = 0; wv < BWARPS; ++wv) {
            e += s_red[0][wv][threadIdx.x];
            t += s_red[1][wv][threadIdx.x];
        }
        atomicAdd(&P.sf[threadIdx.x * PGB_STATE_F64_LEN + PGB_SF_EACC], e);
        atomicAdd(&P.sf[threadIdx.x * PGB_STATE_F64_LEN + PGB_SF_TACC], t);
        __threadfence();
    }
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        const int ticket = atomicAdd(&P.si[PB * PGB_STATE_I32_LEN], 1);
        s_last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (threadIdx.x < PB)
            finalize_column(P.sf + threadIdx.x * PGB_STATE_F64_LEN, P.si + threadIdx.x * PGB_STATE_I32_LEN,
                            P.err_hist ? P.err_hist + (int64_t)threadIdx.x * P.hist_stride : nullptr);
        if (threadIdx.x == 0) P.si[PB * PGB_STATE_I32_LEN] = 0;
    }
}

template <typename T, int PB, bool SYMDEG>
static int launch_panel(const BatchParams &P, cudaStream_t st) {
    static PerDeviceInt ctas_on;
    int &ctas = ctas_on.here();
    if (ctas == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, panel_kernel<T, PB, SYMDEG>, BBLOCK, 0) != cudaSuccess ||
            v < 1)
            v = 2;
        ctas = v;
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, panel_kernel<T, PB, SYMDEG>) == cudaSuccess) {
            const size_t need = (fa.sharedSizeBytes + 1024) * (size_t)v;
            int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (pct > 100) pct = 100;
            cudaFuncSetAttribute(panel_kernel<T, PB, SYMDEG>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
    }
    int grid = sm_count() * ctas;
    const int need = (int)ceil_div(P.n_tiles, BWARPS);
    if (grid > need) grid = need;
    if (grid < 1) return 0;
    panel_kernel<T, PB, SYMDEG><<<grid, BBLOCK, 0, st>>>(P);
    PGB_LAUNCH_OK("panel_kernel");
    return 0;
}

}  // namespace pgb

using namespace pgb;

extern "C" {

int pgb_panel_width(int dtype) { return dtype == PGB_F32 ? 8 : (dtype == PGB_F64 ? 4 : 0); }

int pgb_affine_steps_batched(const pgb_csr *g, int dtype, double alpha, const void *w, const void *sq, const void *c,
                             const void *q, void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64,
                             int32_t *state_i32, double *err_hist, int32_t hist_stride, pgb_span_ws ws,
                             int first_step, int num_launches, void *stream) {
    if (!g) return fail("null graph");
    if (g->values) return fail("pgb_affine_steps_batched: weighted graphs are not supported by the panel kernel yet");
    if (!g->istream) return fail("pgb_affine_steps_batched needs pgb_csr.istream (pgb_build_item_stream)");
    if (g->n == 0) return 0;
    if ((w == nullptr) != (sq == nullptr)) return fail("pgb_affine_steps_batched: w and sq must both be given or both be NULL");
    if (first_step < 1) return fail("pgb_affine_steps_batched: first_step must be >= 1");
    if ((int64_t)g->n_tiles != ceil_div(g->n + g->nnz, g->tile_items)) return fail("graph n_tiles inconsistent");
    BatchParams P;
    memset(&P, 0, sizeof(P));
    P.n = g->n;
    P.nnz = g->nnz;
    P.indptr = g->indptr;
    P.tile_row = g->tile_row;
    P.n_tiles = g->n_tiles;
    P.tile_items = g->tile_items;
    P.istream = g->istream;
    P.out_offset = out_offset;
    P.w = w;
    P.sq = sq;
    P.c = c;
    P.q = q;
    P.alpha = alpha;
    P.sf = state_f64;
    P.si = state_i32;
    P.err_hist = err_hist;
    P.hist_stride = hist_stride;
    P.span_acc = ws.acc;
    P.span_cnt = ws.cnt;
    const bool symdeg = (w == nullptr);
    void *buf[2] = {zbuf0, zbuf1};
    for (int j = 0; j < num_launches; ++j) {
        const int k = first_step + j;
        P.zin = buf[(k - 1) & 1];
        P.zout = buf[k & 1];
        int rc;
        if (dtype == PGB_F32)
            rc = symdeg ? launch_panel<float, 8, true>(P, as_stream(stream)) : launch_panel<float, 8, false>(P, as_stream(stream));
        else if (dtype == PGB_F64)
            rc = symdeg ? launch_panel<double, 4, true>(P, as_stream(stream)) : launch_panel<double, 4, false>(P, as_stream(stream));
        else
            return fail("pgb_affine_steps_batched: unknown dtype %d", dtype);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
