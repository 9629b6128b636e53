// Shared by the per-iteration kernels (spmv_fused.cu: merge-path item stream; hsell.cu: hub-blocked
// sliced-ELL): the launch parameters, the device-side ConvergenceManager and the fused row update.
#pragma once
#include "common.cuh"

namespace pgb {

enum { MODE_CONV = 0, MODE_AFFINE = 1, MODE_POLY = 2 };

struct StepParams {
    int64_t n, nnz;
    const int32_t *indptr, *indices;
    const void *values;
    const int32_t *tile_row;
    int32_t n_tiles;
    const int32_t *istream;  // item-space index stream: row entries then the terminator -1-deg (v3)
    const void *vstream;     // item-space weights (weighted graphs), same positions as istream
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
    int step;                // index of this step in its run (in-kernel dropout draws a new mask per step)
    const pgb_hsell *hsell;  // host pointer: hub-blocked sliced-ELL form of the same graph (or NULL)
    void *partials;          // its per-filter workspace
    void *yacc;              // accumulate mode: y [n_slices + 1][32], zero between steps (or NULL: partial rows)
    // row-partitioned multi-GPU with the exchange fused into the step (pgb_affine_step_peer)
    int n_peers, peer_rank;
    void *peer_zout[PGB_MAX_PEERS];
    void *mc_zout;
    double *peer_acc[PGB_MAX_PEERS];
    const uint32_t *peer_mask;   // [local rows] which ranks read the row's value (NULL: all)
};

// store to a multicast (multimem) address: NVSwitch replicates it into every peer's buffer
__device__ __forceinline__ void multimem_store(float *mc, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_store(double *mc, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc), "d"(v) : "memory");
}

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
    }
    // sum(next ranks) is linear in the current ranks: alpha * sum_i ranks_i*rowsum_i(M) + sum(bias).  Kept current
    // on a stop too: a caller may clear STOP and go on (the plugin route does when the driver's rule differs).
    if (vsi[PGB_SI_QUOTIENT]) vsf[PGB_SF_INVS] = 1.0 / (vsf[PGB_SF_ALPHA] * tacc + vsf[PGB_SF_BIAS]);
    __threadfence();
}

// per-column ConvergenceManager (same rules as finalize_state; used by the panel kernels of spmm_fused.cu and hsell.cu)
__device__ __forceinline__ void finalize_column(double *sf, int32_t *si, double *err_hist) {
    volatile double *vsf = sf;
    volatile int32_t *vsi = si;
    const double tacc = vsf[PGB_SF_TACC], eacc = vsf[PGB_SF_EACC];
    vsf[PGB_SF_TACC] = 0.0;
    vsf[PGB_SF_EACC] = 0.0;
    if (vsi[PGB_SI_STOP] != PGB_RUNNING) return;
    const int k = vsi[PGB_SI_STEPS] + 1;
    vsi[PGB_SI_STEPS] = k;
    const int it = k + 1;
    const double errv = eacc / vsf[PGB_SF_MEAN];
    vsf[PGB_SF_LASTERR] = errv;
    if (err_hist) err_hist[k] = errv;
    int stop = PGB_RUNNING;
    if (it >= vsi[PGB_SI_MAX_ITERS])
        stop = PGB_MAX_ITERS;
    else if (vsi[PGB_SI_ERR_MODE] != PGB_ERR_ITERS && (it % vsi[PGB_SI_END_MODULO]) == 0 && errv <= vsf[PGB_SF_TOL])
        stop = PGB_CONVERGED;
    if (stop != PGB_RUNNING) {
        vsi[PGB_SI_ITERATION] = it;
        vsi[PGB_SI_STOP] = stop;
    } else if (vsi[PGB_SI_QUOTIENT]) {
        vsf[PGB_SF_INVS] = 1.0 / (vsf[PGB_SF_ALPHA] * tacc + vsf[PGB_SF_BIAS]);
    }
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
    T berr, btsum;      // fp32 batch sums (see apply)
    bool batch_sums;

    __device__ __forceinline__ void accumulate_error(double d) {
        if (err_mode == PGB_ERR_MAX)
            err = (err > d) ? err : d;
        else
            err += (err_mode == PGB_ERR_MSQ) ? d * d : d;
    }

    __device__ __forceinline__ void flush_batch() {
        if (err_mode == PGB_ERR_MAX)
            err = (err > (double)berr) ? err : (double)berr;
        else
            err += (double)berr;
        tsum += (double)btsum;
        berr = btsum = (T)0;
    }

    __device__ __forceinline__ RowUpdate(const StepParams &p) : P(p), err(0.0), tsum(0.0), berr((T)0), btsum((T)0),
                                                                batch_sums(false) {
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

    // Everything the update of one row reads besides the gathered sum: issued BEFORE the sum is
    // known so that these loads overlap the gather (hsell update pass).
    struct Loaded {
        T wi, sqi, zi, a, b;   // a: q (affine) / previous ranks (poly) / rscale (conv); b: c (affine) / xlap (conv)
    };

    __device__ __forceinline__ Loaded load(int64_t row, int deg) const {
        Loaded L;
        L.wi = L.sqi = L.zi = L.a = L.b = (T)0;
        if (MODE == MODE_CONV) {
            L.a = P.rscale ? ((const T *)P.rscale)[row] : (T)1;
            if (P.xlap) L.b = ((const T *)P.xlap)[row];
            return L;
        }
        if (SYMDEG) {
            L.wi = deg > 0 ? RowMath<T>::inv((T)deg) : (T)0;
            L.sqi = deg > 0 ? RowMath<T>::root((T)deg) : (T)1;
        } else {
            L.wi = ld_stream((const T *)P.w + row);
            L.sqi = ld_stream((const T *)P.sq + row);
        }
        L.zi = __ldg((const T *)P.zin + P.out_offset + row);
        if (MODE == MODE_AFFINE) {
            // row-aligned streams are touched once per launch: evict-first keeps L1/L2 for the gathers
            L.a = ld_stream((const T *)P.q + row);
            L.b = ld_stream((const T *)P.c + row);
        } else if (coef != (T)0) {
            L.a = ((T *)P.ranks)[row];
        }
        return L;
    }

    // load() may be called with deg = 0 before the row pointers have arrived (so that its loads are not
    // queued behind them); set_degree() then completes the degree-derived factors.
    __device__ __forceinline__ void set_degree(Loaded &L, int deg) const {
        if (SYMDEG && MODE != MODE_CONV) {
            L.wi = deg > 0 ? RowMath<T>::inv((T)deg) : (T)0;
            L.sqi = deg > 0 ? RowMath<T>::root((T)deg) : (T)1;
        }
    }

    // the new iterate: local buffer, or every rank's buffer when the exchange is fused into the step
    __device__ __forceinline__ void store_z(int64_t own, T v) const {
        if (P.n_peers == 0) {
            ((T *)P.zout)[own] = v;
        } else if (P.mc_zout) {
            multimem_store((T *)P.mc_zout + own, v);
        } else {
            const uint32_t m = P.peer_mask ? P.peer_mask[own - P.out_offset] : 0xffffffffu;
            for (int r = 0; r < P.n_peers; ++r)
                if ((m >> r) & 1u) ((T *)P.peer_zout[r])[own] = v;
        }
    }

    __device__ __forceinline__ void apply(int64_t row, T acc, const Loaded &L) {
        const int64_t own = P.out_offset + row;
        if (MODE == MODE_CONV) {
            T y = P.rscale ? L.a * acc : acc;
            if (P.xlap) y = L.b - y;
            ((T *)P.zout)[P.out_perm ? (int64_t)P.out_perm[row] : own] = y;
            return;
        }
        if (MODE == MODE_AFFINE) {
            const T znew = (alpha * L.wi * acc + L.a) * invS;
            store_z(own, znew);
            if (sizeof(T) == 4 && batch_sums) {
                // fp32 mode inside the hsell update pass: a few rows are summed in fp32 and flushed to the
                // fp64 accumulators once per group (flush_batch) — one DADD pair per group instead of per row
                const T d = L.sqi * fabsf((float)(znew - L.zi));
                if (err_mode == PGB_ERR_MAX)
                    berr = (berr > d) ? berr : d;
                else
                    berr += (err_mode == PGB_ERR_MSQ) ? d * d : d;
                btsum += znew * L.b;
            } else {
                double d = (double)L.sqi * fabs((double)znew - (double)L.zi);
                accumulate_error(d);
                tsum += (double)znew * (double)L.b;
            }
        } else {  // MODE_POLY
            const T pw = L.sqi * L.zi;
            if (coef != (T)0) {  // abstract_filters.py:226-228
                const T prev = L.a;
                const T cur = prev + pw * coef;
                ((T *)P.ranks)[row] = cur;
                // fp64: the literal |prev - cur| the reference's Mabs sees; fp32: the exact increment
                double d = (sizeof(T) == 8) ? fabs((double)prev - (double)cur) : fabs((double)coef * (double)pw);
                accumulate_error(d);
            }
            store_z(own, L.wi * acc);
        }
    }

    __device__ __forceinline__ void operator()(int64_t row, T acc, int deg) {
        const Loaded L = load(row, deg);
        apply(row, acc, L);
    }
};

// Grid-level end of a step: one fp64 atomic per CTA for the error numerator and the next normaliser;
// the last CTA plays ConvergenceManager (or just resets the ticket in deferred multi-GPU mode).
template <typename U>
__device__ __forceinline__ void step_epilogue(const StepParams &P, U &update, double *s_red) {
    const bool is_max = update.err_mode == PGB_ERR_MAX;
    const double err = is_max ? block_max(update.err, s_red) : block_sum(update.err, s_red);
    const double tsum = block_sum(update.tsum, s_red);
    if (threadIdx.x == 0) {
        if (is_max)   // non-negative doubles order like their bit patterns
            atomicMax((unsigned long long *)&P.sf[PGB_SF_EACC], (unsigned long long)__double_as_longlong(err));
        else
            atomicAdd(&P.sf[PGB_SF_EACC], err);
        atomicAdd(&P.sf[PGB_SF_TACC], tsum);
        __threadfence();
        const int ticket = atomicAdd(&P.si[PGB_SI_TICKET], 1);
        if (ticket == (int)gridDim.x - 1) {
            __threadfence();
            if (P.finalize) {
                finalize_state(P.sf, P.si, P.err_hist);
            } else {
                P.si[PGB_SI_TICKET] = 0;
                if (P.n_peers > 0) {
                    // hand this rank's sums to every rank (slot = our rank); pgb_state_finalize_peer adds the slots
                    volatile double *vsf = P.sf;
                    const double t = vsf[PGB_SF_TACC], e = vsf[PGB_SF_EACC];
                    vsf[PGB_SF_TACC] = 0.0;
                    vsf[PGB_SF_EACC] = 0.0;
                    for (int r = 0; r < P.n_peers; ++r) {
                        P.peer_acc[r][2 * P.peer_rank] = t;
                        P.peer_acc[r][2 * P.peer_rank + 1] = e;
                    }
                    __threadfence_system();
                }
            }
        }
    }
    if (P.n_peers > 0) __threadfence_system();   // peer stores of this thread are performed before the kernel ends
}

// hsell.cu: one fused step (gather kernel + update kernel) on the hub-blocked sliced-ELL form
template <int MODE>
int hsell_step(const StepParams &P, const pgb_hsell *h, void *partials, int dtype, bool symdeg, cudaStream_t st);

}  // namespace pgb
