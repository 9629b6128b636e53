/*
 * pgb200.h — C-ABI of the B200 graph-filter propagation engine (libpgb200.so).
 *
 * This is the drop-in boundary for pygrank's iterative node-ranking hot path.  The
 * reference is pure Python; its plugin surface for this path is the backend module
 * contract /root/reference/pygrank/core/backend/specification.py:5-117 (conv, degrees,
 * scipy_sparse_to_backend, graph_dropout, sum, abs, ...) bound by
 * /root/reference/pygrank/core/backend/__init__.py:40-84.  A backend module
 * (pygrank_b200/backend.py) implements that contract by calling the entry points below
 * through ctypes; INTEGRATION.md shows the binding a pygrank maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; memory is owned
 *     by the caller (PyTorch tensors); nothing is allocated or freed behind the ABI;
 *   - every entry point returns 0 on success, non-zero on failure, and leaves a message
 *     retrievable with pgb_last_error() (the Python shim raises Exception(msg), the
 *     error style of the reference, e.g. core/backend/__init__.py:42,83);
 *   - `stream` is a cudaStream_t passed as an opaque pointer (0 = default stream); all
 *     work is enqueued on it and no entry point synchronises unless stated;
 *   - dtype: PGB_F32 or PGB_F64 is the arithmetic/storage type of node vectors; all
 *     grid-level reductions accumulate in fp64 in both modes;
 *   - index types: int32 column indices and int32 row pointers (nnz < 2^31 per device;
 *     larger graphs are row-partitioned across devices before they reach this ABI).
 */
#ifndef PGB200_H
#define PGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGB_ABI_VERSION 4

enum { PGB_F32 = 0, PGB_F64 = 1 };

/* pgb_csr_build flags */
enum {
    PGB_BUILD_SYMMETRIZE = 1,      /* also insert (col,row) for every (row,col)            */
    PGB_BUILD_DROP_SELF_LOOPS = 2, /* discard row == col                                    */
    PGB_BUILD_BINARY = 4           /* duplicates collapse to weight 1 (else weights summed) */
};

/* scale kinds for pgb_make_scales: S[S != 0] = 1/S (preprocessing.py:111) after an
 * optional sqrt (preprocessing.py:115,132) */
enum { PGB_SCALE_ONE = 0, PGB_SCALE_RECIP = 1, PGB_SCALE_RSQRT = 2 };

/* error measures of ConvergenceManager (convergence.py:96-101; measures/supervised.py) */
enum {
    PGB_ERR_MABS = 0, /* sum|prev-cur| / n   (Mabs, supervised.py:101-106; default) */
    PGB_ERR_L1 = 1,   /* sum|prev-cur|       (L1,   supervised.py:125-130)          */
    PGB_ERR_MSQ = 2,  /* sum (prev-cur)^2/n  (MSQ,  supervised.py:117-122)          */
    PGB_ERR_ITERS = 3, /* error_type == "iters": never converges early (convergence.py:97-98) */
    PGB_ERR_MAX = 4    /* max|prev-cur|       (MaxDifference, supervised.py:93-98)           */
};

/* stop reasons written to pgb_state_i32[PGB_SI_STOP] */
enum { PGB_RUNNING = 0, PGB_CONVERGED = 1, PGB_MAX_ITERS = 2 };

/* Pull-CSR of the propagation operator: row i lists the sources j with a_ji != 0, so that
 * conv(x, M)_i = (x @ M)_i = R_i * sum_j a_ji * (L_j * x_j)   (numpy.py:64-65 with the
 * factorised normalisation of preprocessing.py:109-138).  `values` holds the RAW weights
 * a_ji in the vector dtype, or NULL for an unweighted graph (all a_ji == 1).
 * `tile_row` is the merge-path partition written by pgb_mergepath_partition. */
typedef struct pgb_csr {
    int64_t n;               /* nodes                                     */
    int64_t nnz;             /* stored entries                            */
    const int32_t *indptr;   /* [n+1]                                     */
    const int32_t *indices;  /* [nnz] ascending within a row              */
    const void *values;      /* [nnz] raw weights (dtype) or NULL         */
    const int32_t *tile_row; /* [n_tiles+1]                               */
    int32_t n_tiles;
    int32_t tile_items;      /* must equal pgb_tile_items()               */
    const int32_t *istream;  /* [n+nnz] item stream: each row's indices then -1-deg (pgb_build_item_stream) */
    const void *vstream;     /* [n+nnz] weights at the same positions (dtype) or NULL                      */
    const struct pgb_hsell *hsell; /* hub-blocked sliced-ELL form (HOST pointer to the struct) or NULL     */
} pgb_csr;

/* Hub-blocked sliced-ELL form of a pull CSR (unweighted, or weighted: hub_vals / tail_vals) whose nodes are ranked by degree (built once
 * per graph and vector dtype with pgb_hsell_count + pgb_hsell_fill).  Rows are cut into slices of 32
 * (one lane per row); columns into n_blocks "hub blocks" of block_cols columns — what one SM keeps of
 * the gather vector in shared memory — and a tail.  The entries of (slice, block) form a UNIT of
 * `rounds` rounds, stored [round][lane]: a hub round is one 32-bit word per lane holding two 16-bit
 * block-local columns (block_cols = padding), a tail round one 32-bit global column per lane (-1 =
 * padding).  The units of a block (slices ascending), then of the next block, form one stream of rounds
 * (each block padded to a whole chunk); the tail units form a second stream.  Streams are cut into
 * CHUNKS of PGB_HSELL_CHUNK rounds, the unit of work of one warp: chunk c starts at word c*32*CHUNK and
 * carries the rounds at which a unit ends inside it (endmask) and the number of its first PIECE
 * (p_first; pieces are numbered in stream order, hub stream first).  A piece ends at every unit end and
 * at the chunk end and yields one row of 32 partial sums, written to partial row piece_row[piece].
 * Partial rows are slice-major: the update pass reads upd_rows[s] = (first row, count) of slice s as one
 * contiguous run.  A slice with more than heavy_parts pieces (hub rows: their units span many chunks) is
 * first reduced by groups of 32 consecutive rows into second-level rows (reduce_items; stored after the
 * first-level rows) and upd_rows names those; if it still has more than heavy_parts it is also listed in
 * heavy_slices (one CTA adds them).  The padding pieces at the end of a block's stream write the last
 * (dump) row.  Chunks are dealt to
 * n_ctas CTAs in contiguous ranges (cta_*_begin).  With n_segments > 1 (row-partitioned multi-GPU) the
 * gather vector is n_segments ranges of seg_len entries (one per rank, each degree-ranked), hub block b
 * is the union of entries [b*block_cols/n_segments, ...) of every range, and the CSR given to the
 * builders uses the virtual column b*block_cols + segment*(block_cols/n_segments) + offset-in-block. */
#define PGB_HSELL_CHUNK 32
typedef struct pgb_hsell {
    int64_t n_rows;          /* local rows                                    */
    int64_t n_slices;        /* ceil(n_rows / 32)                             */
    int64_t n_partials;      /* partial rows (32 values each)                 */
    int64_t seg_len;         /* entries per segment of the gather vector      */
    int32_t n_segments;
    int32_t block_cols;      /* H (<= 65535)                                  */
    int32_t n_blocks;        /* K                                             */
    int32_t n_ctas;          /* CTAs the schedule was cut for                 */
    int32_t n_hub_chunks, n_tail_chunks;
    int32_t n_heavy, heavy_parts;     /* slices listing more than heavy_parts partial rows          */
    int32_t n_reduce, reserved0;      /* second-level reduction items (groups of <= 32 partial rows) */
    const uint32_t *hub_chunks;       /* [n_hub_chunks][2]: p_first, endmask                        */
    const uint32_t *tail_chunks;      /* [n_tail_chunks][2]: p_first, endmask                       */
    const uint32_t *hub_words;        /* [n_hub_chunks*CHUNK*32]                                    */
    const int32_t *tail_cols;         /* [n_tail_chunks*CHUNK*32]                                   */
    const int32_t *piece_row;         /* [pieces] partial row written by each piece                 */
    const int32_t *upd_rows;          /* [n_slices][2]: first partial row, count read by the update pass */
    const int32_t *heavy_slices;      /* [n_heavy]                                                  */
    const int32_t *reduce_items;      /* [n_reduce][3]: first row, count (<= 32), output row        */
    const int32_t *block_chunk_begin; /* [n_blocks+1] first hub chunk of each block                 */
    const int32_t *cta_hub_begin;     /* [n_ctas+1]                                                 */
    const int32_t *cta_tail_begin;    /* [n_ctas+1]                                                 */
    const int32_t *piece_slice;       /* [pieces] slice of each piece (n_slices for padding pieces): accumulate mode */
    const void *hub_vals;             /* weighted graphs: [n_hub_chunks*CHUNK*32][2] edge values of the two halves of
                                         every hub word, in the dtype of the form (0 in padding slots); NULL: all 1 */
    const void *tail_vals;            /* weighted graphs: [n_tail_chunks*CHUNK*32] edge values of the tail entries   */
} pgb_hsell;

/* Cross-tile workspace of one running filter: rows that straddle merge-path tiles are
 * completed by the last tile to arrive.  Zero-filled by the caller once; self-resetting. */
typedef struct pgb_span_ws {
    double *acc;        /* [n_tiles * ncols]  */
    uint32_t *cnt;      /* [n_tiles]          */
    void *partials;     /* [hsell.n_partials][32] of the vector dtype when the graph has an hsell form, else NULL */
    void *yacc;         /* accumulate mode of the hsell step: [hsell.n_slices + 1][32] of the vector dtype, zeroed by
                         * the caller once (the update pass re-zeroes what it reads).  Non-NULL selects the mode: every
                         * piece is added to its slice's row with one coalesced RED.ADD instead of being stored as a
                         * partial row, and the update pass streams y — faster (no partial-row round trip through HBM,
                         * no reduce kernel) but the order of the floating-point additions is not fixed; NULL keeps
                         * the deterministic partial-row path. */
} pgb_span_ws;

/* Device-resident iteration state (mirror of ConvergenceManager, convergence.py:24-101). */
enum { /* indices into state_f64[PGB_STATE_F64_LEN] */
    PGB_SF_ALPHA = 0, /* multiplier of the propagated term in the next normaliser        */
    PGB_SF_BIAS = 1,  /* sum of the affine term (e.g. (1-alpha)*sum(p))                  */
    PGB_SF_INVS = 2,  /* 1/sum(ranks) the CURRENT launch divides by (use_quotient)       */
    PGB_SF_TACC = 3,  /* accumulator: sum_i ranks_i * c_i of the current launch          */
    PGB_SF_EACC = 4,  /* accumulator: error numerator of the current launch              */
    PGB_SF_TOL = 5,   /* max(tol, epsilon) or 0 for tol=None (convergence.py:101)        */
    PGB_SF_MEAN = 6,  /* divisor of the error (n for Mabs/MSQ, 1 for L1)                 */
    PGB_SF_LASTERR = 7,
    PGB_SF_NORM = 8,  /* sum|personalization| (abstract_filters.py:52)                   */
    PGB_SF_PSUM = 9,  /* sum(personalization)                                            */
    PGB_SF_AMUL = 10, /* panel path: this column's multiplier of the gathered sum        */
    PGB_STATE_F64_LEN = 16
};
enum { /* indices into state_i32[PGB_STATE_I32_LEN] */
    PGB_SI_TICKET = 0,
    PGB_SI_STEPS = 1,     /* _step calls completed                                        */
    PGB_SI_STOP = 2,      /* PGB_RUNNING / PGB_CONVERGED / PGB_MAX_ITERS                  */
    PGB_SI_ITERATION = 3, /* ConvergenceManager.iteration when the loop stopped           */
    PGB_SI_MAX_ITERS = 4,
    PGB_SI_END_MODULO = 5,
    PGB_SI_ERR_MODE = 6,
    PGB_SI_QUOTIENT = 7,  /* use_quotient (abstract_filters.py:133-134)                   */
    PGB_STATE_I32_LEN = 16
};

int pgb_abi_version(void);
const char *pgb_last_error(void);
int pgb_tile_items(void);       /* merge-path items (rows + entries) per tile            */
int pgb_device_sm_count(int device);
/* Item-space stream of a CSR (what the fused kernels actually read): for every row its column
 * indices followed by one terminator -1-deg; vstream (optional) carries the weights. */
int pgb_build_item_stream(int64_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices, int dtype,
                          const void *values, int32_t *istream, void *vstream, void *stream);
/* Roofline probe: streams the graph's index array and gathers z[indices[k]] exactly like the fused
 * kernel's phase 2, and nothing else — the measured ceiling of any row-gather formulation for this
 * graph (bench.py reports the fused kernel against it).  scratch: >= 8 bytes. */
int pgb_gather_probe(const pgb_csr *g, int dtype, const void *z, void *scratch, void *stream);
/* 4 (default): hsell when the graph carries that form, else the item-stream kernel; 3: always the
 * item-stream kernel (A/B timing and parity between the two per-iteration kernels). */
int pgb_set_kernel_variant(int variant);

/* ---- synthetic graphs (bench/tests; no reference counterpart, graphs are downloaded in
 *      /root/reference/pygrank/benchmarks/download.py:62-72) ------------------------------ */
int pgb_rmat_edges(int scale, int64_t first_edge, int64_t num_edges, uint64_t seed,
                   uint32_t t1, uint32_t t2, uint32_t t3, int32_t *src, int32_t *dst, void *stream);
int pgb_ba_edges(int64_t n, int m, int64_t first_slot, int64_t num_slots, uint64_t seed,
                 int32_t *src, int32_t *dst, void *stream);

/* ---- K1: graph -> CSR (replaces nx.to_scipy_sparse_array / coo.tocsr,
 *      /root/reference/pygrank/core/utils/preprocessing.py:103 and fastgraph/fastgraph.py:73-78) */
size_t pgb_csr_build_workspace_bytes(int64_t n, int64_t nnz_in, int flags, int weighted);
/* COO (row, col[, val]) -> canonical CSR (ascending columns, duplicates merged).
 * out_indices/out_values need capacity nnz_in * (SYMMETRIZE ? 2 : 1).  *out_nnz_host is
 * written after an internal stream synchronise. */
int pgb_csr_build(int64_t n, int64_t nnz_in, const int32_t *row, const int32_t *col, const double *val,
                  int flags, void *workspace, size_t workspace_bytes, int32_t *out_indptr,
                  int32_t *out_indices, double *out_values, int64_t *out_nnz_host, void *stream);
/* rows[k] = i for indptr[i] <= k < indptr[i+1] (CSR -> COO row expansion) */
int pgb_csr_expand_rows(int64_t n, int64_t nnz, const int32_t *indptr, int32_t *rows, void *stream);
/* perm = node ids sorted by descending degree (stable); iperm[perm[i]] = i */
size_t pgb_degree_order_workspace_bytes(int64_t n);
int pgb_degree_order(int64_t n, const int32_t *indptr, void *workspace, size_t workspace_bytes,
                     int32_t *perm, int32_t *iperm, void *stream);
int pgb_relabel_coo(int64_t nnz, const int32_t *iperm, int32_t *row, int32_t *col, void *stream);
/* Hub-signature order (refines pgb_degree_order; the internal node order is the engine's own choice — the
 * reference keeps the user's order, preprocessing.py:151).  Degree ranks are cut into regions of G nodes up to
 * rank `span` (the hub blocks of pgb_hsell) plus one tail region; INSIDE every region nodes are re-sorted by the
 * set of hub regions their row touches (presence bits, region 0 most significant, rows touching it first; ties
 * keep the degree order).  Column block membership is unchanged, but the 32 rows of a slice now share their
 * blocks, so (slice, block) units that were too sparse for the shared-memory path become dense: on RMAT graphs
 * the low-degree rows' entries move from the L2-gather tail into hub units.  In: perm/iperm from
 * pgb_degree_order and the CSR in ORIGINAL labels; out: perm/iperm of the refined order. */
#define PGB_SIGNATURE_WORDS 8
size_t pgb_hub_order_workspace_bytes(int64_t n, int32_t words);
int pgb_hub_order(int64_t n, const int32_t *indptr, const int32_t *indices, int32_t G, int64_t span,
                  void *workspace, size_t workspace_bytes, int32_t *perm, int32_t *iperm, void *stream);
int pgb_mergepath_partition(int64_t n, int64_t nnz, const int32_t *indptr, int32_t n_tiles,
                            int32_t *tile_row, void *stream);

/* ---- hub-blocked sliced-ELL builders (no reference counterpart: the reference hands scipy's CSR to
 *      csc_matvec as is, core/backend/numpy.py:64-65) ------------------------------------------------ */
/* Largest block_cols the gather kernel can keep in shared memory for this dtype (0: unknown dtype). */
int pgb_hsell_max_block_cols(int dtype);
/* Pass 1 — one warp per slice: hub_rounds[b*n_slices+s] = rounds (2 entries each) of the unit of
 * slice s in block b, or 0 when the slice has fewer than min_entries + round_cost * rounds entries there
 * (a unit costs about round_cost L1 wavefronts per round plus a partial row, a tail entry one wavefront:
 * sparse units stay in the tail); tail_rounds[w*n_slices+s] = longest tail row of the slice inside tail window w.
 * TAIL WINDOWS (n_windows <= 16): a slice's tail is cut by where the gathered entry lies in the gather vector,
 * window w = positions [w*window_len, (w+1)*window_len); the tail stream is window-major, so while the grid sweeps
 * it only one window of z is live in L2 — what keeps the L2 gathers of a row-partitioned graph (gather vector several
 * times the L2) from degenerating into 32-byte HBM sectors.  A slice whose longest tail row has fewer than
 * window_min_rounds entries stays one unit, in the last window (pgb_hsell_fill recognises it by the zero rounds of
 * its other windows).  n_windows = 1: one unit per slice. */
int pgb_hsell_count(int64_t n, const int32_t *indptr, const int32_t *indices, int32_t block_cols, int32_t n_blocks,
                    int32_t min_entries, double round_cost, int32_t n_segments, int64_t seg_len, int32_t n_windows,
                    int64_t window_len, int32_t window_min_rounds, int32_t *hub_rounds, int32_t *tail_rounds,
                    void *stream);
/* Pass 2 — writes the round data and piece_row.  The caller supplies, per (block, slice) in
 * block-major order and per (window, slice) in window-major order for the tail: the first round of the unit in its
 * stream and the number of its first piece (exclusive scans), and slice_ptr (first partial row of every slice: the
 * pieces of a slice take consecutive rows, blocks ascending, then its tail windows ascending).  hub_words must be
 * pre-filled with the padding word (block_cols | block_cols << 16), tail_cols with -1 and piece_row
 * with the dump row.  scratch (int32[nnz], or NULL) enables the bank-aware slot order:
 * within a unit, entry position p of lane l gets a column of shared-memory bank (l + p) mod banks
 * when the row has one (banks = 32 for fp32, 16 for fp64 vectors), which removes most bank conflicts
 * of the gather kernel; the order of a row's entries has no other meaning.
 * WEIGHTED graphs: `values` ([nnz], `dtype`, the CSR's edge values) are written next to the indices — hub_vals
 * [words][2] (the values of the two halves of every hub word), tail_vals [tail words] — both zero-filled by the caller
 * (padding slots multiply the zero padding entry of the gather vector).  values = NULL: unweighted form. */
int pgb_hsell_fill(int64_t n, const int32_t *indptr, const int32_t *indices, int32_t block_cols, int32_t n_blocks,
                   int32_t n_segments, int64_t seg_len, const int32_t *hub_rounds, const int32_t *tail_rounds,
                   const int64_t *hub_round_base, const int64_t *hub_part_base, const int64_t *tail_round_base,
                   const int64_t *tail_part_base, const int32_t *slice_ptr, uint32_t *hub_words, int32_t *tail_cols,
                   int32_t *piece_row, int32_t *scratch, int32_t banks, int32_t n_windows, int64_t window_len,
                   int dtype, const void *values, void *hub_vals, void *tail_vals, void *stream);
/* K8 — graph_dropout inside the gather kernel of the hsell form (torch backends' semantics,
 * /root/reference/pygrank/core/backend/pytorch.py:34-38; called once per iteration, abstract_filters.py:59-62): while
 * p > 0 every stored entry is dropped with probability p by a counter-based hash of (seed, gather launches since this
 * call, slot) and the survivors are rescaled by 1/(1-p); a new mask every launch, nothing streamed, no values array.  Process-global like
 * the backend selection of the reference; p = 0 switches it off.  The linear next-normaliser of the quotient
 * (PGB_SI_QUOTIENT) assumes the unmasked operator: callers run dropout with the quotient off. */
int pgb_hsell_set_dropout(double p, uint64_t seed);
/* Experiment knob: warps (of 32) per CTA that prefer tail units (L2 gathers) over hub units. */
int pgb_hsell_set_tail_warps(int warps);

/* ---- K1/K7: degrees and normalisation (preprocessing.py:104-138; numpy.py:76-77) ----- */
/* out[i] = sum of values (or entry count when values == NULL) of row i, fp64 */
int pgb_csr_row_sums(int64_t n, const int32_t *indptr, const double *values, double *out, void *stream);
/* out = kind(S) with zeros kept zero: S[S != 0] = 1/S[S != 0], after sqrt for RSQRT */
int pgb_make_scales(int64_t n, const double *sums, int kind, double *out, void *stream);
/* data[k] = (left[i] * a_ik) * right[col k] — the normalised CSR exactly as scipy's two
 * diagonal products round it (preprocessing.py:113,138); values == NULL means a_ik = 1 */
int pgb_csr_normalized_values(int64_t n, const int32_t *indptr, const int32_t *indices, const double *values,
                              const double *left, const double *right, double *out_data, void *stream);
/* numpy-compatible pairwise row sums of explicit data (np.add.reduceat as called by
 * scipy's CSR sum(axis=1), numpy.py:76-77); reverse != 0 walks each row backwards (the
 * storage order left by one diagonal product, i.e. "col" normalisation) */
int pgb_csr_row_sums_numpy(int64_t n, const int32_t *indptr, const double *data, int reverse, double *out,
                           void *stream);

/* ---- K2/K4/K5: fused propagation steps ------------------------------------------------ */
/* Plain conv (numpy.py:64-65): y_i = rscale_i * sum_j a_ji z_j, optionally y_i = x_i - that
 * (laplacian, preprocessing.py:122).  z is the PRE-SCALED input (L .* x, see pgb_scale).
 * out_perm != NULL scatters y_i to out[out_perm[i]]. */
int pgb_spmv(const pgb_csr *g, int dtype, const void *z, const void *rscale, const void *x_for_laplacian,
             const int32_t *out_perm, void *out, pgb_span_ws ws, void *stream);

/* Affine recursion in the scaled domain z = ranks / sq (PageRank adhoc.py:34-36 and
 * AbsorbingWalks adhoc.py:166-169 with RecursiveGraphFilter._step's quotient,
 * abstract_filters.py:126-136, and the Mabs check, convergence.py:96-101, fused):
 *     z'_i = (alpha * w_i * sum_j a_ji z_j + q_i) * invS
 *     err += sq_i * |z'_i - z_i|   (or its square)      T += z'_i * c_i
 * w / sq may be NULL for symmetric unweighted graphs: w_i = 1/deg_i, sq_i = sqrt(deg_i)
 * are then derived from the row pointers (deg 0 -> w = 0, sq = 1).
 * Enqueues `num_launches` steps alternating zbuf0/zbuf1 (step k reads buf[(k-1)&1], writes
 * buf[k&1], k = first_step..); a step whose state says STOP != RUNNING exits untouched.
 * finalize != 0: the last CTA of each launch updates the state (single GPU); 0: caller
 * all-reduces the accumulators and calls pgb_state_finalize (row-partitioned multi-GPU). */
int pgb_affine_steps(const pgb_csr *g, int dtype, double alpha, const void *w, const void *sq, const void *c,
                     const void *q, void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64,
                     int32_t *state_i32, double *err_hist, pgb_span_ws ws, int first_step, int num_launches,
                     int finalize, void *stream);

/* Closed-form (polynomial) filter step (ClosedFormGraphFilter._step, abstract_filters.py:248-256,
 * taylor recursion :225-228, power advance :241-246), in the scaled domain zp = power / sq:
 *     ranks_i += coef[k] * sq_i * zp_i ;  err += |delta ranks_i| ;  zp'_i = w_i * sum_j a_ji zp_j
 * coef is a device array indexed by step (coef[k] for k >= 1). */
int pgb_poly_steps(const pgb_csr *g, int dtype, const void *w, const void *sq, const double *coef,
                   void *ranks, void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64,
                   int32_t *state_i32, double *err_hist, pgb_span_ws ws, int first_step, int num_launches,
                   int finalize, void *stream);

/* K3 — batched affine recursion (NodeRanking.propagate, core/signals.py:225-226, as one SpMM-like
 * pass): PB = pgb_panel_width(dtype) seed columns (8 x fp32 / 4 x fp64 = one 32-byte sector per node)
 * advance together.  Vectors z, q are [n][PB] row-major; w/sq/c are per-row as in pgb_affine_steps.
 * state_f64 is [PB][PGB_STATE_F64_LEN]; state_i32 is [PB][PGB_STATE_I32_LEN] followed by one shared
 * ticket word; err_hist is [PB][hist_stride].  A column whose STOP != RUNNING is frozen (z' = z), so
 * every column ends at the iteration where the reference's per-column loop would have stopped.
 * span workspace: acc [n_tiles][PB] doubles, cnt [n_tiles]. */
int pgb_panel_width(int dtype);
int pgb_affine_steps_batched(const pgb_csr *g, int dtype, double alpha, const void *w, const void *sq, const void *c,
                             const void *q, void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64,
                             int32_t *state_i32, double *err_hist, int32_t hist_stride, pgb_span_ws ws,
                             int first_step, int num_launches, void *stream);

/* K3 on the hub-blocked form — the fast path of NodeRanking.propagate (core/signals.py:225-226: one solve per feature
 * column) and of alpha sweeps (algorithms/autotune/optimization.py:160-180 evaluates one candidate after the other).
 * PB = pgb_hsell_panel_width(dtype) columns (4 x fp32 / 2 x fp64 = 16 bytes per node) advance together through a
 * pgb_hsell built with block_cols = pgb_hsell_panel_block_cols() (8192 nodes = 128 KB of [node][PB] per hub block; the
 * builders do not depend on the element type).  The gather kernel is the single-vector one instantiated on 16-byte
 * elements: one 16-bit hub index feeds one LDS.128, one tail index one 16-byte texel, a piece leaves as one 16-byte RED
 * per lane into yacc [(n_slices + 1) * 32][PB] (zero at the first launch; the update pass re-zeroes it).  Every column
 * has its own multiplier (PGB_SF_AMUL), normaliser, error sum and stop decision.
 *
 * The PB columns are SLOTS of a job of n_cols seed columns, scheduled on the device: before every step a column that
 * has stopped is written to `out` (scaled back by its norm when preserve_norm) with its iteration count, stop reason,
 * step count and error history, and the next pending column is loaded into the free slot — the start of
 * GraphFilter.rank, abstract_filters.py:52-56: norm = sum|col| ; pn = col/norm ; z = pn/sq ; q = coef*pn/sq ; state —
 * while the other slots keep iterating.  The host enqueues steps in chunks and polls sched[1] (columns finished);
 * launches after the last column has left are no-ops.  A column with a zero norm leaves with iteration 0 (:53-54). */
typedef struct {
    int32_t n_cols;           /* seed columns of the job                                                         */
    int32_t hist_stride;      /* doubles per row of err_hist / col_err (>= max_iters + 2)                        */
    const void *cols;         /* features, user order: element (i, j) at cols[i*row_stride + j*col_stride]       */
    int64_t row_stride, col_stride;
    void *out;                /* results: element (i, j) at out[i*out_row_stride + j*out_col_stride]             */
    int64_t out_row_stride, out_col_stride;
    const int32_t *perm;      /* internal row i is user node perm[i] (or NULL)                                   */
    const void *sq;           /* per internal row (required)                                                     */
    const void *coefvec;      /* per-row coefficient of the personalization (AbsorbingWalks) or NULL -> coef     */
    const double *col_params; /* [n_cols][3] = (multiplier, alpha_s, coef) per column, or NULL -> the three below */
    double alpha, alpha_s, coef;
    double tol, mean;         /* PGB_SF_TOL, PGB_SF_MEAN of every column                                         */
    int32_t max_iters, end_modulo, err_mode, quotient, preserve_norm;
    int32_t *sched;           /* [4], zero at the start: next column to load, columns finished, spare, spare     */
    int32_t *slot_col;        /* [PB], -1 at the start: column held by every slot                                */
    int32_t *slot_plan;       /* [2*PB] scratch                                                                  */
    double *plan_norm;        /* [PB] scratch                                                                    */
    int32_t *col_result;      /* [n_cols][4]: iteration, stop reason, steps, norm > 0                            */
    double *col_err;          /* [n_cols][hist_stride] error histories (entries 1..steps) or NULL                */
    /* polynomial (closed-form) filters, ClosedFormGraphFilter._step (abstract_filters.py:225-256): every slot accumulates
     * ranks[:, slot] += coef_table[k] * power_k (k = the slot's own step) while its power advances, and leaves with ranks */
    int32_t poly, reserved0;  /* 0: affine recursion; 1: polynomial accumulation (quotient must be 0)            */
    void *ranks;              /* poly: [n][PB] accumulated results (any content: a loaded slot starts from 0)   */
    const double *coef_table; /* poly: device table, coefficient of step k at coef_table[k], k = 1..max_iters    */
} pgb_panel_job;

/* Enqueues `num_launches` steps (step k reads buf[(k-1)&1], writes buf[k&1], k = first_step..).  z, q are [n][PB]
 * row-major (zero at the start), w / c per row as in pgb_affine_steps (w = NULL: w and sq derived from `indptr`).
 * state_f64 [PB][PGB_STATE_F64_LEN]; state_i32 [PB][PGB_STATE_I32_LEN] with STOP != RUNNING at the start, followed by
 * FOUR shared words (ticket; panel stop word, != RUNNING at the start; steps executed; spare); err_hist
 * [PB][hist_stride].  tail_queue: one zeroed uint32.  Single-GPU forms only, accumulate mode, no in-kernel dropout. */
int pgb_hsell_panel_width(int dtype);
int pgb_hsell_panel_block_cols(void);
int pgb_affine_steps_panel(const pgb_hsell *h, const int32_t *indptr, int dtype, const pgb_panel_job *job, const void *w,
                           const void *c, void *q, void *zbuf0, void *zbuf1, double *state_f64, int32_t *state_i32,
                           double *err_hist, void *yacc, uint32_t *tail_queue, int first_step, int num_launches,
                           void *stream);

/* ---- row-partitioned multi-GPU: the exchange fused into the step (no reference counterpart) ----------
 * Every rank keeps the full gather vector (both buffers) and a slot array for the convergence sums in
 * PEER-MAPPED memory (e.g. torch symmetric memory over NVLink / NVSwitch).  The update kernel then
 * writes each new z_i straight into the buffers of all ranks — one multimem.st through the NVSwitch
 * multicast address when mc_zbuf* is set, else one store per peer — and its last CTA writes this rank's
 * (normaliser, error) sums into slot `rank` of every peer's acc array, so the per-iteration all-gather
 * and all-reduce disappear; the caller only needs a cross-device barrier before the next step
 * (which pgb_state_finalize_peer must follow: it adds the slots in rank order, identically everywhere). */
#define PGB_MAX_PEERS 16
typedef struct pgb_peers {
    int32_t n, rank;
    void *zbuf0[PGB_MAX_PEERS];  /* peer-mapped address of every rank's gather-vector buffer 0 (own included) */
    void *zbuf1[PGB_MAX_PEERS];
    void *mc_zbuf0, *mc_zbuf1;   /* multicast addresses of the same buffers, or NULL                          */
    double *acc[PGB_MAX_PEERS];  /* peer-mapped [n][2] slot arrays                                            */
    const uint32_t *row_mask;    /* [local rows] bit r set: rank r reads this row's value (its rows reference the
                                  * column, or it lies in a hub block, or r is the owner); NULL = send to all.
                                  * Unreferenced entries of a peer's buffer are never read, so they stay stale:
                                  * on RMAT-27 / 8 ranks this cuts the bytes every rank receives by ~3x.      */
} pgb_peers;
/* pgb_affine_steps for ONE step (step k reads zbuf[(k-1)&1] locally, writes zbuf[k&1] everywhere);
 * zbuf0/zbuf1 are this rank's own mappings of the buffers named in `peers`. */
int pgb_affine_step_peer(const pgb_csr *g, int dtype, double alpha, const void *w, const void *sq, const void *c,
                         const void *q, void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64,
                         int32_t *state_i32, double *err_hist, pgb_span_ws ws, int step, const pgb_peers *peers,
                         void *stream);
/* pgb_poly_steps for ONE step with the same fused exchange (closed-form filters on a row-partitioned graph). */
int pgb_poly_step_peer(const pgb_csr *g, int dtype, const void *w, const void *sq, const double *coef, void *ranks,
                       void *zbuf0, void *zbuf1, int64_t out_offset, double *state_f64, int32_t *state_i32,
                       double *err_hist, pgb_span_ws ws, int step, const pgb_peers *peers, void *stream);
/* pgb_affine_init whose start vector z0 (this rank's rows) is written into buffer 0 of EVERY rank — replaces
 * the all-gather of the start vector; the caller puts a cross-device barrier before the first step. */
int pgb_affine_init_peer(int64_t n, int dtype, const void *p, const void *warm, const void *sq, const void *c,
                         double coef, const void *coefvec, const int32_t *perm, int64_t out_offset, void *q,
                         double *state_f64, const pgb_peers *peers, void *stream);
/* acc_slots: this rank's own [n][2] slot array (filled by every rank's update kernel). */
int pgb_state_finalize_peer(double *state_f64, int32_t *state_i32, double *err_hist, const double *acc_slots,
                            int32_t n, void *stream);

/* One-thread state update for the deferred (multi-GPU) mode. */
int pgb_state_finalize(double *state_f64, int32_t *state_i32, double *err_hist, void *stream);

/* ---- vector helpers around the fused steps (K6) ----------------------------------------- */
/* out[i] = a[src] * b[src] * scale with src = perm ? perm[i] : i   (b may be NULL) */
int pgb_scale(int64_t n, int dtype, const void *a, const void *b, double scale, const int32_t *perm,
              void *out, void *stream);
/* out[dst] = a[i] * b[i] * (*dev_scale or 1) * scale with dst = perm ? perm[i] : i */
int pgb_unscale(int64_t n, int dtype, const void *a, const void *b, const double *dev_scale, double scale,
                const int32_t *perm, void *out, void *stream);
/* sums[0] += sum|x_i|, sums[1] += sum x_i, sums[2] += sum x_i*y_i (y may be NULL); fp64 */
int pgb_reduce3(int64_t n, int dtype, const void *x, const void *y, double *sums, void *stream);
/* Start of GraphFilter.rank (abstract_filters.py:52-56) in the engine's domain: with
 * norm = state[NORM] (set from pgb_reduce3) and src = perm ? perm[i] : i,
 *     pn = p[src]/norm ; z0[out_offset+i] = (warm ? warm[src] : pn) / sq_i ;
 *     q_i = (coefvec ? coefvec[i] : coef) * pn / sq_i          (q may be NULL)
 * and accumulates state[TACC] += z0_i*c_i (c may be NULL), state[BIAS] += q_i*sq_i. */
int pgb_affine_init(int64_t n, int dtype, const void *p, const void *warm, const void *sq, const void *c,
                    double coef, const void *coefvec, const int32_t *perm, int64_t out_offset, void *z0, void *q,
                    double *state_f64, void *stream);
/* invS for the first step from the accumulated T / BIAS (one thread). */
int pgb_affine_init_finish(double *state_f64, int32_t *state_i32, void *stream);

/* Staging of a panel job.  A slot loads / stores one column at a time; from a row-major [n][B] matrix in user order
 * that is one 32-byte sector per value at random rows, so jobs run on column-major blocks in the engine's row order:
 * direction 0: stage[j][i] = matrix[perm[i]*row_stride + (j0+j)*col_stride]   (j < n_cols; perm NULL = identity)
 * direction 1: matrix[perm[i]*row_stride + (j0+j)] = stage[j][i]              (col_stride must be 1)
 * (tiled transposes: both sides move in full lines).  The job then uses cols = stage, row_stride 1, col_stride n,
 * perm NULL, and out = a second stage with out_row_stride 1, out_col_stride n. */
int pgb_panel_stage(int64_t n, int dtype, int direction, void *matrix, int64_t row_stride, int64_t col_stride,
                    const int32_t *perm, int64_t j0, int32_t n_cols, void *stage, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PGB200_H */
