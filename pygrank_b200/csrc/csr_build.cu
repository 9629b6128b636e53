// K1 — on-device graph -> canonical CSR, degrees, normalisation scales, merge-path partition.
// Replaces the host pipeline of /root/reference/pygrank/core/utils/preprocessing.py:103-138
// (nx.to_scipy_sparse_array / coo.tocsr, two diagonal SpGEMMs) — see include/pgb200.h.
// Sorting/compaction primitives come from CUB (part of the CUDA toolkit, library code); the
// key construction, row-pointer fill, scaling and partition kernels are ours.
#include <cub/cub.cuh>

#include "common.cuh"

namespace pgb {

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BuildLayout {
    size_t keys_a, keys_b, vals_a, vals_b, counters, cub_temp, cub_bytes, total;
};

static int key_bits(int64_t n) {
    int bits = 1;
    while ((1ll << bits) <= n) ++bits;  // must represent the value n itself (dropped-entry row)
    return 32 + bits;
}

static BuildLayout build_layout(int64_t n, int64_t nnz_in, int flags, int weighted) {
    const int64_t m = nnz_in * ((flags & PGB_BUILD_SYMMETRIZE) ? 2 : 1);
    const bool with_vals = weighted && !(flags & PGB_BUILD_BINARY);
    BuildLayout L;
    size_t off = 0;
    L.keys_a = off; off += align256((size_t)m * 8);
    L.keys_b = off; off += align256((size_t)m * 8);
    L.vals_a = off; off += with_vals ? align256((size_t)m * 8) : 0;
    L.vals_b = off; off += with_vals ? align256((size_t)m * 8) : 0;
    L.counters = off; off += 256;
    size_t t_sort = 0, t_sel = 0;
    cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
    if (with_vals) {
        cub::DoubleBuffer<double> dv(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, t_sort, dk, dv, m, 0, key_bits(n));
        cub::DeviceReduce::ReduceByKey(nullptr, t_sel, (uint64_t *)nullptr, (uint64_t *)nullptr, (double *)nullptr,
                                       (double *)nullptr, (int64_t *)nullptr, cub::Sum(), m);
    } else {
        cub::DeviceRadixSort::SortKeys(nullptr, t_sort, dk, m, 0, key_bits(n));
        cub::DeviceSelect::Unique(nullptr, t_sel, (uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t *)nullptr, m);
    }
    L.cub_bytes = (t_sort > t_sel ? t_sort : t_sel) + 256;
    L.cub_temp = off; off += align256(L.cub_bytes);
    L.total = off;
    return L;
}

__global__ void make_keys_kernel(int64_t n, int64_t nnz_in, const int32_t *__restrict__ row,
                                 const int32_t *__restrict__ col, const double *__restrict__ val, int flags,
                                 uint64_t *__restrict__ keys, double *__restrict__ vals) {
    const bool sym = flags & PGB_BUILD_SYMMETRIZE, drop = flags & PGB_BUILD_DROP_SELF_LOOPS;
    const uint64_t dropped = (uint64_t)n << 32;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz_in;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)row[i], c = (uint32_t)col[i];
        const bool bad = (r >= (uint64_t)n) || (c >= (uint64_t)n) || (drop && r == c);
        keys[i] = bad ? dropped : (((uint64_t)r << 32) | c);
        if (vals) vals[i] = val ? val[i] : 1.0;
        if (sym) {
            keys[nnz_in + i] = (bad || r == c) ? dropped : (((uint64_t)c << 32) | r);  // a self loop is stored once
            if (vals) vals[nnz_in + i] = val ? val[i] : 1.0;
        }
    }
}

// counters[0] = number of unique keys (incl. the dropped group); counters[1] <- nnz_out
__global__ void count_valid_kernel(int64_t n, const uint64_t *__restrict__ uniq, int64_t *counters) {
    int64_t u = counters[0];
    if (u > 0 && (uniq[u - 1] >> 32) >= (uint64_t)n) --u;
    counters[1] = u;
}

__global__ void emit_csr_kernel(int64_t n, const uint64_t *__restrict__ uniq, const double *__restrict__ sums,
                                const int64_t *__restrict__ counters, int32_t *__restrict__ indptr,
                                int32_t *__restrict__ indices, double *__restrict__ values) {
    const int64_t nnz = counters[1];
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k <= nnz;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r_prev = (k == 0) ? -1 : (int64_t)(uniq[k - 1] >> 32);
        const int64_t r_cur = (k == nnz) ? n : (int64_t)(uniq[k] >> 32);
        for (int64_t r = r_prev + 1; r <= r_cur; ++r) indptr[r] = (int32_t)k;
        if (k < nnz) {
            indices[k] = (int32_t)(uniq[k] & 0xFFFFFFFFull);
            if (values) values[k] = sums ? sums[k] : 1.0;
        }
    }
}

__global__ void expand_rows_kernel(int64_t n, const int32_t *__restrict__ indptr, int32_t *__restrict__ rows) {
    // one warp per row keeps the writes coalesced for long rows; short rows waste lanes (build-time only)
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps)
        for (int32_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) rows[k] = (int32_t)r;
}

__global__ void degree_keys_kernel(int64_t n, const int32_t *__restrict__ indptr, int32_t *__restrict__ deg,
                                   int32_t *__restrict__ ids) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        deg[i] = indptr[i + 1] - indptr[i];
        ids[i] = (int32_t)i;
    }
}

__global__ void invert_perm_kernel(int64_t n, const int32_t *__restrict__ perm, int32_t *__restrict__ iperm) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        iperm[perm[i]] = (int32_t)i;
}

// Hub-block signature of every row (original labels): bit b (b = 0 most significant of word 0) is set when the
// row has an entry whose degree RANK (iperm[col]) lies in [b*G, (b+1)*G); ranks >= span set nothing.  One warp
// per row.  sig is [words][n].
__global__ void block_signature_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                       const int32_t *__restrict__ iperm, int32_t G, int64_t span, int words,
                                       uint64_t *__restrict__ sig) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        uint64_t acc[PGB_SIGNATURE_WORDS];
#pragma unroll
        for (int w = 0; w < PGB_SIGNATURE_WORDS; ++w) acc[w] = 0ull;
        for (int32_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) {
            const int32_t c = iperm ? iperm[indices[k]] : indices[k];
            if (c < span) {
                const int b = c / G;
#pragma unroll
                for (int w = 0; w < PGB_SIGNATURE_WORDS; ++w)
                    if ((b >> 6) == w) acc[w] |= 1ull << (63 - (b & 63));
            }
        }
#pragma unroll
        for (int w = 0; w < PGB_SIGNATURE_WORDS; ++w) {
            if (w < words) {
                const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)acc[w]);
                const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(acc[w] >> 32));
                if (lane == 0) sig[(int64_t)w * n + r] = ((uint64_t)hi << 32) | lo;
            }
        }
    }
}

// keys of one pass of the signature sort for the current order `ids`: the complemented signature word (rows
// that touch a block sort first), or (sig == NULL) the region of the node's degree rank
__global__ void signature_keys_kernel(int64_t n, const int32_t *__restrict__ ids, const uint64_t *__restrict__ sig,
                                      const int32_t *__restrict__ iperm, int32_t G, int64_t span,
                                      uint64_t *__restrict__ keys) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = ids[i];
        if (sig) {
            keys[i] = ~sig[v];
        } else {
            const int64_t rank = iperm[v];
            keys[i] = (uint64_t)((rank < span ? rank : span) / G);
        }
    }
}

__global__ void relabel_kernel(int64_t nnz, const int32_t *__restrict__ iperm, int32_t *__restrict__ row,
                               int32_t *__restrict__ col) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x) {
        row[i] = iperm[row[i]];
        col[i] = iperm[col[i]];
    }
}

// tile_row[t] = min r in [0, n] with indptr[r+1] + r >= t * items  (row whose end marker is the
// first one at or after the tile's first merge item); tile_row[n_tiles] = n
__global__ void mergepath_partition_kernel(int64_t n, const int32_t *__restrict__ indptr, int32_t n_tiles,
                                           int64_t items, int32_t *__restrict__ tile_row) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= n_tiles;
         t += (int64_t)gridDim.x * blockDim.x) {
        if (t == n_tiles) {
            tile_row[t] = (int32_t)n;
            continue;
        }
        const int64_t target = t * items;
        int64_t lo = 0, hi = n;  // answer in [lo, hi]
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)indptr[mid + 1] + mid >= target)
                hi = mid;
            else
                lo = mid + 1;
        }
        tile_row[t] = (int32_t)lo;
    }
}

__global__ void row_sums_kernel(int64_t n, const int32_t *__restrict__ indptr, const double *__restrict__ values,
                                double *__restrict__ out) {
    // warp per row, fp64; exact for unweighted/integer weights (every BASELINE config)
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int32_t b = indptr[r], e = indptr[r + 1];
        double s = 0.0;
        if (values)
            for (int32_t k = b + lane; k < e; k += 32) s += values[k];
        else if (lane == 0)
            s = (double)(e - b);
        s = warp_sum(s);
        if (lane == 0) out[r] = s;
    }
}

__global__ void make_scales_kernel(int64_t n, const double *__restrict__ sums, int kind, double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = sums[i];
        if (kind == PGB_SCALE_ONE) {
            out[i] = 1.0;
            continue;
        }
        if (kind == PGB_SCALE_RSQRT) s = sqrt(s);       // IEEE sqrt, as np.sqrt
        out[i] = (s != 0.0) ? __ddiv_rn(1.0, s) : s;      // S[S != 0] = 1.0 / S[S != 0]
    }
}

__global__ void normalized_values_kernel(int64_t n, const int32_t *__restrict__ indptr,
                                         const int32_t *__restrict__ indices, const double *__restrict__ values,
                                         const double *__restrict__ left, const double *__restrict__ right,
                                         double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const double l = left ? left[r] : 1.0;
        for (int32_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) {
            double v = values ? values[k] : 1.0;
            if (left) v = __dmul_rn(l, v);                       // (left[i] * a_ik) first ...
            if (right) v = __dmul_rn(v, right[indices[k]]);      // ... then * right[k]
            out[k] = v;
        }
    }
}

// numpy's pairwise summation (the add.reduce inner loop np.add.reduceat runs per segment):
// n < 8 sequential; n <= 128 eight interleaved accumulators combined as a fixed tree plus a
// sequential tail; else split at (n/2 rounded down to a multiple of 8) and recurse.
__device__ __forceinline__ double numpy_pairwise_block(const double *a, int64_t n, int64_t stride) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i * stride];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j * stride];
    int64_t i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[(i + j) * stride];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i * stride];
    return res;
}

// The recursion sum(a, n) = sum(a, n2) + sum(a + n2, n - n2) (n > 128) is evaluated with an explicit
// post-order stack (depth <= 40 covers any int64 length) — device recursion would overflow the
// default 1 KB thread stack on hub rows.
__device__ double numpy_pairwise(const double *a, int64_t n, int64_t stride) {
    if (n <= 128) return numpy_pairwise_block(a, n, stride);
    int64_t off[40], len[40];
    double left[40];
    int stage[40];
    int sp = 0;
    off[0] = 0; len[0] = n; stage[0] = 0; left[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        if (len[sp] <= 128) {
            ret = numpy_pairwise_block(a + off[sp] * stride, len[sp], stride);
            --sp;
            continue;
        }
        int64_t n2 = len[sp] / 2;
        n2 -= n2 % 8;
        if (stage[sp] == 0) {          // descend into the left half
            stage[sp] = 1;
            off[sp + 1] = off[sp]; len[sp + 1] = n2; stage[sp + 1] = 0;
            ++sp;
        } else if (stage[sp] == 1) {   // left half done -> descend into the right half
            left[sp] = ret;
            stage[sp] = 2;
            off[sp + 1] = off[sp] + n2; len[sp + 1] = len[sp] - n2; stage[sp + 1] = 0;
            ++sp;
        } else {                       // both halves done
            ret = left[sp] + ret;
            --sp;
        }
    }
    return ret;
}

__global__ void row_sums_numpy_kernel(int64_t n, const int32_t *__restrict__ indptr, const double *__restrict__ data,
                                      int reverse, double *__restrict__ out) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = indptr[r], e = indptr[r + 1];
        if (e == b) {
            out[r] = 0.0;
            continue;
        }
        // np.add.reduceat(data, starts): segment = data[b:e]; the ufunc reduce takes data[b] as the
        // initial value and pairwise-sums the remaining e-b-1 elements INTO it in one inner-loop call,
        // which for the contiguous add loop is pairwise(data[b+1:e]) added to data[b].
        if (!reverse)
            out[r] = data[b] + numpy_pairwise(data + b + 1, e - b - 1, 1);
        else
            out[r] = data[e - 1] + numpy_pairwise(data + e - 2, e - b - 1, -1);
    }
}

}  // namespace pgb

using namespace pgb;

extern "C" {

size_t pgb_csr_build_workspace_bytes(int64_t n, int64_t nnz_in, int flags, int weighted) {
    if (nnz_in <= 0) return 512;
    return build_layout(n, nnz_in, flags, weighted).total;
}

int pgb_csr_build(int64_t n, int64_t nnz_in, const int32_t *row, const int32_t *col, const double *val, int flags,
                  void *workspace, size_t workspace_bytes, int32_t *out_indptr, int32_t *out_indices,
                  double *out_values, int64_t *out_nnz_host, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (n < 0 || n >= (1ll << 31) - 1) return fail("pgb_csr_build: n=%lld out of range", (long long)n);
    const int64_t m = nnz_in * ((flags & PGB_BUILD_SYMMETRIZE) ? 2 : 1);
    if (m >= (1ll << 31)) return fail("pgb_csr_build: %lld entries exceed the int32 limit of one device", (long long)m);
    if (nnz_in <= 0) {
        PGB_CUDA_OK(cudaMemsetAsync(out_indptr, 0, (size_t)(n + 1) * 4, st));
        if (out_nnz_host) *out_nnz_host = 0;
        return 0;
    }
    const bool weighted = (val != nullptr);
    const bool with_vals = weighted && !(flags & PGB_BUILD_BINARY);
    if (with_vals && !out_values) return fail("pgb_csr_build: weighted build needs out_values");
    const BuildLayout L = build_layout(n, nnz_in, flags, weighted);
    if (workspace_bytes < L.total)
        return fail("pgb_csr_build: workspace %zu < required %zu bytes", workspace_bytes, L.total);
    char *ws = (char *)workspace;
    uint64_t *keys_a = (uint64_t *)(ws + L.keys_a), *keys_b = (uint64_t *)(ws + L.keys_b);
    double *vals_a = with_vals ? (double *)(ws + L.vals_a) : nullptr;
    double *vals_b = with_vals ? (double *)(ws + L.vals_b) : nullptr;
    int64_t *counters = (int64_t *)(ws + L.counters);
    void *cub_temp = ws + L.cub_temp;
    size_t cub_bytes = L.cub_bytes;

    make_keys_kernel<<<stride_grid(nnz_in, 256), 256, 0, st>>>(n, nnz_in, row, col, val, flags, keys_a, vals_a);
    PGB_LAUNCH_OK("make_keys_kernel");

    cub::DoubleBuffer<uint64_t> dk(keys_a, keys_b);
    const uint64_t *uniq = nullptr;
    const double *sums = nullptr;
    if (with_vals) {
        cub::DoubleBuffer<double> dv(vals_a, vals_b);
        PGB_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, dk, dv, m, 0, key_bits(n), st));
        cub_bytes = L.cub_bytes;
        PGB_CUDA_OK(cub::DeviceReduce::ReduceByKey(cub_temp, cub_bytes, dk.Current(), dk.Alternate(), dv.Current(),
                                                   dv.Alternate(), counters, cub::Sum(), m, st));
        uniq = dk.Alternate();
        sums = dv.Alternate();
    } else {
        PGB_CUDA_OK(cub::DeviceRadixSort::SortKeys(cub_temp, cub_bytes, dk, m, 0, key_bits(n), st));
        cub_bytes = L.cub_bytes;
        PGB_CUDA_OK(cub::DeviceSelect::Unique(cub_temp, cub_bytes, dk.Current(), dk.Alternate(), counters, m, st));
        uniq = dk.Alternate();
    }
    count_valid_kernel<<<1, 1, 0, st>>>(n, uniq, counters);
    PGB_LAUNCH_OK("count_valid_kernel");
    emit_csr_kernel<<<stride_grid(m + 1, 256), 256, 0, st>>>(n, uniq, sums, counters, out_indptr, out_indices,
                                                             out_values);
    PGB_LAUNCH_OK("emit_csr_kernel");
    int64_t host_counters[2] = {0, 0};
    PGB_CUDA_OK(cudaMemcpyAsync(host_counters, counters, sizeof(host_counters), cudaMemcpyDeviceToHost, st));
    PGB_CUDA_OK(cudaStreamSynchronize(st));
    if (out_nnz_host) *out_nnz_host = host_counters[1];
    return 0;
}

int pgb_csr_expand_rows(int64_t n, int64_t nnz, const int32_t *indptr, int32_t *rows, void *stream) {
    if (n <= 0 || nnz <= 0) return 0;
    expand_rows_kernel<<<stride_grid(n * 32, 256), 256, 0, as_stream(stream)>>>(n, indptr, rows);
    PGB_LAUNCH_OK("expand_rows_kernel");
    return 0;
}

size_t pgb_degree_order_workspace_bytes(int64_t n) {
    size_t t = 0;
    cub::DoubleBuffer<int32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, t, dk, dv, n);
    return align256(t + 256) + 4 * align256((size_t)n * 4);
}

int pgb_degree_order(int64_t n, const int32_t *indptr, void *workspace, size_t workspace_bytes, int32_t *perm,
                     int32_t *iperm, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (n <= 0) return 0;
    if (workspace_bytes < pgb_degree_order_workspace_bytes(n)) return fail("pgb_degree_order: workspace too small");
    char *ws = (char *)workspace;
    const size_t seg = align256((size_t)n * 4);
    int32_t *deg_a = (int32_t *)ws, *deg_b = (int32_t *)(ws + seg), *id_a = (int32_t *)(ws + 2 * seg),
            *id_b = (int32_t *)(ws + 3 * seg);
    void *cub_temp = ws + 4 * seg;
    size_t cub_bytes = workspace_bytes - 4 * seg;
    degree_keys_kernel<<<stride_grid(n, 256), 256, 0, st>>>(n, indptr, deg_a, id_a);
    PGB_LAUNCH_OK("degree_keys_kernel");
    cub::DoubleBuffer<int32_t> dk(deg_a, deg_b), dv(id_a, id_b);
    PGB_CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(cub_temp, cub_bytes, dk, dv, n, 0, 32, st));
    PGB_CUDA_OK(cudaMemcpyAsync(perm, dv.Current(), (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    invert_perm_kernel<<<stride_grid(n, 256), 256, 0, st>>>(n, perm, iperm);
    PGB_LAUNCH_OK("invert_perm_kernel");
    return 0;
}

size_t pgb_hub_order_workspace_bytes(int64_t n, int32_t words) {
    size_t t = 0;
    cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, t, dk, dv, n);
    return align256(t + 256) + (size_t)(words + 2) * align256((size_t)n * 8) + 2 * align256((size_t)n * 4);
}

int pgb_hub_order(int64_t n, const int32_t *indptr, const int32_t *indices, int32_t G, int64_t span, void *workspace,
                  size_t workspace_bytes, int32_t *perm, int32_t *iperm, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (n <= 0) return 0;
    if (G < 1 || span < 0) return fail("pgb_hub_order: bad block size / span");
    const int64_t bits = ceil_div(span, G);
    const int words = (int)ceil_div(bits, 64);
    if (words > PGB_SIGNATURE_WORDS)
        return fail("pgb_hub_order: %lld signature bits exceed the %d supported", (long long)bits, 64 * PGB_SIGNATURE_WORDS);
    if (workspace_bytes < pgb_hub_order_workspace_bytes(n, words)) return fail("pgb_hub_order: workspace too small");
    char *ws = (char *)workspace;
    const size_t seg8 = align256((size_t)n * 8), seg4 = align256((size_t)n * 4);
    uint64_t *sig = (uint64_t *)ws;
    uint64_t *key_a = (uint64_t *)(ws + (size_t)words * seg8), *key_b = (uint64_t *)(ws + (size_t)(words + 1) * seg8);
    int32_t *id_a = (int32_t *)(ws + (size_t)(words + 2) * seg8), *id_b = (int32_t *)(ws + (size_t)(words + 2) * seg8 + seg4);
    void *cub_temp = ws + (size_t)(words + 2) * seg8 + 2 * seg4;
    size_t cub_bytes = workspace_bytes - ((size_t)(words + 2) * seg8 + 2 * seg4);
    if (words > 0) {
        block_signature_kernel<<<stride_grid(n * 32, 256), 256, 0, st>>>(n, indptr, indices, iperm, G, span, words, sig);
        PGB_LAUNCH_OK("block_signature_kernel");
    }
    // LSD passes of a stable sort: degree rank (the incoming perm) < signature words (last word first) < region
    PGB_CUDA_OK(cudaMemcpyAsync(id_a, perm, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    cub::DoubleBuffer<int32_t> dv(id_a, id_b);
    for (int pass = words - 1; pass >= -1; --pass) {
        cub::DoubleBuffer<uint64_t> dk(key_a, key_b);
        signature_keys_kernel<<<stride_grid(n, 256), 256, 0, st>>>(n, dv.Current(), pass >= 0 ? sig + (int64_t)pass * n : nullptr,
                                                                  iperm, G, span, dk.Current());
        PGB_LAUNCH_OK("signature_keys_kernel");
        size_t tb = cub_bytes;
        PGB_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_temp, tb, dk, dv, n, 0, pass >= 0 ? 64 : 32, st));
    }
    PGB_CUDA_OK(cudaMemcpyAsync(perm, dv.Current(), (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    invert_perm_kernel<<<stride_grid(n, 256), 256, 0, st>>>(n, perm, iperm);
    PGB_LAUNCH_OK("invert_perm_kernel");
    return 0;
}

int pgb_relabel_coo(int64_t nnz, const int32_t *iperm, int32_t *row, int32_t *col, void *stream) {
    if (nnz <= 0) return 0;
    relabel_kernel<<<stride_grid(nnz, 256), 256, 0, as_stream(stream)>>>(nnz, iperm, row, col);
    PGB_LAUNCH_OK("relabel_kernel");
    return 0;
}

int pgb_mergepath_partition(int64_t n, int64_t nnz, const int32_t *indptr, int32_t n_tiles, int32_t *tile_row,
                            void *stream) {
    const int64_t items = pgb_tile_items();
    if ((int64_t)n_tiles != ceil_div(n + nnz, items) && !(n + nnz == 0 && n_tiles == 0))
        return fail("pgb_mergepath_partition: n_tiles=%d but ceil((n+nnz)/%lld)=%lld", n_tiles, (long long)items,
                    (long long)ceil_div(n + nnz, items));
    mergepath_partition_kernel<<<stride_grid(n_tiles + 1, 256), 256, 0, as_stream(stream)>>>(n, indptr, n_tiles, items,
                                                                                              tile_row);
    PGB_LAUNCH_OK("mergepath_partition_kernel");
    return 0;
}

int pgb_csr_row_sums(int64_t n, const int32_t *indptr, const double *values, double *out, void *stream) {
    if (n <= 0) return 0;
    row_sums_kernel<<<stride_grid(n * 32, 256), 256, 0, as_stream(stream)>>>(n, indptr, values, out);
    PGB_LAUNCH_OK("row_sums_kernel");
    return 0;
}

int pgb_make_scales(int64_t n, const double *sums, int kind, double *out, void *stream) {
    if (n <= 0) return 0;
    if (kind < PGB_SCALE_ONE || kind > PGB_SCALE_RSQRT) return fail("pgb_make_scales: unknown kind %d", kind);
    make_scales_kernel<<<stride_grid(n, 256), 256, 0, as_stream(stream)>>>(n, sums, kind, out);
    PGB_LAUNCH_OK("make_scales_kernel");
    return 0;
}

int pgb_csr_normalized_values(int64_t n, const int32_t *indptr, const int32_t *indices, const double *values,
                              const double *left, const double *right, double *out_data, void *stream) {
    if (n <= 0) return 0;
    normalized_values_kernel<<<stride_grid(n * 32, 256), 256, 0, as_stream(stream)>>>(n, indptr, indices, values, left,
                                                                                       right, out_data);
    PGB_LAUNCH_OK("normalized_values_kernel");
    return 0;
}

int pgb_csr_row_sums_numpy(int64_t n, const int32_t *indptr, const double *data, int reverse, double *out,
                           void *stream) {
    if (n <= 0) return 0;
    row_sums_numpy_kernel<<<stride_grid(n, 128), 128, 0, as_stream(stream)>>>(n, indptr, data, reverse, out);
    PGB_LAUNCH_OK("row_sums_numpy_kernel");
    return 0;
}

}  // extern "C"
