// bucket.cu -- rows with 513 .. 65536 intermediate products: one CTA per row, the row's products are
// dealt into column-range BUCKETS of ~100 products in shared memory, every bucket is then sorted by
// one warp in registers (the same network as the warp-per-row bins) and the sorted buckets simply
// concatenate, because their column ranges are disjoint and ascending.
//
// Why: the CTA-wide bitonic network of esc_cta_bitonic.cu spends 17-18 warp instructions per product
// (shared-memory merge stages, a barrier per stage) against 9.5 for the register network, and the
// shared-memory bitmap path of heavy_smem.cu expands every row three times and accumulates with global
// f64 atomics.  Bucketing is one histogram + one scatter per product (shared-memory atomics) and keeps
// all comparisons in registers.
//
// One pass over a row:
//   1. expansion: warps take 32 A entries each, the products are dealt round-robin over the lanes
//      (coalesced B reads); product -> staging slot `pos` (its arrival index), value = fl(a*b),
//      histogram[bucket(col)]++
//   2. exclusive scan of the histogram (<= 128 buckets)
//   3. scatter: key = (col - lo[bucket]) << 13 | pos  into the bucket's segment
//   4. every warp sorts whole buckets in registers (network sized by the bucket: 32..512 keys), counts
//      the distinct columns
//   5. scan of the counts, then every warp sums equal columns left to right (ascending arrival index =
//      the oracle's order: values are bit-identical) and stores its buckets
// Rows with more products than fit the staging area (PMAX) run several passes over disjoint column
// ranges, each pass keeping only the products of its range (their staging order, hence the summation
// order of duplicates, is then not fixed: parity within the stated 1e-12, like the reference whose own
// merge order is schedule dependent, scheduler.rs:386-396).
// A row whose columns are so skewed that a bucket exceeds 512 products or a pass exceeds PMAX is put
// on an overflow list and recomputed by the bitonic / bitmap kernels.
//
// Reference logic replaced: the window of A scalars fanned out over the lanes (scheduler.rs:551-556),
// multiply (simulator.rs:86-111), sort (:143-171), merge-accumulate (:199-230), psum append (:955-983).
#include "cta_common.cuh"

namespace spada {

constexpr int BK_SEQ_BITS = 13;                // staging index of a product inside one pass (< 8192)
constexpr uint32_t BK_SEQ_MASK = (1u << BK_SEQ_BITS) - 1u;
constexpr int BK_COL_BITS = 32 - BK_SEQ_BITS;  // in-bucket column offset
constexpr int BK_NB_MAX = 128;                 // buckets per pass
constexpr int BK_TARGET = 100;                 // mean products per bucket
constexpr int BK_BUCKET_MAX = 512;             // largest register network (E = 16)
constexpr uint64_t BK_WIDTH_MAX = (1ull << BK_COL_BITS) - 256;

// bucket(d) = (d * mult) >> 32 for d = col - c0 in [0, range): monotone, < nb
// (written with 32-bit halves and __umulhi: nvcc 12.9 strength-reduced the plain 64-bit form
//  `&s_cur[(uint64_t)d * mult >> 32]` inside the scatter loop into a wrong 32-bit multiply)
__device__ __forceinline__ uint32_t bk_bucket(uint32_t d, uint64_t mult) {
    return d * (uint32_t)(mult >> 32) + __umulhi(d, (uint32_t)mult);
}

__host__ __device__ inline uint32_t bk_num_buckets(uint32_t products, uint64_t range) {
    uint64_t nb = ((uint64_t)products + BK_TARGET - 1) / BK_TARGET;
    const uint64_t nb_min = (range + BK_WIDTH_MAX - 1) / BK_WIDTH_MAX;  // the in-bucket offset must fit BK_COL_BITS
    if (nb < nb_min) nb = nb_min;
    if (nb > (uint64_t)BK_NB_MAX) nb = BK_NB_MAX;
    if (nb > range) nb = range;
    if (nb < 1) nb = 1;
    return (uint32_t)nb;
}

bool bucket_supported(int64_t b_cols) { return b_cols >= 1 && (uint64_t)b_cols <= BK_WIDTH_MAX * (uint64_t)BK_NB_MAX; }

// sorts the sz (<= 32*E) keys of one bucket in registers, leaves them sorted in place and returns the
// number of distinct columns
template <int E>
__device__ __forceinline__ int bk_sort_count(uint32_t* keys, int sz, int lane) {
    uint32_t x[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = r * 32 + lane;  // striped: conflict free; the network does not care where a key starts
        x[r] = i < sz ? keys[i] : 0xffffffffu;
    }
    warp_sort<uint32_t, E>(x, lane);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = lane * E + r;  // sorted order is the blocked layout
        if (i < sz) keys[i] = x[r];
    }
    const uint32_t prev = __shfl_up_sync(FULL, x[E - 1], 1);
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const uint32_t pv = (r == 0) ? prev : x[r - 1];
        const bool first = (lane == 0 && r == 0);
        if (x[r] != 0xffffffffu && (first || (x[r] >> BK_SEQ_BITS) != (pv >> BK_SEQ_BITS))) ++cnt;
    }
    cnt = __reduce_add_sync(FULL, cnt);
    __syncwarp();
    return cnt;
}

// sums equal columns of one sorted bucket left to right and stores it at c_col/c_val + gbase
__device__ __forceinline__ void bk_reduce_store(const uint32_t* keys, const double* vals, int sz, int lane, int64_t gbase,
                                                uint32_t col_base, int32_t* __restrict__ c_col,
                                                double* __restrict__ c_val) {
    int out_base = 0;
    uint32_t prev_last = 0;
    for (int cb = 0; cb < sz; cb += 32) {
        const int i = cb + lane;
        const bool valid = i < sz;
        const uint32_t ki = valid ? keys[i] : 0xffffffffu;
        const uint32_t col = ki >> BK_SEQ_BITS;
        uint32_t col_prev = __shfl_up_sync(FULL, col, 1);
        if (lane == 0) col_prev = prev_last;
        const bool head = valid && (i == 0 || col_prev != col);
        const unsigned hm = __ballot_sync(FULL, head);
        prev_last = __shfl_sync(FULL, col, 31);
        if (head) {
            double sum = vals[ki & BK_SEQ_MASK];
            for (int j = i + 1; j < sz; ++j) {
                const uint32_t kj = keys[j];
                if ((kj >> BK_SEQ_BITS) != col) break;
                sum = __dadd_rn(sum, vals[kj & BK_SEQ_MASK]);
            }
            const int o = out_base + __popc(hm & ((1u << lane) - 1u));
            c_col[gbase + o] = (int32_t)(col_base + col);
            c_val[gbase + o] = sum;
        }
        out_base += __popc(hm);
    }
}

// exclusive scan of n (<= 128) counters by warp 0: out[i] = sum of in[0..i), out[n] = total
__device__ __forceinline__ uint32_t bk_scan128(const uint32_t* in, uint32_t* out, int n, int lane) {
    uint32_t v[4], s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = lane * 4 + q;
        v[q] = i < n ? in[i] : 0u;
        s += v[q];
    }
    int total;
    uint32_t ex = (uint32_t)warp_excl_scan((int)s, lane, total);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = lane * 4 + q;
        if (i < n) out[i] = ex;
        ex += v[q];
    }
    if (lane == 0) out[n] = (uint32_t)total;
    return (uint32_t)total;
}

template <int PMAX, int THREADS>
__device__ __forceinline__ void bucket_row(const DevCsr& a, const DevCsr& b, int64_t row_begin, uint32_t r, uint32_t p,
                                           const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col,
                                           double* __restrict__ c_val, uint32_t* __restrict__ row_nnz_out,
                                           uint32_t* __restrict__ ovf, unsigned char* s_raw) {
    constexpr int NW = THREADS / 32;
    constexpr int UN = EXPAND_UNROLL;
    constexpr uint32_t PCAP = (uint32_t)PMAX * 3u / 4u;  // expected products of one pass of a multi-pass row
    static_assert(PMAX <= (1 << BK_SEQ_BITS), "staging index must fit the key");
    double* s_val = reinterpret_cast<double*>(s_raw);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_raw + sizeof(double) * PMAX);
    uint32_t* s_key = s_stage + PMAX;
    __shared__ uint32_t s_hist[BK_NB_MAX], s_off[BK_NB_MAX + 1], s_cur[BK_NB_MAX], s_lo[BK_NB_MAX], s_cnt[BK_NB_MAX],
        s_out[BK_NB_MAX + 1];
    __shared__ int s_wtot[NW];
    __shared__ uint32_t s_count, s_ovf;

    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    const int64_t cbase = c_ptr[r];
    const uint64_t n = (uint64_t)b.cols;
    const uint32_t passes = p <= (uint32_t)PMAX ? 1u : (p + PCAP - 1u) / PCAP;
    const bool single = passes == 1u;
    uint32_t row_total = 0;

    for (uint32_t pi = 0; pi < passes; ++pi) {
        const uint32_t c0 = (uint32_t)(n * pi / passes), c1 = (uint32_t)(n * (pi + 1) / passes);
        const uint64_t range = (uint64_t)(c1 - c0);
        if (range == 0) continue;  // uniform: more passes than columns
        const uint32_t nb = bk_num_buckets(single ? p : PCAP, range);
        const uint64_t mult = ((uint64_t)nb << 32) / range;
        for (int t = threadIdx.x; t < (int)nb; t += THREADS) {
            s_hist[t] = 0u;
            s_lo[t] = (uint32_t)((((uint64_t)t << 32) + mult - 1) / mult);  // smallest d with bucket(d) == t
        }
        if (threadIdx.x == 0) {
            s_count = 0u;
            s_ovf = 0u;
        }
        __syncthreads();

        // ---- 1. expansion ---------------------------------------------------------------------
        int run = 0;  // arrival index of the first product of the current round
        for (int64_t pb = a_begin; pb < a_end; pb += THREADS) {
            const int64_t pa = pb + threadIdx.x;
            int len = 0;
            int64_t bs = 0;
            double av = 0.0;
            if (pa < a_end) {
                const int32_t k = ldg_i32(a.col + pa);
                av = ldg_f64(a.val + pa);
                bs = ldg_i64(b.ptr + k);
                len = (int)(ldg_i64(b.ptr + k + 1) - bs);
            }
            int wtotal;
            const int off = warp_excl_scan(len, lane, wtotal);
            if (lane == 0) s_wtot[warp] = wtotal;
            __syncthreads();
            int base = run, all = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int t = s_wtot[w];
                if (w < warp) base += t;
                all += t;
            }
            for (int tb = 0; tb < wtotal; tb += 32 * UN) {
                int64_t q[UN];
                double aj[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    q[u] = 0;
                    aj[u] = 0.0;
                    if (tb + u * 32 < wtotal) {  // warp-uniform
                        const int t = tb + u * 32 + lane;
                        int j = 0;
#pragma unroll
                        for (int st = 16; st > 0; st >>= 1) {
                            const int o = __shfl_sync(FULL, off, j + st);
                            if (o <= t) j += st;
                        }
                        const int oj = __shfl_sync(FULL, off, j);
                        const int64_t bsj = shfl_i64(bs, j);
                        aj[u] = shfl_f64(av, j);
                        q[u] = bsj + (t - oj);
                    }
                }
                uint32_t c[UN];
                double bv[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    c[u] = 0;
                    bv[u] = 0.0;
                    if (tb + u * 32 + lane < wtotal) {
                        c[u] = (uint32_t)ldg_i32(b.col + q[u]);
                        bv[u] = ldg_f64(b.val + q[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    if (tb + u * 32 < wtotal) {  // warp-uniform
                        const int t = tb + u * 32 + lane;
                        const bool valid = t < wtotal;
                        uint32_t pos;
                        bool keep;
                        if (single) {
                            pos = (uint32_t)(base + t);
                            keep = valid;
                        } else {
                            keep = valid && c[u] >= c0 && c[u] < c1;
                            const unsigned km = __ballot_sync(FULL, keep);
                            uint32_t wb = 0;
                            if (km) {
                                const int leader = __ffs(km) - 1;
                                if (lane == leader) wb = atomicAdd(&s_count, (uint32_t)__popc(km));
                                wb = __shfl_sync(FULL, wb, leader);
                            }
                            pos = wb + (uint32_t)__popc(km & ((1u << lane) - 1u));
                            if (keep && pos >= (uint32_t)PMAX) {
                                s_ovf = 1u;  // pass does not fit the staging area
                                keep = false;
                            }
                        }
                        if (keep) {
                            const uint32_t d = c[u] - c0;
                            s_stage[pos] = d;
                            s_val[pos] = __dmul_rn(aj[u], bv[u]);
                            atomicAdd(&s_hist[bk_bucket(d, mult)], 1u);
                        }
                    }
                }
            }
            run += all;
            __syncthreads();
        }
        const uint32_t P = single ? (uint32_t)run : (s_count < (uint32_t)PMAX ? s_count : (uint32_t)PMAX);

        // ---- 2. bucket offsets ------------------------------------------------------------------
        if (warp == 0) {
            bk_scan128(s_hist, s_off, (int)nb, lane);
            bool big = false;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const int i = lane * 4 + qd;
                if (i < (int)nb) {
                    s_cur[i] = s_off[i];
                    big |= s_hist[i] > (uint32_t)BK_BUCKET_MAX;
                }
            }
            if (big) s_ovf = 1u;
        }
        __syncthreads();
        if (s_ovf) {  // skewed columns: hand the row to the fallback kernels
            if (threadIdx.x == 0) ovf[1 + atomicAdd(&ovf[0], 1u)] = r;
            return;
        }

        // ---- 3. scatter -----------------------------------------------------------------------------
        for (uint32_t pos = threadIdx.x; pos < P; pos += THREADS) {
            const uint32_t d = s_stage[pos];
            const uint32_t bk = bk_bucket(d, mult);
            const uint32_t slot = atomicAdd(&s_cur[bk], 1u);
            s_key[slot] = ((d - s_lo[bk]) << BK_SEQ_BITS) | pos;
        }
        __syncthreads();

        // ---- 4. sort the buckets ------------------------------------------------------------------
        for (uint32_t bk = warp; bk < nb; bk += NW) {
            const int sz = (int)s_hist[bk];
            uint32_t* kb = s_key + s_off[bk];
            int cnt = 0;
            if (sz > 256) cnt = bk_sort_count<16>(kb, sz, lane);
            else if (sz > 128) cnt = bk_sort_count<8>(kb, sz, lane);
            else if (sz > 64) cnt = bk_sort_count<4>(kb, sz, lane);
            else if (sz > 32) cnt = bk_sort_count<2>(kb, sz, lane);
            else if (sz > 0) cnt = bk_sort_count<1>(kb, sz, lane);
            if (lane == 0) s_cnt[bk] = (uint32_t)cnt;
        }
        __syncthreads();

        // ---- 5. place and store ----------------------------------------------------------------------
        if (warp == 0) bk_scan128(s_cnt, s_out, (int)nb, lane);
        __syncthreads();
        for (uint32_t bk = warp; bk < nb; bk += NW)
            bk_reduce_store(s_key + s_off[bk], s_val, (int)s_hist[bk], lane, cbase + row_total + s_out[bk], c0 + s_lo[bk],
                            c_col, c_val);
        row_total += s_out[nb];
        __syncthreads();
    }
    if (row_nnz_out && threadIdx.x == 0) row_nnz_out[r] = row_total;
}

// CTAs per SM the register budget is sized for (shared memory allows as many)
constexpr int bk_min_blocks(int pmax) { return pmax <= 1024 ? 8 : (pmax <= 2048 ? 4 : (pmax <= 4096 ? 3 : 2)); }

template <int PMAX, int THREADS>
__global__ void __launch_bounds__(THREADS, bk_min_blocks(PMAX))
k_bucket_rows(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
              const uint32_t* __restrict__ flops, const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col,
              double* __restrict__ c_val, uint32_t* __restrict__ row_nnz_out, uint32_t* __restrict__ ovf) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    bucket_row<PMAX, THREADS>(a, b, row_begin, r, flops[r], c_ptr, c_col, c_val, row_nnz_out, ovf, s_raw);
}

// ---- fallback: the rows of the overflow list (<= 4096 products) through the CTA-wide bitonic network -----
__global__ void __launch_bounds__(ESC_CTA_THREADS)
k_bitonic_numeric_list(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ flops,
                       const uint32_t* __restrict__ ovf, const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
                       uint32_t* __restrict__ row_nnz_out) {
    constexpr int N = 4096;
    typedef uint64_t K;  // any B.cols
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    double* vals = reinterpret_cast<double*>(s_raw + sizeof(K) * N);
    __shared__ CtaStage st;
    const uint32_t n_rows = ovf[0];
    for (uint32_t i = blockIdx.x; i < n_rows; i += gridDim.x) {
        const uint32_t r = ovf[1 + i];
        if (flops[r] > ESC_MAX_PRODUCTS) continue;  // heavy rows: launch_heavy_smem_list
        const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
        const int p = bitonic_cta_expand<K, N, true>(a, b, a_begin, a_end, keys, vals, st);
        for (int t = p + threadIdx.x; t < N; t += ESC_CTA_THREADS) keys[t] = KeyTraits<K>::sentinel;
        __syncthreads();
        bitonic_cta_sort<K, N>(keys);
        const int total = cta_reduce_store<K, N>(keys, vals, p, c_ptr[r], c_col, c_val, st);
        if (row_nnz_out && threadIdx.x == 0) row_nnz_out[r] = (uint32_t)total;
        __syncthreads();
    }
}

template <int PMAX, int THREADS>
static void bucket_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm, uint32_t rows,
                          const uint32_t* flops, const int64_t* c_ptr, int32_t* c_col, double* c_val,
                          uint32_t* row_nnz_out, uint32_t* ovf, cudaStream_t s) {
    const size_t smem = (sizeof(double) + 2 * sizeof(uint32_t)) * PMAX;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_bucket_rows<PMAX, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    k_bucket_rows<PMAX, THREADS><<<rows, THREADS, smem, s>>>(a, b, row_begin, perm, rows, flops, c_ptr, c_col, c_val,
                                                             row_nnz_out, ovf);
}

// bins 6..8 (<= 1024 / 2048 / 4096 products): one pass; bin 9 (<= 65536): column-range passes of <= 6144
void launch_bucket_rows(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm, uint32_t rows,
                        const uint32_t* flops, const int64_t* c_ptr, int32_t* c_col, double* c_val,
                        uint32_t* row_nnz_out, uint32_t* ovf, cudaStream_t s) {
    if (rows == 0) return;
    switch (bin) {
        case 6: bucket_launch<1024, 128>(a, b, row_begin, perm, rows, flops, c_ptr, c_col, c_val, row_nnz_out, ovf, s); break;
        case 7: bucket_launch<2048, 256>(a, b, row_begin, perm, rows, flops, c_ptr, c_col, c_val, row_nnz_out, ovf, s); break;
        case 8: bucket_launch<4096, 256>(a, b, row_begin, perm, rows, flops, c_ptr, c_col, c_val, row_nnz_out, ovf, s); break;
        default: bucket_launch<6144, 512>(a, b, row_begin, perm, rows, flops, c_ptr, c_col, c_val, row_nnz_out, ovf, s); break;
    }
}

// recomputes the overflow rows with <= 4096 products (device-side count in ovf[0])
void launch_bucket_fallback(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* flops,
                            const uint32_t* ovf, uint32_t max_rows, const int64_t* c_ptr, int32_t* c_col, double* c_val,
                            uint32_t* row_nnz_out, cudaStream_t s) {
    if (max_rows == 0) return;
    const size_t smem = (sizeof(uint64_t) + sizeof(double)) * 4096;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_bitonic_numeric_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const unsigned grid = max_rows < 296u ? max_rows : 296u;
    k_bitonic_numeric_list<<<grid, ESC_CTA_THREADS, smem, s>>>(a, b, row_begin, flops, ovf, c_ptr, c_col, c_val,
                                                               row_nnz_out);
}

}  // namespace spada
