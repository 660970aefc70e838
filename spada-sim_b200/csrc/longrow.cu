// longrow.cu -- stages 2 and 3 for LONG rows (more than 4096 intermediate products, bins >= BIN_LONG0).
//
// The reference K-tiles a row that is longer than a window into several windows, each of which produces a sorted
// PARTIAL row with its own psum address (scheduler.rs:522-524, 548-550), and merges the partial rows later: pairs
// on a PE (merge_task, scheduler.rs:381-480) or up to eight at a time on an adder tree (in_cache_merge_task,
// scheduler.rs:820-920; MergeTree::compare_pop, adder_tree.rs:145-188: smallest column first, ties to the LEFT
// leaf; Adder::add, adder_tree.rs:73-83: consecutive equal [row, col] summed left to right).  This file is that
// scheme restated for the GPU, with one difference that makes the result deterministic and bit-identical to the
// CPU oracle: partial rows are kept UNREDUCED (every product survives with its column) until the last merge, so
// the sum of one C[i,j] is the pure ascending-k, left-to-right chain  ((p1 + p2) + p3) + ...  -- a tree of
// partially reduced rows would compute (p1 + p2) + (p3 + p4), a different rounding.  No atomics anywhere.
//
//   chunk   the products of a row are numbered in arrival order (ascending k, then B's stored order); chunk c holds
//           the products [c * 4096, (c + 1) * 4096).  Cutting by product number (not by A entry) makes every chunk but
//           the last exactly 4096 long, so a B row of any length is just a sequence of chunks and every level's
//           runs have power-of-two lengths.  k_long_prefix records where every A entry's products start.
//   sort    one CTA per chunk: flattened expansion into shared memory, bitonic sort of (column << 12 | arrival)
//           keys (register chunks merged through shared memory, cta_common.cuh), (column, value) pairs written out
//           in sorted order.
//   merge   level l merges runs of 4096 * 2^(l-1) products pairwise.  One CTA per 4096 outputs: merge-path
//           partition on the two runs (binary search on a diagonal, ties X-before-Y = earlier k first, which
//           keeps equal columns in arrival order: a stable merge), both input slices staged in shared memory,
//           every thread merges 16 outputs serially, the tile is written back with coalesced stores.
//   sum     heads (first entry of every run of equal columns) are counted per 4096 outputs and scanned; every head
//           sums its run left to right and the finished row goes to its scratch row, nnz recorded.
//
// HBM traffic per product: 12 B written by the sort, 24 B per merge level, 4 + 24 B for the sums.
#include "cta_common.cuh"

namespace spada {

constexpr int LR_THREADS = ESC_CTA_THREADS;
constexpr int LR_ITEMS = LONG_UNIT / LR_THREADS;   // outputs per thread of a merge tile
static_assert(LR_ITEMS == 16, "merge tiles are 256 threads x 16 outputs");

// largest i in [0, n) with off[i] <= v; off is non-decreasing and off[0] <= v.  The whole warp probes 32 positions
// per step (a 32-ary search: 4 dependent loads for a million rows instead of 20).
template <typename T, typename V>
__device__ __forceinline__ int64_t warp_search_le(const T* __restrict__ off, int64_t n, V v, int lane) {
    int64_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const int64_t step = (hi - lo + 31) >> 5;
        const int64_t idx = lo + step * (lane + 1);
        const bool le = idx < hi && (V)off[idx] <= v;
        const int c = __popc(__ballot_sync(FULL, le));
        const int64_t nhi = lo + step * (c + 1);
        lo += step * c;
        hi = nhi < hi ? nhi : hi;
    }
    return lo;
}

// ---- per-wave tables ----------------------------------------------------------------------------------------
__global__ void k_long_setup(const uint32_t* __restrict__ rows_list, uint32_t n, const uint32_t* __restrict__ flops,
                             uint32_t* __restrict__ p, uint32_t* __restrict__ u) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t P = flops[rows_list[i]];
    p[i] = P;
    u[i] = (uint32_t)(((uint64_t)P + LONG_UNIT - 1) >> LONG_UNIT_LOG);
}

// aseq[e] = arrival number of the first product of A entry e inside its row (exclusive scan of the B-row lengths)
__global__ void __launch_bounds__(LR_THREADS)
k_long_prefix(DevCsr a, int64_t row_begin, const uint32_t* __restrict__ rows_list, const uint32_t* __restrict__ b_len,
              uint32_t* __restrict__ aseq) {
    __shared__ uint32_t s_w[LR_THREADS / 32];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t r = rows_list[blockIdx.x];
    const int64_t a0 = a.ptr[row_begin + r], a1 = a.ptr[row_begin + r + 1];
    uint32_t carry = 0;
    for (int64_t pb = a0; pb < a1; pb += LR_THREADS) {
        const int64_t e = pb + threadIdx.x;
        const uint32_t len = e < a1 ? __ldg(b_len + ldg_i32(a.col + e)) : 0u;
        uint32_t x = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        uint32_t base = 0, all = 0;
#pragma unroll
        for (int w = 0; w < LR_THREADS / 32; ++w) {
            const uint32_t t = s_w[w];
            if (w < warp) base += t;
            all += t;
        }
        if (e < a1) aseq[e] = carry + base + (x - len);
        carry += all;
        __syncthreads();
    }
}

// which row of the wave a chunk / tile belongs to, found once per CTA by warp 0
struct UnitInfo {
    uint32_t i;        // index inside the wave
    uint32_t row;      // row of A (relative to row_begin)
    uint32_t P;        // products of the row
    uint32_t t;        // chunk / tile number inside the row
    int64_t base;      // start of the row inside the ping-pong buffers
};
__device__ __forceinline__ bool find_unit(int64_t g, const int64_t* __restrict__ unit_off,
                                          const int64_t* __restrict__ prod_off, const uint32_t* __restrict__ p,
                                          const uint32_t* __restrict__ rows_list, uint32_t n, UnitInfo* s_info) {
    if (g >= unit_off[n]) return false;   // uniform over the CTA
    if (threadIdx.x < 32) {
        const uint32_t i = (uint32_t)warp_search_le<int64_t, int64_t>(unit_off, n, g, lane_id());
        if (threadIdx.x == 0) {
            s_info->i = i;
            s_info->row = rows_list[i];
            s_info->P = p[i];
            s_info->t = (uint32_t)(g - unit_off[i]);
            s_info->base = prod_off[i];
        }
    }
    __syncthreads();
    return true;
}

// ---- sort: one CTA per chunk ---------------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(LR_THREADS)
k_long_chunk_sort(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ rows_list, uint32_t n,
                  const uint32_t* __restrict__ p, const int64_t* __restrict__ unit_off,
                  const int64_t* __restrict__ prod_off, const uint32_t* __restrict__ aseq,
                  int32_t* __restrict__ out_col, double* __restrict__ out_val) {
    constexpr int N = LONG_UNIT;
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    double* vals = reinterpret_cast<double*>(s_raw + sizeof(K) * N);
    __shared__ CtaStage st;
    __shared__ UnitInfo info;
    __shared__ int64_t s_e[2];
    if (!find_unit(blockIdx.x, unit_off, prod_off, p, rows_list, n, &info)) return;
    const int lane = lane_id();
    const uint32_t s0 = info.t << LONG_UNIT_LOG;
    const uint32_t cnt = info.P - s0 < (uint32_t)N ? info.P - s0 : (uint32_t)N;
    const int64_t a0 = a.ptr[row_begin + info.row], a1 = a.ptr[row_begin + info.row + 1];
    // A entries that meet the chunk: e0 = last entry starting at or before s0 (the last of a tie is the non-empty
    // one), e1 = first entry starting at or after s0 + cnt
    if (threadIdx.x < 32) {
        const int64_t e0 = warp_search_le<uint32_t, uint32_t>(aseq + a0, a1 - a0, s0, lane);
        const int64_t e1 = warp_search_le<uint32_t, uint32_t>(aseq + a0, a1 - a0, s0 + cnt - 1u, lane) + 1;
        if (lane == 0) {
            s_e[0] = a0 + e0;
            s_e[1] = a0 + e1;
        }
    }
    __syncthreads();
    const int64_t e0 = s_e[0], e1 = s_e[1];
    for (int64_t pb = e0; pb < e1; pb += LR_THREADS) {
        const int64_t e = pb + threadIdx.x;
        int off = (int)cnt;
        int64_t bs = 0;
        double av = 0.0;
        if (e < e1) {
            const int32_t k = ldg_i32(a.col + e);
            av = ldg_f64(a.val + e);
            int len;
            b_row(b, k, bs, len);
            const uint32_t q0 = aseq[e];
            const uint32_t lo = q0 > s0 ? q0 : s0;   // the entry's first product inside the chunk
            off = (int)(lo - s0);
            bs += (int64_t)(lo - q0);
        }
        st.off[threadIdx.x] = off;
        st.bs[threadIdx.x] = bs;
        st.av[threadIdx.x] = av;
        if (threadIdx.x == 0) st.off[LR_THREADS] = (pb + LR_THREADS < e1) ? (int)(aseq[pb + LR_THREADS] - s0) : (int)cnt;
        __syncthreads();
        const int n_ent = (int)((e1 - pb) < LR_THREADS ? (e1 - pb) : LR_THREADS);
        const int first = st.off[0], end = st.off[LR_THREADS];
        for (int t0 = first + (int)threadIdx.x; t0 < end; t0 += 2 * LR_THREADS) {
            int t[2] = {t0, t0 + LR_THREADS};
            int64_t q[2];
            int j[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                int lo = 0, hi = n_ent;   // largest j with off[j] <= t
                if (t[u] < end) {
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (st.off[mid] <= t[u]) lo = mid; else hi = mid;
                    }
                }
                j[u] = lo;
                q[u] = st.bs[lo] + (t[u] - st.off[lo]);
            }
            uint32_t c[2];
            double bv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                c[u] = 0;
                bv[u] = 0.0;
                if (t[u] < end) {
                    c[u] = (uint32_t)ldg_i32(b.col + q[u]);
                    bv[u] = ldg_f64(b.val + q[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (t[u] < end) {
                    keys[t[u]] = ((K)c[u] << LONG_UNIT_LOG) | (K)t[u];
                    vals[t[u]] = __dmul_rn(st.av[j[u]], bv[u]);
                }
        }
        __syncthreads();
    }
    for (int t = (int)cnt + threadIdx.x; t < N; t += LR_THREADS) keys[t] = KeyTraits<K>::sentinel;
    __syncthreads();
    bitonic_cta_sort<K, N>(keys);
    const int64_t dst = info.base + s0;
    for (int t = threadIdx.x; t < (int)cnt; t += LR_THREADS) {
        const K key = keys[t];
        out_col[dst + t] = (int32_t)(uint32_t)(key >> LONG_UNIT_LOG);
        out_val[dst + t] = vals[(int)(key & (K)(N - 1))];
    }
}

// ---- merge: one CTA per 4096 outputs of one level --------------------------------------------------------------
// number of X elements among the first d outputs of the stable merge of X (nx) and Y (ny), ties X first
template <typename P>
__device__ __forceinline__ int merge_path(P x, int nx, P y, int ny, int d) {
    int lo = d > ny ? d - ny : 0, hi = d < nx ? d : nx;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] <= y[d - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(LR_THREADS, 3)
k_long_merge(const uint32_t* __restrict__ rows_list, uint32_t n, uint32_t i_lo, int level,
             const uint32_t* __restrict__ p, const int64_t* __restrict__ unit_off, const int64_t* __restrict__ prod_off,
             const int32_t* __restrict__ in_col, const double* __restrict__ in_val, int32_t* __restrict__ out_col,
             double* __restrict__ out_val) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    int32_t* s_col = reinterpret_cast<int32_t*>(s_raw);
    double* s_val = reinterpret_cast<double*>(s_raw + sizeof(int32_t) * LONG_UNIT);
    __shared__ UnitInfo info;
    __shared__ int s_part[2];
    if (!find_unit(unit_off[i_lo] + blockIdx.x, unit_off, prod_off, p, rows_list, n, &info)) return;
    const uint32_t P = info.P;
    if (level > long_levels(P)) return;   // the row was finished by an earlier level (cannot happen inside level bins)
    const uint64_t RL = (uint64_t)LONG_UNIT << (level - 1);
    const uint64_t o0 = (uint64_t)info.t << LONG_UNIT_LOG;
    const uint64_t o1 = o0 + LONG_UNIT < P ? o0 + LONG_UNIT : P;
    const uint64_t pbase = o0 / (2 * RL) * (2 * RL);
    const uint64_t xe = pbase + RL < P ? pbase + RL : P;
    const uint64_t ye = pbase + 2 * RL < P ? pbase + 2 * RL : P;
    const int nx = (int)(xe - pbase), ny = (int)(ye - xe);
    const int32_t* X = in_col + info.base + pbase;
    const int32_t* Y = in_col + info.base + xe;
    if (threadIdx.x == 0) s_part[0] = merge_path(X, nx, Y, ny, (int)(o0 - pbase));
    if (threadIdx.x == 32) s_part[1] = merge_path(X, nx, Y, ny, (int)(o1 - pbase));
    __syncthreads();
    const int i0 = s_part[0], i1 = s_part[1];
    const int j0 = (int)(o0 - pbase) - i0, j1 = (int)(o1 - pbase) - i1;
    const int cx = i1 - i0, cy = j1 - j0, tot = cx + cy;
    const double* Xv = in_val + info.base + pbase;
    const double* Yv = in_val + info.base + xe;
    for (int t = threadIdx.x; t < cx; t += LR_THREADS) {
        s_col[t] = X[i0 + t];
        s_val[t] = Xv[i0 + t];
    }
    for (int t = threadIdx.x; t < cy; t += LR_THREADS) {
        s_col[cx + t] = Y[j0 + t];
        s_val[cx + t] = Yv[j0 + t];
    }
    __syncthreads();
    const int d = threadIdx.x * LR_ITEMS;
    int32_t oc[LR_ITEMS];
    double ov[LR_ITEMS];
    if (d < tot) {
        int i = merge_path(s_col, cx, s_col + cx, cy, d);
        int j = d - i;
        int32_t xk = i < cx ? s_col[i] : 0x7fffffff;
        int32_t yk = j < cy ? s_col[cx + j] : 0x7fffffff;
#pragma unroll
        for (int q = 0; q < LR_ITEMS; ++q) {
            if (d + q < tot) {
                const bool tx = j >= cy || (i < cx && xk <= yk);
                const int src = tx ? i : cx + j;
                oc[q] = tx ? xk : yk;
                ov[q] = s_val[src];
                if (tx) {
                    ++i;
                    xk = i < cx ? s_col[i] : 0x7fffffff;
                } else {
                    ++j;
                    yk = j < cy ? s_col[cx + j] : 0x7fffffff;
                }
            }
        }
    }
    __syncthreads();
    if (d < tot) {
#pragma unroll
        for (int q = 0; q < LR_ITEMS; ++q)
            if (d + q < tot) {
                s_col[d + q] = oc[q];
                s_val[d + q] = ov[q];
            }
    }
    __syncthreads();
    const int64_t dst = info.base + (int64_t)o0;
    for (int t = threadIdx.x; t < tot; t += LR_THREADS) {
        out_col[dst + t] = s_col[t];
        out_val[dst + t] = s_val[t];
    }
}

// ---- sums ---------------------------------------------------------------------------------------------------
// heads (first entry of a run of equal columns) that start inside every tile of 4096 sorted products
__global__ void __launch_bounds__(LR_THREADS)
k_long_count(const uint32_t* __restrict__ rows_list, uint32_t n, const uint32_t* __restrict__ p,
             const int64_t* __restrict__ unit_off, const int64_t* __restrict__ prod_off,
             const int32_t* __restrict__ col0, const int32_t* __restrict__ col1, uint32_t* __restrict__ unit_heads) {
    __shared__ UnitInfo info;
    __shared__ int s_w[LR_THREADS / 32];
    if (!find_unit(blockIdx.x, unit_off, prod_off, p, rows_list, n, &info)) return;
    const int32_t* col = ((long_levels(info.P) & 1) ? col1 : col0) + info.base;
    const uint32_t o0 = info.t << LONG_UNIT_LOG;
    const uint32_t o1 = info.P - o0 < (uint32_t)LONG_UNIT ? info.P : o0 + LONG_UNIT;
    int cnt = 0;
    for (uint32_t pos = o0 + threadIdx.x; pos < o1; pos += LR_THREADS)
        if (pos == 0 || col[pos] != col[pos - 1]) ++cnt;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
    if (lane_id() == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < LR_THREADS / 32; ++w) t += s_w[w];
        unit_heads[blockIdx.x] = (uint32_t)t;
    }
}

// every head sums its run left to right (the oracle's order: ascending k) and writes the entry of the finished row
__global__ void __launch_bounds__(LR_THREADS)
k_long_reduce(const uint32_t* __restrict__ rows_list, uint32_t n, const uint32_t* __restrict__ p,
              const int64_t* __restrict__ unit_off, const int64_t* __restrict__ prod_off,
              const int32_t* __restrict__ col0, const int32_t* __restrict__ col1, const double* __restrict__ val0,
              const double* __restrict__ val1, const int64_t* __restrict__ unit_hoff, const int64_t* __restrict__ t_ptr,
              int32_t* __restrict__ t_col, double* __restrict__ t_val, uint32_t* __restrict__ row_nnz) {
    __shared__ UnitInfo info;
    __shared__ int s_w[LR_THREADS / 32];
    if (!find_unit(blockIdx.x, unit_off, prod_off, p, rows_list, n, &info)) return;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const bool odd = long_levels(info.P) & 1;
    const int32_t* col = (odd ? col1 : col0) + info.base;
    const double* val = (odd ? val1 : val0) + info.base;
    const uint32_t P = info.P;
    const uint32_t o0 = info.t << LONG_UNIT_LOG;
    const uint32_t o1 = P - o0 < (uint32_t)LONG_UNIT ? P : o0 + LONG_UNIT;
    const int64_t row_h0 = unit_hoff[unit_off[info.i]];
    int64_t dst = t_ptr[info.row] + (unit_hoff[blockIdx.x] - row_h0);
    if (info.t == 0 && threadIdx.x == 0) row_nnz[info.row] = (uint32_t)(unit_hoff[unit_off[info.i + 1]] - row_h0);
    for (uint32_t pb = o0; pb < o1; pb += LR_THREADS) {
        const uint32_t pos = pb + threadIdx.x;
        int32_t c = 0;
        bool head = false;
        if (pos < o1) {
            c = col[pos];
            head = pos == 0 || col[pos - 1] != c;
        }
        const unsigned hm = __ballot_sync(FULL, head);
        if (lane == 0) s_w[warp] = __popc(hm);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < LR_THREADS / 32; ++w) {
            const int t = s_w[w];
            if (w < warp) before += t;
            all += t;
        }
        if (head) {
            double sum = val[pos];
            for (uint32_t j = pos + 1; j < P && col[j] == c; ++j) sum = __dadd_rn(sum, val[j]);
            const int64_t o = dst + before + __popc(hm & ((1u << lane) - 1u));
            st_out(t_col + o, c);
            st_out(t_val + o, sum);
        }
        dst += all;
        __syncthreads();
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
void launch_long_prefix(const DevCsr& a, int64_t row_begin, const uint32_t* rows_list, uint32_t n_rows,
                        const uint32_t* b_len, uint32_t* aseq, cudaStream_t s) {
    if (n_rows) k_long_prefix<<<n_rows, LR_THREADS, 0, s>>>(a, row_begin, rows_list, b_len, aseq);
}

template <typename K>
static void chunk_sort_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* aseq, const LongWave& w,
                              cudaStream_t s) {
    const size_t smem = (sizeof(K) + sizeof(double)) * LONG_UNIT;
    static PerDeviceOnce attr;
    if (attr.first())
        cudaFuncSetAttribute(k_long_chunk_sort<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_long_chunk_sort<K><<<(unsigned)w.unit_bound, LR_THREADS, smem, s>>>(a, b, row_begin, w.rows_list, w.n_rows, w.p,
                                                                         w.unit_off, w.prod_off, aseq, w.col[0], w.val[0]);
}

uint32_t launch_long_wave(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* flops,
                          const uint32_t* aseq, const LongWave& w, const int64_t* t_ptr, int32_t* t_col, double* t_val,
                          uint32_t* row_nnz, PlanCounters* ctr, cudaStream_t s, const LongStages* stages) {
    if (w.n_rows == 0) return 0;
    uint32_t kernels = 0;
    auto on = [&](const char* what, uint32_t grid) { if (stages) stages->on(what, grid); };
    auto off = [&]() { if (stages) stages->off(); };
    on("long_sort", (uint32_t)w.unit_bound);
    k_long_setup<<<(w.n_rows + 255) / 256, 256, 0, s>>>(w.rows_list, w.n_rows, flops, w.p, w.u);
    launch_scan_u32_i64(w.p, w.n_rows, w.prod_off, w.tile_state, ctr, s);
    launch_scan_u32_i64(w.u, w.n_rows, w.unit_off, w.tile_state, ctr, s);
    // sort: 32-bit (column << 12 | arrival) keys whenever they fit
    if ((uint64_t)b.cols <= (1ull << (32 - LONG_UNIT_LOG))) chunk_sort_launch<uint32_t>(a, b, row_begin, aseq, w, s);
    else chunk_sort_launch<uint64_t>(a, b, row_begin, aseq, w, s);
    kernels += 4;
    off();
    // merge levels: rows are listed by ascending level count, level l takes the list from level_lo[l] on
    const size_t msmem = (sizeof(int32_t) + sizeof(double)) * LONG_UNIT;
    static PerDeviceOnce attr;
    if (attr.first()) cudaFuncSetAttribute(k_long_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
    on("long_merge", w.level_grid[1]);
    for (int l = 1; l <= w.max_level; ++l) {
        if (w.level_grid[l] == 0) continue;
        k_long_merge<<<w.level_grid[l], LR_THREADS, msmem, s>>>(w.rows_list, w.n_rows, w.level_lo[l], l, w.p, w.unit_off,
                                                               w.prod_off, w.col[(l - 1) & 1], w.val[(l - 1) & 1],
                                                               w.col[l & 1], w.val[l & 1]);
        ++kernels;
    }
    off();
    on("long_sums", (uint32_t)w.unit_bound);
    cudaMemsetAsync(w.unit_heads, 0, (size_t)w.unit_bound * sizeof(uint32_t), s);
    k_long_count<<<(unsigned)w.unit_bound, LR_THREADS, 0, s>>>(w.rows_list, w.n_rows, w.p, w.unit_off, w.prod_off, w.col[0],
                                                              w.col[1], w.unit_heads);
    launch_scan_u32_i64(w.unit_heads, (int64_t)w.unit_bound, w.unit_hoff, w.tile_state, ctr, s);
    k_long_reduce<<<(unsigned)w.unit_bound, LR_THREADS, 0, s>>>(w.rows_list, w.n_rows, w.p, w.unit_off, w.prod_off,
                                                               w.col[0], w.col[1], w.val[0], w.val[1], w.unit_hoff, t_ptr,
                                                               t_col, t_val, row_nnz);
    kernels += 3;
    off();
    return kernels;
}

}  // namespace spada
