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
//   sum     heads (first entry of every run of equal columns) are counted per 4096 outputs (that is the row's nnz);
//           after the row_ptr scan every head sums its run left to right straight into C -- and, in a sharded run,
//           into every peer's C: for long rows the sums ARE the placement.
//
// Buffers: a long row's scratch row (capacity = its products) holds its sorted products; the merge levels ping-pong
// between it and a pong buffer that only has to hold one WAVE of rows.  A row with L levels starts in buffer L & 1
// (0 = scratch row, 1 = pong), so that its last level lands in the scratch row.
//
// HBM traffic per product: 12 B written by the sort, 24 B per merge level, 4 B for the count, 12 + 12 B for the sums.
#include <algorithm>
#include <cstdio>

#include "cta_common.cuh"

namespace spada {

constexpr int LR_THREADS = ESC_CTA_THREADS;

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
    uint32_t i;        // index inside the long-row list
    uint32_t row;      // row of A (relative to row_begin)
    uint32_t P;        // products of the row
    uint32_t t;        // chunk / tile number inside the row
    int64_t base0;     // start of the row's scratch row (buffer 0)
    int64_t base1;     // start of the row inside the pong buffer of its wave (buffer 1)
};
// the tables of the whole long-row list + the two buffers (see the header comment)
struct LongDev {
    const uint32_t* rows_list;
    const uint32_t* p;
    const uint32_t* unit_row;
    const int64_t* unit_off;
    const int64_t* prod_off;
    const int64_t* t_ptr;
    int32_t* col[2];
    double* val[2];
    uint32_t wave_lo;   // first list index of the wave in flight (origin of the pong buffer)
};
__device__ __forceinline__ void find_unit(int64_t g, const LongDev& D, UnitInfo* s_info) {
    if (threadIdx.x == 0) {
        const uint32_t i = D.unit_row[g];
        s_info->i = i;
        s_info->row = D.rows_list[i];
        s_info->P = D.p[i];
        s_info->t = (uint32_t)(g - D.unit_off[i]);
        s_info->base0 = D.t_ptr[s_info->row];
        s_info->base1 = D.prod_off[i] - D.prod_off[D.wave_lo];
    }
    __syncthreads();
}

// unit_row[g] = row (index inside the list) of chunk / tile g: one binary search per unit
__global__ void k_long_units(const int64_t* __restrict__ unit_off, uint32_t n, uint32_t* __restrict__ unit_row) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= unit_off[n]) return;
    uint32_t lo = 0, hi = n;   // unit_off[lo] <= g < unit_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (unit_off[mid] <= g) lo = mid; else hi = mid;
    }
    unit_row[g] = lo;
}

// ---- sort: one CTA per chunk ---------------------------------------------------------------------------------
// N = sort capacity of this chunk: 4096, or 2048 / 1024 for a row's last chunk when it is that short (a row of 5000
// products is one full chunk and one of 904: sorting the second one at full width would waste a quarter of the work)
template <typename K, bool SPLIT, int N>
__device__ __forceinline__ void chunk_body(const DevCsr a, const DevCsr b, const uint32_t* __restrict__ aseq,
                                           uint32_t s0, uint32_t cnt, int64_t e0, int64_t e1, K* keys, double* vals,
                                           CtaStage& st, uint32_t* top, int32_t* __restrict__ out_col,
                                           double* __restrict__ out_val, int64_t dst) {
    constexpr int SB = Log2<N>::v;
    const int lane = lane_id();
    if constexpr (SPLIT)
        for (int t = threadIdx.x; t < N / 32 + 1; t += LR_THREADS) top[t] = 0u;   // the loop below syncs before it writes
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
        for (int t0 = first + (int)threadIdx.x; t0 < end; t0 += CTA_EXPAND_UNROLL * LR_THREADS) {
            int t[CTA_EXPAND_UNROLL];
            int64_t q[CTA_EXPAND_UNROLL];
            int j[CTA_EXPAND_UNROLL];
#pragma unroll
            for (int u = 0; u < CTA_EXPAND_UNROLL; ++u) t[u] = t0 + u * LR_THREADS;
#pragma unroll
            for (int u = 0; u < CTA_EXPAND_UNROLL; ++u) {
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
            uint32_t c[CTA_EXPAND_UNROLL];
            double bv[CTA_EXPAND_UNROLL];
#pragma unroll
            for (int u = 0; u < CTA_EXPAND_UNROLL; ++u) {
                c[u] = 0;
                bv[u] = 0.0;
                if (t[u] < end) {
                    c[u] = (uint32_t)ldg_i32(b.col + q[u]);
                    bv[u] = ldg_f64(b.val + q[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < CTA_EXPAND_UNROLL; ++u)
                if (t[u] < end) {
                    keys[KeySlot<K, N>::at(t[u])] = ((K)c[u] << SB) | (K)t[u];   // SPLIT: the top bit falls off
                    vals[t[u]] = __dmul_rn(st.av[j[u]], bv[u]);
                    if constexpr (SPLIT) top_bit_mark(top, t[u], (c[u] >> (32 - SB)) & 1u, lane);
                }
        }
        __syncthreads();
    }
    for (int t = (int)cnt + threadIdx.x; t < N; t += LR_THREADS) keys[KeySlot<K, N>::at(t)] = KeyTraits<K>::sentinel;
    __syncthreads();
    bitonic_cta_sort<K, N>(keys);
    int n0 = 0x7fffffff;
    if constexpr (SPLIT) n0 = cta_split_top<N>(keys, (int)cnt, top, st);
    constexpr uint32_t TOP = sizeof(K) == 4 ? 1u << (32 - SB) : 0u;
    for (int t = threadIdx.x; t < (int)cnt; t += LR_THREADS) {
        const K key = keys[KeySlot<K, N>::at(t)];
        out_col[dst + t] = (int32_t)((uint32_t)(key >> SB) | (t >= n0 ? TOP : 0u));
        out_val[dst + t] = vals[(int)(key & (K)(N - 1))];
    }
}

// SPLIT: columns up to 2^21 with 32-bit keys -- sorted without their top bit, then split by it (cta_split_top; the
// 64-bit network measured 3.2x slower per product)
template <typename K, bool SPLIT>
__global__ void __launch_bounds__(LR_THREADS)
k_long_chunk_sort(DevCsr a, DevCsr b, int64_t row_begin, LongDev D, uint32_t i_hi, const uint32_t* __restrict__ aseq) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    double* vals = reinterpret_cast<double*>(s_raw + sizeof(K) * LONG_UNIT);
    __shared__ CtaStage st;
    __shared__ UnitInfo info;
    __shared__ int64_t s_e[2];
    __shared__ uint32_t top[SPLIT ? LONG_UNIT / 32 + 1 : 1];
    const int64_t g = D.unit_off[D.wave_lo] + blockIdx.x;
    if (g >= D.unit_off[i_hi]) return;   // uniform over the CTA
    find_unit(g, D, &info);
    const int lane = lane_id();
    const int buf = long_levels(info.P) & 1;   // where the row's chunks go so that its last level ends in the scratch row
    const uint32_t s0 = info.t << LONG_UNIT_LOG;
    const uint32_t cnt = info.P - s0 < (uint32_t)LONG_UNIT ? info.P - s0 : (uint32_t)LONG_UNIT;
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
    const int64_t dst = (buf ? info.base1 : info.base0) + s0;
    if (cnt <= 1024u)
        chunk_body<K, SPLIT, 1024>(a, b, aseq, s0, cnt, s_e[0], s_e[1], keys, vals, st, top, buf ? D.col[1] : D.col[0], buf ? D.val[1] : D.val[0], dst);
    else if (cnt <= 2048u)
        chunk_body<K, SPLIT, 2048>(a, b, aseq, s0, cnt, s_e[0], s_e[1], keys, vals, st, top, buf ? D.col[1] : D.col[0], buf ? D.val[1] : D.val[0], dst);
    else
        chunk_body<K, SPLIT, LONG_UNIT>(a, b, aseq, s0, cnt, s_e[0], s_e[1], keys, vals, st, top, buf ? D.col[1] : D.col[0], buf ? D.val[1] : D.val[0], dst);
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

// Every level runs two kernels.  k_long_partition: one thread per output tile finds the tile's input slices with two
// merge-path searches (millions of independent searches hide their latency) and leaves a 32-byte descriptor.
// k_long_merge: persistent CTAs with two shared-memory stages; thread 0 reads a descriptor and starts the four TMA
// bulk copies (cp.async.bulk, SASS UBLKCP) of a tile two tiles ahead, so the input slices of tile k+1 are in flight
// while tile k is merged; completion is counted by the stage's mbarrier.  Bulk copies need 16-byte aligned addresses
// and sizes: every slice is fetched from the aligned address below its start to the aligned address above its end,
// the merge indexes past the few extra elements.
constexpr int RD_PAD = 4;   // spare elements of k_long_reduce's stage for the bulk copies' alignment
constexpr int MG_THREADS = 512;
constexpr int MG_ITEMS = LONG_UNIT / MG_THREADS;   // outputs per thread
constexpr int MG_PAD = 16;
// The merged tile goes back through the stage for coalesced stores.  Every thread writes MG_ITEMS consecutive outputs:
// one spare word per 32 column ids and one spare double per 16 values spread the lanes of a warp over all banks
// (unpadded: 8-way conflicts on the columns, 16-way on the values).
__device__ __forceinline__ int mg_col_at(int e) { return e + (e >> 5); }
__device__ __forceinline__ int mg_val_at(int e) { return e + (e >> 4); }
struct MergeStage {
    int32_t col[LONG_UNIT + LONG_UNIT / 32 + MG_PAD];
    double val[LONG_UNIT + LONG_UNIT / 16 + MG_PAD];
};
struct __align__(16) MergeTile {
    int64_t ax, ay;       // element index (inside the input buffer) of the first X / Y element the tile consumes
    int64_t out;          // where the tile's outputs go (inside the other buffer)
    uint32_t cnt;         // elements taken from X (bits 0-12) and from Y (13-25); bit 26: the input is buffer 1;
                          // bit 27: the row's last level -- the tile also counts the run heads that start inside it
    int32_t prev;         // last level only: column of the output element before the tile, -1: the tile opens the row
};
constexpr uint32_t MT_BUF1 = 1u << 26, MT_LAST = 1u << 27;
__device__ __forceinline__ int mt_cx(const MergeTile& T) { return (int)(T.cnt & 0x1fffu); }
__device__ __forceinline__ int mt_cy(const MergeTile& T) { return (int)((T.cnt >> 13) & 0x1fffu); }

__global__ void __launch_bounds__(256)
k_long_partition(LongDev D, uint32_t i_lo, uint32_t i_hi, int level, MergeTile* __restrict__ tiles) {
    const int64_t g0 = D.unit_off[i_lo];
    const int64_t g = g0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D.unit_off[i_hi]) return;
    const uint32_t i = D.unit_row[g];
    const uint32_t P = D.p[i];
    const int L = long_levels(P);
    MergeTile T{};
    if (level <= L) {   // rows finished by an earlier level take no part (cannot happen inside level bins)
        const int bin = (L - level + 1) & 1;   // buffer the level reads; it writes the other one
        const int64_t b0 = D.t_ptr[D.rows_list[i]], b1 = D.prod_off[i] - D.prod_off[D.wave_lo];
        const int64_t base_in = bin ? b1 : b0, base_out = bin ? b0 : b1;
        const int32_t* in_col = bin ? D.col[1] : D.col[0];
        const uint64_t RL = (uint64_t)LONG_UNIT << (level - 1);
        const uint64_t o0 = (uint64_t)(g - D.unit_off[i]) << LONG_UNIT_LOG;
        const uint64_t o1 = o0 + LONG_UNIT < P ? o0 + LONG_UNIT : P;
        const uint64_t pbase = o0 / (2 * RL) * (2 * RL);
        const uint64_t xe = pbase + RL < P ? pbase + RL : P;
        const uint64_t ye = pbase + 2 * RL < P ? pbase + 2 * RL : P;
        const int nx = (int)(xe - pbase), ny = (int)(ye - xe);
        const int32_t* X = in_col + base_in + pbase;
        const int32_t* Y = in_col + base_in + xe;
        const int d0 = (int)(o0 - pbase), d1 = (int)(o1 - pbase);
        const int i0 = merge_path(X, nx, Y, ny, d0);
        const int i1 = (d1 == nx + ny) ? nx : merge_path(X, nx, Y, ny, d1);
        T.ax = base_in + (int64_t)pbase + i0;
        T.ay = base_in + (int64_t)xe + (d0 - i0);
        T.out = base_out + (int64_t)o0;
        T.cnt = (uint32_t)(i1 - i0) | (uint32_t)((d1 - i1) - (d0 - i0)) << 13 | (bin ? MT_BUF1 : 0u);
        if (level == L) {   // the output of this level is the row in its final order
            int32_t prev = -1;
            if (i0 > 0) prev = X[i0 - 1];
            if (d0 - i0 > 0) prev = max(prev, Y[d0 - i0 - 1]);   // the later of the two in merge order has the larger column
            T.prev = prev;
            T.cnt |= MT_LAST;
        }
    }
    tiles[g - g0] = T;
}

__global__ void __launch_bounds__(MG_THREADS, 2)
k_long_merge(const MergeTile* __restrict__ tiles, LongDev D, uint32_t i_lo, uint32_t i_hi,
             uint32_t* __restrict__ unit_heads, uint32_t* __restrict__ row_nnz) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    MergeStage* stage = reinterpret_cast<MergeStage*>(s_raw);
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int s_heads;
    const int64_t n_tiles = D.unit_off[i_hi] - D.unit_off[i_lo];
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        s_heads = 0;
    }
    __syncthreads();

    // where the X and Y slices of a tile sit inside a stage (both start on 16-byte boundaries of the source)
    auto layout = [](const MergeTile& T, int& xc, int& yc, int& xv, int& yv, int& nxc, int& nyc, int& nxv, int& nyv) {
        const int cx = mt_cx(T), cy = mt_cy(T);
        const int dxc = (int)(T.ax & 3), dyc = (int)(T.ay & 3);
        nxc = cx ? (dxc + cx + 3) & ~3 : 0;
        nyc = cy ? (dyc + cy + 3) & ~3 : 0;
        xc = dxc;
        yc = nxc + dyc;
        const int dxv = (int)(T.ax & 1), dyv = (int)(T.ay & 1);
        nxv = cx ? (dxv + cx + 1) & ~1 : 0;
        nyv = cy ? (dyv + cy + 1) & ~1 : 0;
        xv = dxv;
        yv = nxv + dyv;
    };
    auto issue = [&](const MergeTile& T, int sidx) {   // thread 0 only
        const int cx = mt_cx(T), cy = mt_cy(T);
        if (cx + cy == 0) return;
        const bool bin = T.cnt & MT_BUF1;
        const int32_t* in_col = bin ? D.col[1] : D.col[0];
        const double* in_val = bin ? D.val[1] : D.val[0];
        int xc, yc, xv, yv, nxc, nyc, nxv, nyv;
        layout(T, xc, yc, xv, yv, nxc, nyc, nxv, nyv);
        MergeStage& st = stage[sidx];
        mbar_expect_tx(&bar[sidx], (uint32_t)(nxc + nyc) * 4u + (uint32_t)(nxv + nyv) * 8u);
        if (cx) {
            tma_load_1d(st.col, in_col + (T.ax - xc), (uint32_t)nxc * 4u, &bar[sidx]);
            tma_load_1d(st.val, in_val + (T.ax - xv), (uint32_t)nxv * 8u, &bar[sidx]);
        }
        if (cy) {
            tma_load_1d(st.col + nxc, in_col + (T.ay - (yc - nxc)), (uint32_t)nyc * 4u, &bar[sidx]);
            tma_load_1d(st.val + nxv, in_val + (T.ay - (yv - nxv)), (uint32_t)nyv * 8u, &bar[sidx]);
        }
    };

    const int64_t G = gridDim.x;
    int64_t t = blockIdx.x;
    if (threadIdx.x == 0) {
        if (t < n_tiles) issue(tiles[t], 0);
        if (t + G < n_tiles) issue(tiles[t + G], 1);
    }
    uint32_t phase[2] = {0u, 0u};
    for (int k = 0; t < n_tiles; ++k, t += G) {
        const int cur = k & 1;
        const MergeTile T = tiles[t];
        const bool bin = T.cnt & MT_BUF1, last = T.cnt & MT_LAST;
        int32_t* __restrict__ out_col = bin ? D.col[0] : D.col[1];
        double* __restrict__ out_val = bin ? D.val[0] : D.val[1];
        MergeTile ahead{};   // descriptor of the tile two ahead: loaded now, used when this stage is free again
        const bool has_ahead = threadIdx.x == 0 && t + 2 * G < n_tiles;
        if (has_ahead) ahead = tiles[t + 2 * G];
        const int cx = mt_cx(T), cy = mt_cy(T), tot = cx + cy;
        if (tot == 0) {                       // uniform; nothing was issued for this tile
            if (has_ahead) issue(ahead, cur);
            continue;
        }
        int ixc, iyc, ixv, iyv, nxc, nyc, nxv, nyv;
        layout(T, ixc, iyc, ixv, iyv, nxc, nyc, nxv, nyv);
        mbar_wait(&bar[cur], phase[cur]);
        phase[cur] ^= 1u;
        MergeStage& st = stage[cur];
        const int32_t* xc = st.col + ixc;
        const int32_t* yc = st.col + iyc;
        const double* xv = st.val + ixv;
        const double* yv = st.val + iyv;
        const int d = threadIdx.x * MG_ITEMS;
        int32_t oc[MG_ITEMS];
        double ov[MG_ITEMS];
        if (d < tot) {
            int i = merge_path(xc, cx, yc, cy, d);
            int j = d - i;
            int32_t xk = i < cx ? xc[i] : 0x7fffffff;
            int32_t yk = j < cy ? yc[j] : 0x7fffffff;
#pragma unroll
            for (int q = 0; q < MG_ITEMS; ++q) {
                if (d + q < tot) {
                    const bool tx = j >= cy || (i < cx && xk <= yk);
                    oc[q] = tx ? xk : yk;
                    ov[q] = tx ? xv[i] : yv[j];
                    if (tx) {
                        ++i;
                        xk = i < cx ? xc[i] : 0x7fffffff;
                    } else {
                        ++j;
                        yk = j < cy ? yc[j] : 0x7fffffff;
                    }
                }
            }
        }
        __syncthreads();
        if (d < tot) {
#pragma unroll
            for (int q = 0; q < MG_ITEMS; ++q)
                if (d + q < tot) {
                    st.col[mg_col_at(d + q)] = oc[q];
                    st.val[mg_val_at(d + q)] = ov[q];
                }
        }
        __syncthreads();
        int heads = 0;   // last level: run heads (first product of a column) among this thread's outputs
        for (int e = threadIdx.x; e < tot; e += MG_THREADS) {
            const int32_t c = st.col[mg_col_at(e)];
            out_col[T.out + e] = c;
            out_val[T.out + e] = st.val[mg_val_at(e)];
            if (last) heads += c != (e ? st.col[mg_col_at(e - 1)] : T.prev);
        }
        if (last) {   // uniform
            heads = (int)__reduce_add_sync(FULL, (unsigned)heads);
            if (heads && (threadIdx.x & 31) == 0) atomicAdd(&s_heads, heads);
        }
        // the stage goes back to the async proxy: the bulk copies of the tile two ahead write it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (has_ahead) issue(ahead, cur);
        if (last && threadIdx.x == 32) {
            // the tile is unit g of the wave: its head count places the row's sums later; the row's nnz is their total
            // (integer atomics: deterministic).  s_heads is clear again before the next tile's barrier lets anybody add.
            const int64_t g = D.unit_off[i_lo] + t;
            const uint32_t h = (uint32_t)s_heads;
            s_heads = 0;
            unit_heads[g] = h;
            atomicAdd(row_nnz + D.rows_list[D.unit_row[g]], h);
        }
    }
}

// ---- sums ---------------------------------------------------------------------------------------------------
// Every head sums its run left to right (the oracle's order: ascending k) and writes the entry of the finished row.
// The tile's 4096 sorted products are staged in shared memory first (coalesced loads, all in flight at once), so
// the dependent chain of a run (compare the next column, add the next value) runs at shared-memory latency, and all
// heads of the tile run side by side: warp w owns the positions [512 w, 512 (w + 1)), lanes interleaved.  A run that
// leaves the tile (at most one: the tile's last) is streamed by the whole CTA, 4096 products at a time, with one
// thread doing the adds in order -- the diagonal of A x A^T is one run of nnz(A row) products.
// ND = 1: one destination (this GPU's C); otherwise every destination of dst_all (sharded runs: the peers' C too)
template <int ND>
__device__ __forceinline__ void reduce_store_col(const CopyDst& dst_all, int64_t o, int32_t c) {
    if (ND == 1) {
        st_out(dst_all.col[0] + o, c);
    } else {
#pragma unroll 1
        for (int d = dst_all.n - 1; d >= 0; --d) st_out(dst_all.col[d] + o, c);
    }
}
template <int ND>
__device__ __forceinline__ void reduce_store_val(const CopyDst& dst_all, int64_t o, double v) {
    if (ND == 1) {
        st_out(dst_all.val[0] + o, v);
    } else {
#pragma unroll 1
        for (int d = dst_all.n - 1; d >= 0; --d) st_out(dst_all.val[d] + o, v);
    }
}

template <int ND>
__global__ void __launch_bounds__(LR_THREADS)
k_long_reduce(LongDev D, uint32_t n, const int64_t* __restrict__ unit_hoff, const int64_t* __restrict__ c_ptr, CopyDst dst_all) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    int32_t* s_col = reinterpret_cast<int32_t*>(s_raw);
    double* s_val = reinterpret_cast<double*>(s_raw + sizeof(int32_t) * (LONG_UNIT + RD_PAD));
    __shared__ __align__(8) uint64_t bar;
    __shared__ UnitInfo info;
    __shared__ int s_w[LR_THREADS / 32];
    __shared__ int s_prev;            // column of the product before the tile
    __shared__ int s_open_rank;       // output slot of the run that leaves the tile, -1: none
    __shared__ double s_open_sum;
    __shared__ int s_end;
    // rows are listed by ascending product count: the tiles are taken from the far end, so that the longest runs
    // (strictly sequential sums, e.g. the diagonal of A x A^T) start first and the short tiles fill in beside them
    const int64_t n_units = D.unit_off[n];
    const int64_t unit = n_units - 1 - (int64_t)blockIdx.x;
    if (unit < 0) return;
    find_unit(unit, D, &info);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int32_t* __restrict__ col = D.col[0] + info.base0;
    const double* __restrict__ val = D.val[0] + info.base0;
    const uint32_t P = info.P;
    const uint32_t o0 = info.t << LONG_UNIT_LOG;
    const int cnt = (int)(P - o0 < (uint32_t)LONG_UNIT ? P - o0 : (uint32_t)LONG_UNIT);
    const int64_t row_h0 = unit_hoff[D.unit_off[info.i]];
    const int64_t dst = c_ptr[info.row] + (unit_hoff[unit] - row_h0) +
                        shard_offset(dst_all.off, dst_all.shard_nnz, dst_all.shard_idx);
    // the whole tile in flight at once: two bulk copies (from the 16-byte boundary below the tile's first product to
    // the one above its last; sc / sv index past the few extra elements) instead of 32 loads per thread in batches
    const int dxc = (int)((reinterpret_cast<uintptr_t>(col + o0) >> 2) & 3), dxv = (int)((reinterpret_cast<uintptr_t>(val + o0) >> 3) & 1);
    const int32_t* sc = s_col + dxc;
    const double* sv = s_val + dxv;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        const uint32_t bc = (uint32_t)((dxc + cnt + 3) & ~3) * 4u, bv = (uint32_t)((dxv + cnt + 1) & ~1) * 8u;
        mbar_expect_tx(&bar, bc + bv);
        tma_load_1d(s_col, col + o0 - dxc, bc, &bar);
        tma_load_1d(s_val, val + o0 - dxv, bv, &bar);
        s_prev = o0 ? col[o0 - 1] : -1;
        s_open_rank = -1;
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    constexpr int SEG = LONG_UNIT / (LR_THREADS / 32);   // positions per warp
    constexpr int ITERS = SEG / 32;
    unsigned hm[ITERS];
    int mine = 0;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int t = warp * SEG + it * 32 + lane;
        bool head = false;
        if (t < cnt) {
            const int32_t c = sc[t];
            head = (t == 0) ? (s_prev != c) : (sc[t - 1] != c);
        }
        hm[it] = __ballot_sync(FULL, head);
        mine += __popc(hm[it]);
    }
    if (lane == 0) s_w[warp] = mine;
    __syncthreads();
    int rank = 0;
#pragma unroll
    for (int w = 0; w < LR_THREADS / 32; ++w)
        if (w < warp) rank += s_w[w];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        if ((hm[it] >> lane) & 1u) {
            const int t = warp * SEG + it * 32 + lane;
            const int32_t c = sc[t];
            double sum = sv[t];
            int j = t + 1;
            while (j < cnt && sc[j] == c) {
                sum = __dadd_rn(sum, sv[j]);
                ++j;
            }
            const int o = rank + __popc(hm[it] & ((1u << lane) - 1u));
            reduce_store_col<ND>(dst_all, dst + o, c);
            if (j == cnt && o0 + (uint32_t)cnt < P) {   // the run may go on in the next tile: finished below
                s_open_rank = o;
                s_open_sum = sum;
            } else {
                reduce_store_val<ND>(dst_all, dst + o, sum);
            }
        }
        rank += __popc(hm[it]);
    }
    __syncthreads();
    if (s_open_rank < 0) return;   // uniform
    const int32_t c = sc[cnt - 1];
    double sum = s_open_sum;
    for (uint32_t q0 = o0 + (uint32_t)cnt; q0 < P; q0 += LONG_UNIT) {
        const int len = (int)(P - q0 < (uint32_t)LONG_UNIT ? P - q0 : (uint32_t)LONG_UNIT);
        __syncthreads();
        if (threadIdx.x == 0) s_end = len;
        for (int t = threadIdx.x; t < len; t += LR_THREADS) {
            s_col[t] = col[q0 + t];
            s_val[t] = val[q0 + t];
        }
        __syncthreads();
        // first position whose column differs: the sorted order makes "differs" monotone, so probing the chunk's
        // positions with one atomicMin per warp that sees the change is enough
        for (int tb = 0; tb < len; tb += LR_THREADS) {
            const int t = tb + threadIdx.x;
            const bool diff = t < len && s_col[t] != c;
            const unsigned dm = __ballot_sync(FULL, diff);
            if (diff && (dm & ((1u << lane) - 1u)) == 0) atomicMin(&s_end, t);
        }
        __syncthreads();
        const int e = s_end;
        if (threadIdx.x == 0) {
            int j = 0;
            for (; j + 8 <= e; j += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = s_val[j + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) sum = __dadd_rn(sum, v[u]);
            }
            for (; j < e; ++j) sum = __dadd_rn(sum, s_val[j]);
        }
        if (e < len) break;   // uniform: s_end is shared
    }
    if (threadIdx.x == 0) reduce_store_val<ND>(dst_all, dst + s_open_rank, sum);
}

// ---- narrow outputs: long rows of a product with few columns ------------------------------------------------
// When B has few columns (cari: 400) a long row lands on few outputs, every one hit hundreds of times; carrying all
// those products through merge levels is wasted traffic.  One CTA per row keeps a dense accumulator of B.cols values
// in shared memory and walks the A row in stored order: all threads take the elements of ONE B row (distinct columns,
// no conflicts), then a barrier, then the next B row -- every column is updated in ascending k, the oracle's order,
// without atomics.  The first product of a column is stored as it is (not added to +0.0: -0.0 survives).  The
// elements of the next B row are fetched into registers while the current one is applied.
constexpr int DN_THREADS = 128;
constexpr int DN_BATCH = 256;     // A entries staged per batch

__global__ void __launch_bounds__(DN_THREADS)
k_dense_rows(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ rows_list, const int64_t* __restrict__ t_ptr,
             int32_t* __restrict__ t_col, double* __restrict__ t_val, uint32_t* __restrict__ row_nnz) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int ncols = (int)b.cols;
    double* acc = reinterpret_cast<double*>(s_raw);
    unsigned char* hit = s_raw + sizeof(double) * (size_t)ncols;
    __shared__ int64_t s_bs[DN_BATCH];
    __shared__ int s_len[DN_BATCH];
    __shared__ double s_av[DN_BATCH];
    __shared__ int s_w[DN_THREADS / 32];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t r = rows_list[blockIdx.x];
    const int64_t a0 = a.ptr[row_begin + r], a1 = a.ptr[row_begin + r + 1];
    for (int c = threadIdx.x; c < ncols; c += DN_THREADS) hit[c] = 0;
    for (int64_t pb = a0; pb < a1; pb += DN_BATCH) {
        __syncthreads();
        for (int t = threadIdx.x; t < DN_BATCH; t += DN_THREADS) {
            const int64_t e = pb + t;
            int64_t bs = 0;
            int len = 0;
            double av = 0.0;
            if (e < a1) {
                av = ldg_f64(a.val + e);
                b_row(b, ldg_i32(a.col + e), bs, len);
            }
            s_bs[t] = bs;
            s_len[t] = len;
            s_av[t] = av;
        }
        __syncthreads();
        const int n_ent = (int)((a1 - pb) < DN_BATCH ? (a1 - pb) : DN_BATCH);
        // software pipeline: the first DN_THREADS elements of entry i + 1 are loaded while entry i is applied
        int32_t nc = 0;
        double nv = 0.0;
        if ((int)threadIdx.x < s_len[0]) {
            nc = ldg_i32(b.col + s_bs[0] + threadIdx.x);
            nv = ldg_f64(b.val + s_bs[0] + threadIdx.x);
        }
        for (int i = 0; i < n_ent; ++i) {
            const int32_t c0 = nc;
            const double v0 = nv;
            if (i + 1 < n_ent && (int)threadIdx.x < s_len[i + 1]) {
                nc = ldg_i32(b.col + s_bs[i + 1] + threadIdx.x);
                nv = ldg_f64(b.val + s_bs[i + 1] + threadIdx.x);
            }
            const int len = s_len[i];
            const double av = s_av[i];
            if ((int)threadIdx.x < len) {
                const double prod = __dmul_rn(av, v0);
                acc[c0] = hit[c0] ? __dadd_rn(acc[c0], prod) : prod;
                hit[c0] = 1;
            }
            for (int t = threadIdx.x + DN_THREADS; t < len; t += DN_THREADS) {   // B rows longer than the CTA
                const int32_t c = ldg_i32(b.col + s_bs[i] + t);
                const double prod = __dmul_rn(av, ldg_f64(b.val + s_bs[i] + t));
                acc[c] = hit[c] ? __dadd_rn(acc[c], prod) : prod;
                hit[c] = 1;
            }
            __syncthreads();   // the next B row may hit the same columns: strictly after this one
        }
    }
    __syncthreads();
    // compaction in ascending column order: every thread owns a contiguous range of columns
    const int per = (ncols + DN_THREADS - 1) / DN_THREADS;
    const int cbeg = threadIdx.x * per, cend = cbeg + per < ncols ? cbeg + per : ncols;
    int cnt = 0;
    for (int c = cbeg; c < cend; ++c) cnt += hit[c];
    int x = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    int base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < DN_THREADS / 32; ++w) {
        if (w < warp) base += s_w[w];
        all += s_w[w];
    }
    int64_t o = t_ptr[r] + base + (x - cnt);
    for (int c = cbeg; c < cend; ++c)
        if (hit[c]) {
            st_out(t_col + o, (int32_t)c);
            st_out(t_val + o, acc[c]);
            ++o;
        }
    if (threadIdx.x == 0) row_nnz[r] = (uint32_t)all;
}

bool dense_rows_fit(int64_t b_cols) { return b_cols > 0 && b_cols <= DENSE_MAX_COLS; }

void launch_dense_rows(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list, uint32_t n_rows,
                       const int64_t* t_ptr, int32_t* t_col, double* t_val, uint32_t* row_nnz, cudaStream_t s) {
    if (!n_rows) return;
    const size_t smem = (size_t)b.cols * (sizeof(double) + 1) + 16;
    static PerDeviceOnce attr;
    if (attr.first())
        cudaFuncSetAttribute(k_dense_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DENSE_MAX_COLS * 9 + 16));
    k_dense_rows<<<n_rows, DN_THREADS, smem, s>>>(a, b, row_begin, rows_list, t_ptr, t_col, t_val, row_nnz);
}

// ---- host side ------------------------------------------------------------------------------------------------
void launch_long_prefix(const DevCsr& a, int64_t row_begin, const uint32_t* rows_list, uint32_t n_rows,
                        const uint32_t* b_len, uint32_t* aseq, cudaStream_t s) {
    if (n_rows) k_long_prefix<<<n_rows, LR_THREADS, 0, s>>>(a, row_begin, rows_list, b_len, aseq);
}

static LongDev long_dev(const LongPlan& P, uint32_t wave_lo) {
    LongDev D{};
    D.rows_list = P.rows_list;
    D.p = P.p;
    D.unit_row = P.unit_row;
    D.unit_off = P.unit_off;
    D.prod_off = P.prod_off;
    D.t_ptr = P.t_ptr;
    D.col[0] = P.s_col;
    D.val[0] = P.s_val;
    D.col[1] = P.pong_col;
    D.val[1] = P.pong_val;
    D.wave_lo = wave_lo;
    return D;
}

// tables of the whole long-row list: products and chunks per row, their prefix sums, chunk -> row
uint32_t launch_long_setup(const LongPlan& P, const uint32_t* flops, PlanCounters* ctr, cudaStream_t s) {
    if (P.n_rows == 0) return 0;
    k_long_setup<<<(P.n_rows + 255) / 256, 256, 0, s>>>(P.rows_list, P.n_rows, flops, P.p, P.u);
    launch_scan_u32_i64(P.p, P.n_rows, P.prod_off, P.tile_state, ctr, s);
    launch_scan_u32_i64(P.u, P.n_rows, P.unit_off, P.tile_state, ctr, s);
    k_long_units<<<(unsigned)((P.unit_bound + 255) / 256), 256, 0, s>>>(P.unit_off, P.n_rows, P.unit_row);
    cudaMemsetAsync(P.unit_heads, 0, (size_t)P.unit_bound * sizeof(uint32_t), s);
    return 5;
}

template <typename K, bool SPLIT>
static void chunk_sort_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* aseq, const LongDev& D,
                              const LongWaveRange& w, cudaStream_t s) {
    const size_t smem = (sizeof(K) + sizeof(double)) * LONG_UNIT;
    static PerDeviceOnce attr;
    if (attr.first())
        cudaFuncSetAttribute(k_long_chunk_sort<K, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_long_chunk_sort<K, SPLIT><<<(unsigned)w.unit_bound, LR_THREADS, smem, s>>>(a, b, row_begin, D, w.hi, aseq);
}

// one wave: chunk sorts, merge levels (the rows of the wave end in their scratch rows), head counts -> row_nnz
uint32_t launch_long_wave(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* aseq, const LongPlan& P,
                          const LongWaveRange& w, uint32_t* row_nnz, cudaStream_t s, const LongStages* stages) {
    if (w.hi <= w.lo) return 0;
    uint32_t kernels = 0;
    auto on = [&](const char* what, uint32_t grid, uint64_t products) { if (stages) stages->on(what, grid, products); };
    auto off = [&]() { if (stages) stages->off(); };
    const LongDev D = long_dev(P, w.lo);
    on("long_sort", (uint32_t)w.unit_bound, w.products_bound);
    // sort: 32-bit (column << 12 | arrival) keys whenever they fit
    if ((uint64_t)b.cols <= (1ull << (32 - LONG_UNIT_LOG))) chunk_sort_launch<uint32_t, false>(a, b, row_begin, aseq, D, w, s);
    else if ((uint64_t)b.cols <= (1ull << (33 - LONG_UNIT_LOG))) chunk_sort_launch<uint32_t, true>(a, b, row_begin, aseq, D, w, s);
    else chunk_sort_launch<uint64_t, false>(a, b, row_begin, aseq, D, w, s);
    kernels += 1;
    off();
    // merge levels: rows are listed by ascending level count, level l takes the list from level_lo[l] on
    const size_t mgsmem = 2 * sizeof(MergeStage);
    static PerDeviceOnce attr;
    if (attr.first()) cudaFuncSetAttribute(k_long_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mgsmem);
    static int merge_ctas = 0;   // persistent grid: two CTAs per SM
    if (!merge_ctas) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        merge_ctas = 2 * sms;
    }
    MergeTile* tiles = reinterpret_cast<MergeTile*>(P.tiles);
    for (int l = 1; l <= w.max_level; ++l) {
        if (w.level_grid[l] == 0) continue;
        char lname[24];
        snprintf(lname, sizeof(lname), "long_merge_L%d", l);
        on(lname, w.level_grid[l], w.level_products[l]);
        const unsigned grid = std::min<unsigned>(w.level_grid[l], (unsigned)merge_ctas);
        k_long_partition<<<(w.level_grid[l] + 255) / 256, 256, 0, s>>>(D, w.level_lo[l], w.hi, l, tiles);
        k_long_merge<<<grid, MG_THREADS, mgsmem, s>>>(tiles, D, w.level_lo[l], w.hi, P.unit_heads, row_nnz);
        kernels += 2;
        off();
    }
    return kernels;
}

// after the last wave: where every tile's run heads go inside its row
uint32_t launch_long_heads_scan(const LongPlan& P, PlanCounters* ctr, cudaStream_t s) {
    if (P.n_rows == 0) return 0;
    launch_scan_u32_i64(P.unit_heads, (int64_t)P.unit_bound, P.unit_hoff, P.tile_state, ctr, s);
    return 1;
}

// second half: every head sums its run left to right straight into C (every destination of dst)
uint32_t launch_long_reduce(const LongPlan& P, const int64_t* c_ptr, const CopyDst& dst, cudaStream_t s) {
    if (P.n_rows == 0) return 0;
    const size_t msmem = (sizeof(int32_t) + sizeof(double)) * (LONG_UNIT + RD_PAD);
    static PerDeviceOnce rattr;
    if (rattr.first()) {
        cudaFuncSetAttribute(k_long_reduce<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
        cudaFuncSetAttribute(k_long_reduce<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
    }
    if (dst.n == 1)
        k_long_reduce<1><<<(unsigned)P.unit_bound, LR_THREADS, msmem, s>>>(long_dev(P, 0), P.n_rows, P.unit_hoff, c_ptr, dst);
    else
        k_long_reduce<8><<<(unsigned)P.unit_bound, LR_THREADS, msmem, s>>>(long_dev(P, 0), P.n_rows, P.unit_hoff, c_ptr, dst);
    return 1;
}

}  // namespace spada
