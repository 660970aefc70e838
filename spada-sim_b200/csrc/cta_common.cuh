// cta_common.cuh -- pieces shared by the CTA-per-row sort kernels (esc_cta_bitonic.cu, esc_tiled.cu):
// the CTA-wide flattened product expansion, the hybrid bitonic sort (register chunks merged through
// shared memory) and the ballot-based segmented reduce + store.
#pragma once
#include "common.cuh"
#include "sort.cuh"

namespace spada {

constexpr int ESC_CTA_THREADS = 256;

// =============================================================================================
// CTA-per-row kernels, N = 1024 / 2048 / 4096 products at most; 8 warps, chunk = N/8 keys per warp
// =============================================================================================

// Shared staging of one batch of A entries (one per thread): arrival offset, start of the B row,
// A value.  The products of the batch are then dealt to ALL threads of the CTA (thread t takes
// products t, t+256, ...), each finding its A entry by binary search over the offsets -- the
// expansion stays busy on every warp even when the A row has few, long-ish B rows to visit.
struct CtaStage {
    int off[ESC_CTA_THREADS + 1];
    int64_t bs[ESC_CTA_THREADS];
    double av[ESC_CTA_THREADS];
    int wtot[ESC_CTA_THREADS / 32];
};

// PACKED: keys carry the arrival index (column << log2(N/G) | seq inside the group); LOAD_COL = false: values only
template <typename K, int N, bool NUMERIC, bool PACKED = NUMERIC, bool LOAD_COL = true, int G = 1>
__device__ __forceinline__ int bitonic_cta_expand(const DevCsr& a, const DevCsr& b, int64_t a_begin, int64_t a_end,
                                                  K* keys, double* vals, CtaStage& st) {
    constexpr int SB = Log2<N / G>::v;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int seq_base = 0;
    for (int64_t pb = a_begin; pb < a_end; pb += ESC_CTA_THREADS) {
        const int64_t p = pb + threadIdx.x;
        int len = 0;
        int64_t bs = 0;
        double av = 0.0;
        if (p < a_end) {
            int32_t k = ldg_i32(a.col + p);
            if (NUMERIC) av = ldg_f64(a.val + p);
            b_row(b, k, bs, len);
        }
        int wtotal;
        int woff = warp_excl_scan(len, lane, wtotal);
        if (lane == 0) st.wtot[warp] = wtotal;
        __syncthreads();
        int base = 0, all = 0;
#pragma unroll
        for (int w = 0; w < ESC_CTA_THREADS / 32; ++w) {
            int t = st.wtot[w];
            if (w < warp) base += t;
            all += t;
        }
        st.off[threadIdx.x] = base + woff;
        st.bs[threadIdx.x] = bs;
        if (NUMERIC) st.av[threadIdx.x] = av;
        if (threadIdx.x == 0) st.off[ESC_CTA_THREADS] = all;
        __syncthreads();
        const int n_ent = (int)((a_end - pb) < ESC_CTA_THREADS ? (a_end - pb) : ESC_CTA_THREADS);
        for (int t0 = threadIdx.x; t0 < all; t0 += 2 * ESC_CTA_THREADS) {
            // two products per step: independent searches and gathers in flight
            int t[2] = {t0, t0 + ESC_CTA_THREADS};
            int64_t q[2];
            int j[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                int lo = 0, hi = n_ent;  // largest j in [0, n_ent) with off[j] <= t
                if (t[u] < all) {
                    while (hi - lo > 1) {
                        int mid = (lo + hi) >> 1;
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
                if (t[u] < all) {
                    if (LOAD_COL) c[u] = (uint32_t)ldg_i32(b.col + q[u]);
                    if (NUMERIC) bv[u] = ldg_f64(b.val + q[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (t[u] < all) {
                    int sq = seq_base + t[u];
                    if (LOAD_COL) keys[sq] = PACKED ? (((K)c[u] << SB) | (K)(sq & (N / G - 1))) : (K)c[u];
                    if (NUMERIC) vals[sq] = __dmul_rn(st.av[j[u]], bv[u]);
                }
            }
        }
        seq_base += all;
        __syncthreads();
    }
    return seq_base;
}

// G > 1: the N keys are G independent groups of N / G keys (each sorted on its own: the merge phases stop at N / G)
template <typename K, int N, int G = 1>
__device__ __forceinline__ void bitonic_cta_sort(K* keys) {
    constexpr int WARPS = ESC_CTA_THREADS / 32;
    constexpr int CH = N / WARPS;  // keys per warp chunk
    constexpr int E = CH / 32;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    K x[E];
    load_blocked<K, E>(x, keys + warp * CH, lane);
    warp_sort<K, E>(x, lane);
    store_blocked<K, E>(x, keys + warp * CH, lane);
    __syncthreads();
#pragma unroll 1
    for (int k = 2 * CH; k <= N / G; k <<= 1) {
        // flip stage: i against its mirror image inside the block of k keys
        for (int t = threadIdx.x; t < N / 2; t += ESC_CTA_THREADS) {
            const int h = k >> 1;
            const int i = ((t & ~(h - 1)) << 1) | (t & (h - 1));
            const int l = i ^ (k - 1);
            const K ka = keys[i], kb = keys[l];
            if (ka > kb) {
                keys[i] = kb;
                keys[l] = ka;
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int j = k >> 2; j >= CH; j >>= 1) {
            for (int t = threadIdx.x; t < N / 2; t += ESC_CTA_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const K ka = keys[i], kb = keys[l];
                if (ka > kb) {
                    keys[i] = kb;
                    keys[l] = ka;
                }
            }
            __syncthreads();
        }
        load_blocked<K, E>(x, keys + warp * CH, lane);
        warp_merge_tail<K, E>(x, lane);
        store_blocked<K, E>(x, keys + warp * CH, lane);
        __syncthreads();
    }
}

// Two sorted groups of N / 2 packed keys (column << log2(N/2) | arrival inside the group; group 0 = the earlier
// arrivals) merged into one sorted sequence: a stable merge by column, ties to group 0, so equal columns stay in
// arrival order.  32-bit keys then serve columns up to 2^(32 - log2(N/2)) -- twice as many as one group of N.
// On return keys[i] = column and vals[i] = value of the i-th product in (column, arrival) order.
template <int N>
__device__ __forceinline__ void cta_merge_groups2(uint32_t* keys, double* vals, int cnt) {
    constexpr int H = N / 2;
    constexpr int SBH = Log2<H>::v;
    constexpr int ITEMS = N / ESC_CTA_THREADS;
    const int c0 = cnt < H ? cnt : H, c1 = cnt - c0;
    const uint32_t* X = keys;
    const uint32_t* Y = keys + H;
    const int d = threadIdx.x * ITEMS;
    uint32_t oc[ITEMS];
    double ov[ITEMS];
    if (d < cnt) {
        int lo = d > c1 ? d - c1 : 0, hi = d < c0 ? d : c0;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((X[mid] >> SBH) <= (Y[d - 1 - mid] >> SBH)) lo = mid + 1; else hi = mid;
        }
        int i = lo, j = d - lo;
        uint32_t xk = i < c0 ? X[i] : 0xffffffffu, yk = j < c1 ? Y[j] : 0xffffffffu;
#pragma unroll
        for (int q = 0; q < ITEMS; ++q) {
            if (d + q < cnt) {
                const bool tx = j >= c1 || (i < c0 && (xk >> SBH) <= (yk >> SBH));
                const uint32_t key = tx ? xk : yk;
                oc[q] = key >> SBH;
                ov[q] = vals[(int)(key & (uint32_t)(H - 1)) + (tx ? 0 : H)];
                if (tx) {
                    ++i;
                    xk = i < c0 ? X[i] : 0xffffffffu;
                } else {
                    ++j;
                    yk = j < c1 ? Y[j] : 0xffffffffu;
                }
            }
        }
    }
    __syncthreads();
    if (d < cnt) {
#pragma unroll
        for (int q = 0; q < ITEMS; ++q)
            if (d + q < cnt) {
                keys[d + q] = oc[q];
                vals[d + q] = ov[q];
            }
    }
    __syncthreads();
}

// Segmented sums + store of a sorted row held in shared memory.  Warp w owns the positions
// [w*32*ITEMS, (w+1)*32*ITEMS), lanes interleaved (position = base + e*32 + lane: conflict-free
// shared-memory reads, coalesced stores); run heads are found with ballots, one scan over the eight
// warp totals places them, then every head sums its run left to right.
// SORTED: keys[i] is the column itself and vals[i] its value (after cta_merge_groups2) instead of packed keys that
// index vals by arrival
template <typename K, int N, bool SORTED = false>
__device__ __forceinline__ int cta_reduce_store(const K* keys, const double* vals, int p, int64_t cbase,
                                                int32_t* __restrict__ c_col, double* __restrict__ c_val, CtaStage& st,
                                                uint32_t col_offset = 0) {
    constexpr int SB = SORTED ? 0 : Log2<N>::v;
    constexpr int ITEMS = N / ESC_CTA_THREADS;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int w0 = warp * 32 * ITEMS;
    unsigned hm[ITEMS];
    uint32_t col[ITEMS];
    uint32_t carry = (w0 > 0 && w0 <= p) ? (uint32_t)(keys[w0 - 1] >> SB) : 0xffffffffu;
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < ITEMS; ++e) {
        const int i = w0 + e * 32 + lane;
        col[e] = (i < p) ? (uint32_t)(keys[i] >> SB) : 0xffffffffu;
        uint32_t cp = __shfl_up_sync(FULL, col[e], 1);
        if (lane == 0) cp = carry;
        const bool head = (i < p) && (i == 0 || cp != col[e]);
        hm[e] = __ballot_sync(FULL, head);
        carry = __shfl_sync(FULL, col[e], 31);
        cnt += __popc(hm[e]);
    }
    __syncthreads();
    if (lane == 0) st.wtot[warp] = cnt;
    __syncthreads();
    int o = 0, total = 0;
#pragma unroll
    for (int w = 0; w < ESC_CTA_THREADS / 32; ++w) {
        if (w < warp) o += st.wtot[w];
        total += st.wtot[w];
    }
#pragma unroll
    for (int e = 0; e < ITEMS; ++e) {
        if ((hm[e] >> lane) & 1u) {
            const int i = w0 + e * 32 + lane;
            double sum = vals[SORTED ? i : (int)(keys[i] & (K)(N - 1))];
            for (int j = i + 1; j < p; ++j) {
                const K kj = keys[j];
                if ((uint32_t)(kj >> SB) != col[e]) break;
                sum = __dadd_rn(sum, vals[SORTED ? j : (int)(kj & (K)(N - 1))]);
            }
            const int oo = o + __popc(hm[e] & ((1u << lane) - 1u));
            st_out(c_col + (cbase + oo), (int32_t)(col[e] + col_offset));
            st_out(c_val + (cbase + oo), sum);
        }
        o += __popc(hm[e]);
    }
    return total;
}

}  // namespace spada
