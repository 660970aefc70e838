// cta_common.cuh -- pieces shared by the CTA-wide sort kernels (esc_cta_bitonic.cu: one row per CTA; longrow.cu: one
// chunk of a long row per CTA):
// the CTA-wide flattened product expansion, the hybrid bitonic sort (register chunks merged through
// shared memory) and the ballot-based segmented reduce + store.
#pragma once
#include "common.cuh"
#include "sort.cuh"

namespace spada {

constexpr int ESC_CTA_THREADS = 256;
#ifndef SPADA_CTA_EXPAND_UNROLL
#define SPADA_CTA_EXPAND_UNROLL 2
#endif
constexpr int CTA_EXPAND_UNROLL = SPADA_CTA_EXPAND_UNROLL;   // products a thread expands per step (B gathers in flight)

// Where key e of a CTA-wide sort lives in shared memory.  The register phases move every lane's N/256 consecutive keys
// as 16-byte vectors: with U vectors per lane the eight lanes of a quarter-warp (one 128-byte transaction) would meet
// in 8/U bank groups (4-way conflicts for 4096 32-bit keys, 8-way for 64-bit ones).  XOR-ing the vector index with the
// lane bits above it spreads them over all eight; consecutive keys read by consecutive lanes (the shared-memory
// phases, the expansion, the reduce) stay inside one 128-byte line, permuted, so they remain conflict-free.
#ifndef SPADA_KEY_SWIZZLE
#define SPADA_KEY_SWIZZLE 1
#endif
template <typename K, int N>
struct KeySlot {
    static constexpr int LE = sizeof(K) == 4 ? 2 : 1;                       // log2(keys per 16-byte vector)
    static constexpr int U = (N / ESC_CTA_THREADS) * (int)sizeof(K) / 16;   // vectors per lane
    // two vectors per lane (2-way conflicts only): the index arithmetic costs more than the conflicts -- measured on
    // B200: sort_pass<2048> 0.635 -> 0.726 ms with the map, sort_pass<4096> 0.657 -> 0.645, chunk sort 0.946 -> 0.895
    static constexpr bool ON = SPADA_KEY_SWIZZLE && U > 2;
    __device__ __forceinline__ static int at(int e) {
        if constexpr (ON) return e ^ (((e >> (LE + 3)) & (U - 1)) << LE);
        return e;
    }
    __device__ __forceinline__ static int vec(int v) {   // the same map on vector indices
        if constexpr (ON) return v ^ ((v >> 3) & (U - 1));
        return v;
    }
};
// a lane's E consecutive keys (logical positions first + lane * E ...) to and from registers
template <typename K, int N, int E>
__device__ __forceinline__ void load_lane_keys(K (&x)[E], const K* keys, int first, int lane) {
    constexpr int V = E * (int)sizeof(K) / 16;
    static_assert(V >= 1, "at least one 16-byte vector per lane");
    const uint4* base = reinterpret_cast<const uint4*>(keys);
    const int v0 = (first + lane * E) >> KeySlot<K, N>::LE;
    uint4 tmp[V];
#pragma unroll
    for (int i = 0; i < V; ++i) tmp[i] = base[KeySlot<K, N>::vec(v0 + i)];
    memcpy(x, tmp, sizeof(tmp));
}
template <typename K, int N, int E>
__device__ __forceinline__ void store_lane_keys(const K (&x)[E], K* keys, int first, int lane) {
    constexpr int V = E * (int)sizeof(K) / 16;
    uint4* base = reinterpret_cast<uint4*>(keys);
    const int v0 = (first + lane * E) >> KeySlot<K, N>::LE;
    uint4 tmp[V];
    memcpy(tmp, x, sizeof(tmp));
#pragma unroll
    for (int i = 0; i < V; ++i) base[KeySlot<K, N>::vec(v0 + i)] = tmp[i];
}

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

// One bit short of 32-bit keys (column ids of 33 - log2 N bits): the key carries the column without its top bit and the
// top bit goes to a bitmap indexed by arrival; after the sort one stable split by that bit restores the full order
// (cta_split_top).  The lanes of a warp hold consecutive arrivals: one ballot, at most two words of the bitmap.
__device__ __forceinline__ void top_bit_mark(uint32_t* top, int idx, bool is_top, int lane) {
    const unsigned am = __activemask();
    const unsigned bal = __ballot_sync(am, is_top);
    if (bal && lane == __ffs(am) - 1) {
        const int idx0 = idx - lane;   // arrival of lane 0 (active or not)
        const int sh = idx0 & 31;
        atomicOr(&top[idx0 >> 5], bal << sh);
        if (sh && (bal >> (32 - sh))) atomicOr(&top[(idx0 >> 5) + 1], bal >> (32 - sh));
    }
}

// Keys carry the arrival index (column << log2 N | arrival).
// SPLIT (32-bit keys only): the column's bit 32 - log2 N is recorded in `top` instead of in the key.
template <typename K, int N, bool SPLIT = false>
__device__ __forceinline__ int bitonic_cta_expand(const DevCsr& a, const DevCsr& b, int64_t a_begin, int64_t a_end,
                                                  K* keys, double* vals, CtaStage& st, uint32_t* top = nullptr) {
    constexpr int SB = Log2<N>::v;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int seq_base = 0;
    for (int64_t pb = a_begin; pb < a_end; pb += ESC_CTA_THREADS) {
        const int64_t p = pb + threadIdx.x;
        int len = 0;
        int64_t bs = 0;
        double av = 0.0;
        if (p < a_end) {
            int32_t k = ldg_i32(a.col + p);
            av = ldg_f64(a.val + p);
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
        st.av[threadIdx.x] = av;
        if (threadIdx.x == 0) st.off[ESC_CTA_THREADS] = all;
        __syncthreads();
        const int n_ent = (int)((a_end - pb) < ESC_CTA_THREADS ? (a_end - pb) : ESC_CTA_THREADS);
        for (int t0 = threadIdx.x; t0 < all; t0 += CTA_EXPAND_UNROLL * ESC_CTA_THREADS) {
            // several products per step: independent searches and gathers in flight
            int t[CTA_EXPAND_UNROLL];
            int64_t q[CTA_EXPAND_UNROLL];
            int j[CTA_EXPAND_UNROLL];
#pragma unroll
            for (int u = 0; u < CTA_EXPAND_UNROLL; ++u) t[u] = t0 + u * ESC_CTA_THREADS;
#pragma unroll
            for (int u = 0; u < CTA_EXPAND_UNROLL; ++u) {
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
            uint32_t c[CTA_EXPAND_UNROLL];
            double bv[CTA_EXPAND_UNROLL];
#pragma unroll
            for (int u = 0; u < CTA_EXPAND_UNROLL; ++u) {
                c[u] = 0;
                bv[u] = 0.0;
                if (t[u] < all) {
                    c[u] = (uint32_t)ldg_i32(b.col + q[u]);
                    bv[u] = ldg_f64(b.val + q[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < CTA_EXPAND_UNROLL; ++u) {
                if (t[u] < all) {
                    int sq = seq_base + t[u];
                    keys[KeySlot<K, N>::at(sq)] = ((K)c[u] << SB) | (K)(sq & (N - 1));   // SPLIT: the top bit falls off
                    vals[sq] = __dmul_rn(st.av[j[u]], bv[u]);
                    if constexpr (SPLIT) top_bit_mark(top, sq, (c[u] >> (32 - SB)) & 1u, lane);
                }
            }
        }
        seq_base += all;
        __syncthreads();
    }
    return seq_base;
}

template <typename K, int N>
__device__ __forceinline__ void bitonic_cta_sort(K* keys) {
    constexpr int WARPS = ESC_CTA_THREADS / 32;
    constexpr int CH = N / WARPS;  // keys per warp chunk
    constexpr int E = CH / 32;
    using S = KeySlot<K, N>;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    K x[E];
    load_lane_keys<K, N, E>(x, keys, warp * CH, lane);
    warp_sort<K, E>(x, lane);
    store_lane_keys<K, N, E>(x, keys, warp * CH, lane);
    __syncthreads();
#pragma unroll 1
    for (int k = 2 * CH; k <= N; k <<= 1) {
        // flip stage: i against its mirror image inside the block of k keys
        for (int t = threadIdx.x; t < N / 2; t += ESC_CTA_THREADS) {
            const int h = k >> 1;
            const int i = S::at(((t & ~(h - 1)) << 1) | (t & (h - 1)));
            const int l = S::at((((t & ~(h - 1)) << 1) | (t & (h - 1))) ^ (k - 1));
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
                const int i0 = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int i = S::at(i0), l = S::at(i0 | j);
                const K ka = keys[i], kb = keys[l];
                if (ka > kb) {
                    keys[i] = kb;
                    keys[l] = ka;
                }
            }
            __syncthreads();
        }
        load_lane_keys<K, N, E>(x, keys, warp * CH, lane);
        warp_merge_tail<K, E>(x, lane);
        store_lane_keys<K, N, E>(x, keys, warp * CH, lane);
        __syncthreads();
    }
}

// After a sort by (column without its top bit, arrival): one stable split by the top bit (bitmap `top`, indexed by
// arrival) puts every key where the full (column, arrival) order wants it -- 32-bit keys then serve twice as many
// columns (the 64-bit network measured 3.2x slower per product; two half-size groups merged serially in shared
// memory, round 2's first version of this, cost as much as the sort itself: 16-way bank conflicts on its staging).
// Only the keys move; the values stay where their arrival index finds them.  Positions are dealt to the lanes
// interleaved (conflict-free reads); the keys of either half keep their order, so a warp writes two contiguous runs.
// Returns n0, the number of products whose top bit is clear: position i holds a column with the top bit set iff i >= n0.
template <int N>
__device__ __forceinline__ int cta_split_top(uint32_t* keys, int cnt, const uint32_t* top, CtaStage& st) {
    constexpr int ITEMS = N / ESC_CTA_THREADS;
    constexpr int WARPS = ESC_CTA_THREADS / 32;
    using S = KeySlot<uint32_t, N>;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int w0 = warp * 32 * ITEMS;
    uint32_t key[ITEMS];
    unsigned lowm[ITEMS];   // ballot of "valid and top bit clear"
    int nlow = 0;
#pragma unroll
    for (int e = 0; e < ITEMS; ++e) {
        const int i = w0 + e * 32 + lane;
        bool low = false;
        key[e] = 0;
        if (i < cnt) {
            key[e] = keys[S::at(i)];
            const int t = (int)(key[e] & (uint32_t)(N - 1));
            low = !((top[t >> 5] >> (t & 31)) & 1u);
        }
        lowm[e] = __ballot_sync(FULL, low);
        nlow += __popc(lowm[e]);
    }
    __syncthreads();   // every key is in registers
    if (lane == 0) st.wtot[warp] = nlow;
    __syncthreads();
    int zbase = 0, n0 = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        if (w < warp) zbase += st.wtot[w];
        n0 += st.wtot[w];
    }
#pragma unroll
    for (int e = 0; e < ITEMS; ++e) {
        const int i = w0 + e * 32 + lane;
        if (i < cnt) {
            const int z = zbase + __popc(lowm[e] & ((1u << lane) - 1u));   // clear-bit products before i
            keys[S::at(((lowm[e] >> lane) & 1u) ? z : n0 + (i - z))] = key[e];
        }
        zbase += __popc(lowm[e]);
    }
    __syncthreads();
    return n0;
}

// Segmented sums + store of a sorted row held in shared memory.  Warp w owns the positions
// [w*32*ITEMS, (w+1)*32*ITEMS), lanes interleaved (position = base + e*32 + lane: conflict-free
// shared-memory reads, coalesced stores); run heads are found with ballots, one scan over the eight
// warp totals places them, then every head sums its run left to right.
// n0 < p (after cta_split_top): the columns at positions >= n0 carry one more bit, the one the key has no room for
template <typename K, int N>
__device__ __forceinline__ int cta_reduce_store(const K* keys, const double* vals, int p, int64_t cbase,
                                                int32_t* __restrict__ c_col, double* __restrict__ c_val, CtaStage& st,
                                                int n0 = 0x7fffffff) {
    constexpr int SB = Log2<N>::v;
    constexpr int ITEMS = N / ESC_CTA_THREADS;
    constexpr uint32_t TOP = sizeof(K) == 4 ? 1u << (32 - SB) : 0u;
    using S = KeySlot<K, N>;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int w0 = warp * 32 * ITEMS;
    unsigned hm[ITEMS];
    uint32_t col[ITEMS];
    uint32_t carry = (w0 > 0 && w0 <= p) ? (uint32_t)(keys[S::at(w0 - 1)] >> SB) | (w0 - 1 >= n0 ? TOP : 0u) : 0xffffffffu;
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < ITEMS; ++e) {
        const int i = w0 + e * 32 + lane;
        col[e] = (i < p) ? (uint32_t)(keys[S::at(i)] >> SB) | (i >= n0 ? TOP : 0u) : 0xffffffffu;
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
            const int end = i < n0 && n0 < p ? n0 : p;   // a run never crosses the split
            const uint32_t low = col[e] & ~TOP;
            double sum = vals[(int)(keys[S::at(i)] & (K)(N - 1))];
            for (int j = i + 1; j < end; ++j) {
                const K kj = keys[S::at(j)];
                if ((uint32_t)(kj >> SB) != low) break;
                sum = __dadd_rn(sum, vals[(int)(kj & (K)(N - 1))]);
            }
            const int oo = o + __popc(hm[e] & ((1u << lane) - 1u));
            st_out(c_col + (cbase + oo), (int32_t)col[e]);
            st_out(c_val + (cbase + oo), sum);
        }
        o += __popc(hm[e]);
    }
    return total;
}

}  // namespace spada
