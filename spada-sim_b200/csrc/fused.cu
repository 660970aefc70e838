// fused.cu -- stages 2+3+4 in ONE pass for rows with at most 512 intermediate products (bins 1..5).
//
// The two-phase path sorts every such row twice (once to count, once to compute).  Here a row is
// expanded, sorted and reduced once; the reduced row waits in the warp's shared memory while the
// CTA learns where it goes: eight consecutive rows form a tile, the tile's nnz total is
// published and the exclusive prefix over all earlier tiles is obtained by decoupled look-back
// (the same single-pass scan as plan.cu, one tile per CTA; tile id = a ticket drawn at CTA start).
// Then row_ptr is written and the rows are copied to their final place with coalesced stores.
//
// Rows with more than 512 products (CTA-per-row sort bins, long rows) are computed beforehand into
// scratch rows by their own kernels, which also record their nnz; the tile sums simply include those
// counts, and the scratch rows are copied to their place afterwards against the finished row_ptr.
//
// C is allocated with capacity = number of intermediate products (an upper bound on nnz(C))
// because nnz(C) is only known when this kernel ends; the engine falls back to the two-phase
// path when that bound does not fit.
//
// Reference logic replaced: one PE pass (simulator.rs:86-111, 143-171, 199-230), write_psums
// (simulator.rs:955-983) and the indptr maintenance of CsrMatStorage::write (storage.rs:196-210).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sort.cuh"

namespace spada {

#ifndef SPADA_FUSED_WARPS
#define SPADA_FUSED_WARPS 8
#endif
constexpr int FUSED_WARPS = SPADA_FUSED_WARPS;  // warps per tile of the tiny-row kernel (4 rows each; at most 8: one 32-row scan)
static_assert(FUSED_WARPS <= 8, "the tile scan covers 32 rows");
#ifndef SPADA_LIGHT_WARPS
#define SPADA_LIGHT_WARPS 4
#endif
constexpr int LIGHT_WARPS = SPADA_LIGHT_WARPS;  // warps per tile of the general kernel; each warp owns RPW consecutive rows

#define FST_AGG (1ull << 62)
#define FST_PREFIX (2ull << 62)
#define FST_MASK (3ull << 62)

// sort the warp's N = 32*E packed keys in registers, leave them sorted in shared memory (blocked
// layout = sorted order) and return the number of distinct columns among the first p keys
template <typename K, int E, int SBK>
__device__ __forceinline__ int sort_count(K* keys, int p, int lane) {
    constexpr int N = 32 * E;
    for (int t = p + lane; t < N; t += 32) keys[t] = KeyTraits<K>::sentinel;
    __syncwarp();
    K x[E];
    load_blocked<K, E>(x, keys, lane);
    warp_sort<K, E>(x, lane);
    __syncwarp();
    store_blocked<K, E>(x, keys, lane);
    const K prev = shfl_up_key(x[E - 1]);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const bool first = (lane == 0 && i == 0);
        const K pv = (i == 0) ? prev : x[i - 1];
        if (x[i] != KeyTraits<K>::sentinel && (first || (uint32_t)(x[i] >> SBK) != (uint32_t)(pv >> SBK))) ++cnt;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
    __syncwarp();
    return cnt;
}

// sums equal columns of the sorted row left to right and stores the row at c_col/c_val + base
template <typename K, int SBK>
__device__ __forceinline__ void reduce_store(const K* keys, const double* vals, int p, int lane, int64_t base,
                                             int32_t* __restrict__ c_col, double* __restrict__ c_val) {
    int out_base = 0;
    uint32_t prev_last = 0;
    for (int cb = 0; cb < p; cb += 32) {
        const int i = cb + lane;
        const bool valid = i < p;
        const K ki = valid ? keys[i] : KeyTraits<K>::sentinel;
        const uint32_t col = (uint32_t)(ki >> SBK);
        uint32_t col_prev = __shfl_up_sync(FULL, col, 1);
        if (lane == 0) col_prev = prev_last;
        const bool head = valid && (i == 0 || col_prev != col);
        const unsigned hm = __ballot_sync(FULL, head);
        prev_last = __shfl_sync(FULL, col, 31);
        if (head) {
            double sum = vals[(int)(ki & (K)((1u << SBK) - 1u))];
            for (int j = i + 1; j < p; ++j) {
                const K kj = keys[j];
                if ((uint32_t)(kj >> SBK) != col) break;
                sum = __dadd_rn(sum, vals[(int)(kj & (K)((1u << SBK) - 1u))]);
            }
            const int o = out_base + __popc(hm & ((1u << lane) - 1u));
            st_out(c_col + (base + o), (int32_t)col);
            st_out(c_val + (base + o), sum);
        }
        out_base += __popc(hm);
    }
}

template <typename K, int NMAX, int RPW>
__global__ void __launch_bounds__(LIGHT_WARPS * 32)
k_fused_light(DevCsr a, DevCsr b, int64_t row_begin, int64_t m, const uint32_t* __restrict__ flops,
              const uint32_t* __restrict__ pre_nnz, int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col,
              double* __restrict__ c_val, unsigned long long* tile_state, uint32_t* ticket) {
    constexpr int SBK = Log2<NMAX>::v;
    constexpr int TILE_ROWS = LIGHT_WARPS * RPW;
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ uint32_t s_nnz[TILE_ROWS];
    __shared__ unsigned long long s_excl;
    __shared__ uint32_t s_tile;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    K* keys_w = reinterpret_cast<K*>(s_raw) + (size_t)warp * RPW * NMAX;
    double* vals_w = reinterpret_cast<double*>(s_raw + sizeof(K) * NMAX * TILE_ROWS) + (size_t)warp * RPW * NMAX;

    // Tile id = a ticket drawn when the CTA starts: every tile this one looks back at is held by a CTA that is
    // already running (or done), whatever order the hardware dispatches the grid in and whatever else shares the
    // GPU (side streams, NCCL kernels) -- the look-back can always make progress.  (Persistent CTAs that draw the next
    // ticket while the current tile is placed were measured: ER 5.2 -> 11.9 ms, Poisson 1.34 -> 1.53 ms; reverted.)
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t r0 = (int64_t)tile * TILE_ROWS + warp * RPW;

    int nnz[RPW];
    int prod[RPW];   // products of a row computed here, -1 otherwise
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
        nnz[q] = 0;
        prod[q] = -1;
        const int64_t r = r0 + q;
        K* keys = keys_w + q * NMAX;
        double* vals = vals_w + q * NMAX;
        if (r < m) {
            const uint32_t pf = flops[r];
            const int bn = bin_of(pf);
            if (bn >= 1 && bn <= 5) {
                const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
                int seq = 0;
                for (int64_t pb = a_begin; pb < a_end; pb += 32) {
                    int bt;
                    expand_batch<true, false>(a, b, pb + lane, a_end, lane, seq, bt,
                                              [&](int sq, uint32_t c, double av, double bv) {
                                                  keys[sq] = ((K)c << SBK) | (K)sq;
                                                  vals[sq] = __dmul_rn(av, bv);
                                              });
                    seq += bt;
                }
                prod[q] = seq;
                // the window shape of this row: 32 lanes x E keys per lane, E picked by its product count
                switch (bn) {
                    case 1: nnz[q] = sort_count<K, 1, SBK>(keys, seq, lane); break;
                    case 2: if constexpr (NMAX >= 64) nnz[q] = sort_count<K, 2, SBK>(keys, seq, lane); break;
                    case 3: if constexpr (NMAX >= 128) nnz[q] = sort_count<K, 4, SBK>(keys, seq, lane); break;
                    case 4: if constexpr (NMAX >= 256) nnz[q] = sort_count<K, 8, SBK>(keys, seq, lane); break;
                    default: if constexpr (NMAX >= 512) nnz[q] = sort_count<K, 16, SBK>(keys, seq, lane); break;
                }
            } else if (bn >= 6) {
                nnz[q] = (int)pre_nnz[r];
            }
        }
        if (lane == 0) s_nnz[warp * RPW + q] = (uint32_t)nnz[q];
    }
    __syncthreads();
    unsigned long long tile_total = 0, my_off = 0;
#pragma unroll
    for (int w = 0; w < TILE_ROWS; ++w) {
        unsigned long long t = s_nnz[w];
        if (w < warp * RPW) my_off += t;
        tile_total += t;
    }
    if (warp == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&tile_state[0], FST_PREFIX | tile_total);
        } else {
            if (lane == 0) atomicExch(&tile_state[tile], FST_AGG | tile_total);
            int64_t pidx = (int64_t)tile - 1;
            while (true) {
                int64_t idx = pidx - lane;
                unsigned long long s;
                do {
                    s = (idx >= 0) ? *((volatile unsigned long long*)&tile_state[idx]) : FST_PREFIX;
                } while (__any_sync(FULL, (s & FST_MASK) == 0));
                unsigned has_prefix = __ballot_sync(FULL, (s & FST_MASK) == FST_PREFIX);
                unsigned long long val = s & ~FST_MASK;
                if (has_prefix) {
                    int first = __ffs(has_prefix) - 1;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(FULL, val, d);
                excl += val;
                if (has_prefix) break;
                pidx -= 32;
            }
            if (lane == 0) atomicExch(&tile_state[tile], FST_PREFIX | (excl + tile_total));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    int64_t base = (int64_t)(s_excl + my_off);
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
        const int64_t r = r0 + q;
        if (r < m) {
            if (lane == 0) {
                c_ptr[r] = base;
                if (r == m - 1) c_ptr[m] = base + nnz[q];
            }
            if (prod[q] >= 0)
                reduce_store<K, SBK>(keys_w + q * NMAX, vals_w + q * NMAX, prod[q], lane, base, c_col, c_val);
        }
        base += nnz[q];
    }
}

// ---------------------------------------------------------------------------------------------
// Tiny rows (at most 32 products, e.g. the 5-point stencil: 25 products, 13 outputs per row): the
// column-wise end of Spada's window shapes.  One product per lane, the whole row lives in
// registers: a 32-key bitonic network (15 compare-exchange steps), shuffles for the values, a
// ballot for the run heads.  32 consecutive rows form a tile (8 warps x 4 rows) that is placed by
// the same decoupled look-back as above.
constexpr int TINY_RPW = 4;
constexpr int TINY_TILE = FUSED_WARPS * TINY_RPW;
constexpr int TINY_LD = 40;  // row stride of the staging arrays: the four 8-lane groups of a warp land on disjoint banks

// one row (1..32 products) by a whole warp: leaves the finished row in rc/rv, returns its nnz
template <typename K>
__device__ __forceinline__ int tiny_row_warp(const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t r,
                                             uint32_t* rc, double* rv, int lane) {
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    K x[1];
    double prod = 0.0;
    int p;
    if (a_end - a_begin <= 32) {
        // the whole A row is one batch: one product per lane, nothing staged
        int64_t bs = 0;
        int len = 0;
        double av = 0.0;
        if (lane < (int)(a_end - a_begin)) {
            const int32_t k = ldg_i32(a.col + a_begin + lane);
            av = ldg_f64(a.val + a_begin + lane);
            b_row(b, k, bs, len);
        }
        const int off = warp_excl_scan(len, lane, p);
        int j = 0;
#pragma unroll
        for (int st = 16; st > 0; st >>= 1) {
            const int o = __shfl_sync(FULL, off, j + st);
            if (o <= lane) j += st;
        }
        const int oj = __shfl_sync(FULL, off, j);
        const int64_t bsj = shfl_i64(bs, j);
        const double aj = shfl_f64(av, j);
        x[0] = KeyTraits<K>::sentinel;
        if (lane < p) {
            const int64_t qq = bsj + (lane - oj);
            x[0] = ((K)(uint32_t)ldg_i32(b.col + qq) << 5) | (K)lane;
            prod = __dmul_rn(aj, ldg_f64(b.val + qq));
        }
    } else {
        // more than 32 A entries (most of them meeting empty B rows): stage through shared memory
        p = 0;
        for (int64_t pb = a_begin; pb < a_end; pb += 32) {
            int bt;
            expand_batch<true, false>(a, b, pb + lane, a_end, lane, p, bt, [&](int sq, uint32_t c, double av, double bv) {
                rc[sq] = c;
                rv[sq] = __dmul_rn(av, bv);
            });
            p += bt;
        }
        __syncwarp();
        x[0] = lane < p ? (((K)rc[lane] << 5) | (K)lane) : KeyTraits<K>::sentinel;
        prod = lane < p ? rv[lane] : 0.0;
        __syncwarp();
    }
    const bool have = lane < p;
    warp_sort<K, 1>(x, lane);
    const uint32_t col = (uint32_t)(x[0] >> 5);
    const double v = shfl_f64(prod, (int)(x[0] & (K)31));
    const uint32_t col_prev = __shfl_up_sync(FULL, col, 1);
    const bool head = have && (lane == 0 || col_prev != col);
    const unsigned hm = __ballot_sync(FULL, head);
    // run length of a head = distance to the next head (or to the end of the row)
    const unsigned later = lane < 31 ? (hm >> (lane + 1)) : 0u;
    const int run = head ? (later ? __ffs(later) : (p - lane)) : 0;
    const int max_run = __reduce_max_sync(FULL, run);
    double sum = v;
    for (int d = 1; d < max_run; ++d) {
        const double nv = shfl_f64(v, (lane + d) & 31);
        if (d < run) sum = __dadd_rn(sum, nv);
    }
    if (head) {
        const int pos = __popc(hm & ((1u << lane) - 1u));
        rc[pos] = col;
        rv[pos] = sum;
    }
    return __popc(hm);
}

// the tile's 32 finished rows wait in shared memory: scan their counts, look back for the tile's base,
// write row_ptr and copy the rows out (a warp stores its four rows as one contiguous stream)
__device__ __forceinline__ void tiny_tile_finish(uint32_t tile, int64_t m, uint32_t (*s_col)[TINY_LD], double (*s_val)[TINY_LD],
                                                 uint32_t* s_nnz, uint32_t* s_off, uint32_t* s_light,
                                                 unsigned long long* s_excl, int64_t* __restrict__ c_ptr,
                                                 int32_t* __restrict__ c_col, double* __restrict__ c_val,
                                                 unsigned long long* tile_state, int lane, int warp) {
    if (warp == 0) {
        // exclusive scan of the 32 row counts, then the look-back for the tile's base
        const uint32_t n = lane < TINY_TILE ? s_nnz[lane] : 0u;
        uint32_t x = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane < TINY_TILE) s_off[lane] = x - n;
        const unsigned long long tile_total = __shfl_sync(FULL, x, 31);
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&tile_state[0], FST_PREFIX | tile_total);
        } else {
            if (lane == 0) atomicExch(&tile_state[tile], FST_AGG | tile_total);
            int64_t pidx = (int64_t)tile - 1;
            while (true) {
                int64_t idx = pidx - lane;
                unsigned long long s;
                do {
                    s = (idx >= 0) ? *((volatile unsigned long long*)&tile_state[idx]) : FST_PREFIX;
                } while (__any_sync(FULL, (s & FST_MASK) == 0));
                unsigned has_prefix = __ballot_sync(FULL, (s & FST_MASK) == FST_PREFIX);
                unsigned long long val = s & ~FST_MASK;
                if (has_prefix) {
                    int first = __ffs(has_prefix) - 1;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(FULL, val, d);
                excl += val;
                if (has_prefix) break;
                pidx -= 32;
            }
            if (lane == 0) atomicExch(&tile_state[tile], FST_PREFIX | (excl + tile_total));
        }
        if (lane == 0) *s_excl = excl;
    }
    __syncthreads();
    // row_ptr: one thread per row of the tile
    if (threadIdx.x < TINY_TILE) {
        const int64_t r = (int64_t)tile * TINY_TILE + threadIdx.x;
        if (r < m) {
            const int64_t base = (int64_t)(*s_excl + s_off[threadIdx.x]);
            c_ptr[r] = base;
            if (r == m - 1) c_ptr[m] = base + s_nnz[threadIdx.x];
        }
    }
    // entries: a warp stores its four rows as one contiguous stream
    const int rt0 = warp * TINY_RPW;
    const uint32_t n0 = s_nnz[rt0], n1 = s_nnz[rt0 + 1], n2 = s_nnz[rt0 + 2], n3 = s_nnz[rt0 + 3];
    const uint32_t c1 = n0, c2 = n0 + n1, c3 = n0 + n1 + n2, tot = c3 + n3;
    const int64_t wbase = (int64_t)(*s_excl + s_off[rt0]);
    const uint32_t lightmask = (*s_light >> rt0) & 0xfu;
    for (uint32_t e = lane; e < tot; e += 32) {
        const int q = (e >= c1) + (e >= c2) + (e >= c3);
        if ((lightmask >> q) & 1u) {
            const uint32_t idx = e - (q == 0 ? 0u : (q == 1 ? c1 : (q == 2 ? c2 : c3)));
            st_out(c_col + (wbase + e), (int32_t)s_col[rt0 + q][idx]);
            st_out(c_val + (wbase + e), s_val[rt0 + q][idx]);
        }
    }
}

template <typename K>
__global__ void __launch_bounds__(FUSED_WARPS * 32)
k_fused_tiny(DevCsr a, DevCsr b, int64_t row_begin, int64_t m, const uint32_t* __restrict__ flops,
             const uint32_t* __restrict__ pre_nnz, int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col,
             double* __restrict__ c_val, unsigned long long* tile_state, uint32_t* ticket) {
    __shared__ uint32_t s_col[TINY_TILE][TINY_LD];
    __shared__ double s_val[TINY_TILE][TINY_LD];
    __shared__ uint32_t s_nnz[TINY_TILE];   // nnz of every row of the tile
    __shared__ uint32_t s_off[TINY_TILE];   // exclusive offsets inside the tile
    __shared__ uint32_t s_light;            // bit rt set <=> row rt was computed here
    __shared__ unsigned long long s_excl;
    __shared__ uint32_t s_tile;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_light = 0u;
        s_tile = atomicAdd(ticket, 1u);   // tile id = start order (see k_fused_light)
    }
    __syncthreads();
    const uint32_t tile = s_tile;

#pragma unroll 1
    for (int q = 0; q < TINY_RPW; ++q) {
        const int rt = warp * TINY_RPW + q;
        const int64_t r = (int64_t)tile * TINY_TILE + rt;
        int nnz = 0;
        if (r < m) {
            const uint32_t pf = flops[r];
            if (pf >= 1u && pf <= 32u) {
                nnz = tiny_row_warp<K>(a, b, row_begin, r, s_col[rt], s_val[rt], lane);
                if (lane == 0) atomicOr(&s_light, 1u << rt);
            } else if (pf > 32u) {
                nnz = (int)pre_nnz[r];   // computed into a scratch row beforehand
            }
        }
        if (lane == 0) s_nnz[rt] = (uint32_t)nnz;
    }
    __syncthreads();
    tiny_tile_finish(tile, m, s_col, s_val, s_nnz, s_off, &s_light, &s_excl, c_ptr, c_col, c_val, tile_state, lane, warp);
}

// ---------------------------------------------------------------------------------------------
// Tiny rows, four at a time: a warp is cut into four groups of 8 lanes, every group owns one row
// (at most 8 A entries, at most 32 products = 8 lanes x 4 keys).  Spada's window [R, L/R] with
// R = 4 rows sharing the 32 lanes: the scan of the B-row lengths, the search of a product's A entry
// and the 32-key sorting network all stay inside the group (shuffles of width 8, xor masks < 8), so one
// warp instruction advances four rows.  Measured against k_fused_tiny on the 5-point stencil: see
// profiles/.  Rows that do not fit (more than 8 A entries) fall back to the whole-warp path above.
template <typename K>
__device__ __forceinline__ K shfl_up1_w8(K v);
template <>
__device__ __forceinline__ uint32_t shfl_up1_w8<uint32_t>(uint32_t v) { return __shfl_up_sync(FULL, v, 1, 8); }
template <>
__device__ __forceinline__ uint64_t shfl_up1_w8<uint64_t>(uint64_t v) {
    const uint32_t lo = __shfl_up_sync(FULL, (uint32_t)v, 1, 8);
    const uint32_t hi = __shfl_up_sync(FULL, (uint32_t)(v >> 32), 1, 8);
    return ((uint64_t)hi << 32) | lo;
}

template <typename K>
__global__ void __launch_bounds__(FUSED_WARPS * 32)
k_fused_tiny4(DevCsr a, DevCsr b, int64_t row_begin, int64_t m, const uint32_t* __restrict__ flops,
              const uint32_t* __restrict__ pre_nnz, int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col,
              double* __restrict__ c_val, unsigned long long* tile_state, uint32_t* ticket) {
    __shared__ uint32_t s_col[TINY_TILE][TINY_LD];   // finished rows (compacted), staged for the coalesced store
    __shared__ double s_val[TINY_TILE][TINY_LD];     // first the products by arrival index, then the finished rows
    __shared__ uint32_t t_col[TINY_TILE][TINY_LD];   // the sorted row, read by the run sums
    __shared__ double t_val[TINY_TILE][TINY_LD];
    __shared__ uint32_t s_nnz[TINY_TILE];
    __shared__ uint32_t s_off[TINY_TILE];
    __shared__ uint32_t s_light;
    __shared__ unsigned long long s_excl;
    __shared__ uint32_t s_tile;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int g = lane >> 3, sub = lane & 7;
    if (threadIdx.x == 0) {
        s_light = 0u;
        s_tile = atomicAdd(ticket, 1u);   // tile id = start order (see k_fused_light)
    }
    __syncthreads();
    const uint32_t tile = s_tile;

    {
        const int rt = warp * TINY_RPW + g;
        const int64_t r = (int64_t)tile * TINY_TILE + rt;
        uint32_t pf = 0;
        int64_t a_begin = 0;
        int a_len = 0;
        if (r < m) {
            pf = flops[r];
            a_begin = a.ptr[row_begin + r];
            a_len = (int)(a.ptr[row_begin + r + 1] - a_begin);
        }
        const bool mine = pf >= 1u && pf <= 32u;   // computed here
        const bool fits = !mine || a_len <= 8;
        if (__all_sync(FULL, fits)) {
            // ---- four rows at once -------------------------------------------------------------------
            int len = 0;
            int64_t bs = 0;
            double av = 0.0;
            if (mine && sub < a_len) {
                const int32_t k = ldg_i32(a.col + a_begin + sub);
                av = ldg_f64(a.val + a_begin + sub);
                b_row(b, k, bs, len);
            }
            int x = len;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const int y = __shfl_up_sync(FULL, x, d, 8);
                if (sub >= d) x += y;
            }
            const int off = x - len;
            const int p = __shfl_sync(FULL, x, 7, 8);
            K key[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int sq = e * 8 + sub;
                int j = 0;
#pragma unroll
                for (int st = 4; st > 0; st >>= 1) {
                    const int o = __shfl_sync(FULL, off, j + st, 8);
                    if (o <= sq) j += st;
                }
                const int oj = __shfl_sync(FULL, off, j, 8);
                const int bs_lo = __shfl_sync(FULL, (int)(bs & 0xffffffffll), j, 8);
                const int bs_hi = __shfl_sync(FULL, (int)(bs >> 32), j, 8);
                const long long avb = __double_as_longlong(av);
                const int av_lo = __shfl_sync(FULL, (int)(avb & 0xffffffffll), j, 8);
                const int av_hi = __shfl_sync(FULL, (int)(avb >> 32), j, 8);
                key[e] = KeyTraits<K>::sentinel;
                if (sq < p) {
                    const int64_t qq = (((int64_t)bs_hi << 32) | (uint32_t)bs_lo) + (sq - oj);
                    const double aj = __longlong_as_double(((long long)av_hi << 32) | (uint32_t)av_lo);
                    key[e] = ((K)(uint32_t)ldg_i32(b.col + qq) << 5) | (K)sq;
                    s_val[rt][sq] = __dmul_rn(aj, ldg_f64(b.val + qq));
                }
            }
            __syncwarp();
            ChunkSort<K, 4, 32>::run(key, lane);   // strides < 8 lanes: every group sorts its own 32 keys
            uint32_t col[4];
            double v[4];
            unsigned validbits = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool valid = key[q] != KeyTraits<K>::sentinel;
                col[q] = (uint32_t)(key[q] >> 5);
                v[q] = valid ? s_val[rt][(int)(key[q] & (K)31)] : 0.0;
                validbits |= (valid ? 1u : 0u) << q;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if ((validbits >> q) & 1u) {
                    t_col[rt][sub * 4 + q] = col[q];
                    t_val[rt][sub * 4 + q] = v[q];
                }
            const K prevk = shfl_up1_w8<K>(key[3]);
            unsigned headbits = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t pc = (q == 0) ? (uint32_t)(prevk >> 5) : col[q - 1];
                const bool first = (sub == 0 && q == 0);
                if (((validbits >> q) & 1u) && (first || pc != col[q])) headbits |= 1u << q;
            }
            const int hc = __popc(headbits);
            int hx = hc;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const int y = __shfl_up_sync(FULL, hx, d, 8);
                if (sub >= d) hx += y;
            }
            const int nnz = __shfl_sync(FULL, hx, 7, 8);
            __syncwarp();   // products all fetched, sorted row visible: s_val may now take the finished row
            int o = hx - hc;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if ((headbits >> q) & 1u) {
                    double sum = v[q];
                    for (int jj = sub * 4 + q + 1; jj < p && t_col[rt][jj] == col[q]; ++jj) sum = __dadd_rn(sum, t_val[rt][jj]);
                    s_col[rt][o] = col[q];
                    s_val[rt][o] = sum;
                    ++o;
                }
            if (sub == 0) {
                s_nnz[rt] = mine ? (uint32_t)nnz : ((r < m && pf > 32u) ? pre_nnz[r] : 0u);
                if (mine) atomicOr(&s_light, 1u << rt);
            }
        } else {
            // ---- a row with more than 8 A entries: the warp takes its four rows one after the other -------
#pragma unroll 1
            for (int q = 0; q < TINY_RPW; ++q) {
                const int rtq = warp * TINY_RPW + q;
                const int64_t rq = (int64_t)tile * TINY_TILE + rtq;
                int nnz = 0;
                if (rq < m) {
                    const uint32_t pq = flops[rq];
                    if (pq >= 1u && pq <= 32u) {
                        nnz = tiny_row_warp<K>(a, b, row_begin, rq, s_col[rtq], s_val[rtq], lane);
                        if (lane == 0) atomicOr(&s_light, 1u << rtq);
                    } else if (pq > 32u) {
                        nnz = (int)pre_nnz[rq];
                    }
                }
                if (lane == 0) s_nnz[rtq] = (uint32_t)nnz;
            }
        }
    }
    __syncthreads();
    tiny_tile_finish(tile, m, s_col, s_val, s_nnz, s_off, &s_light, &s_excl, c_ptr, c_col, c_val, tile_state, lane, warp);
}

// ---------------------------------------------------------------------------------------------
// Mixed row lengths (heavy-tailed graphs): tiles are cut by WORK, not by row count -- Spada's window [R, L/R] with
// R adapting along the matrix (rowwise_perf_adjust.rs:36-77 groups consecutive rows of similar length; 121-252 picks
// R per group): a tile holds as many consecutive rows as fit a fixed budget of sort slots, so a stretch of short rows
// shares one tile (R up to 64) while a 512-product row has a tile almost to itself (R = 1..4).  Every tile costs about
// the same, which is what keeps the decoupled look-back flowing: with a fixed row count per tile (k_fused_light) a
// tile finished with its slowest row and everything placed after it waited (measured 0.9x on the power-law config).
// A row occupies the sort capacity of its bin (32 << bin slots) inside the tile's shared memory; rows are handed to
// the tile's warps one at a time; rows with more than 512 products were computed into scratch rows beforehand and only
// contribute their nnz.
#ifndef SPADA_TILE_WARPS
#define SPADA_TILE_WARPS 8
#endif
#ifndef SPADA_TILE_CAP
#define SPADA_TILE_CAP 2048
#endif
constexpr int TILE_CAP = SPADA_TILE_CAP;   // sort slots per tile
constexpr int TILE_CUT = TILE_CAP - 512;   // a tile ends where the running slot count crosses a multiple of this
constexpr int TILE_MAX_ROWS = 64;
constexpr int TILE_WARPS = SPADA_TILE_WARPS;
static_assert(TILE_CUT + 512 <= TILE_CAP, "a row that starts below the cut must fit");

__global__ void k_tile_weights(const uint32_t* __restrict__ flops, int64_t m, uint32_t* __restrict__ w) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const uint32_t f = flops[r];
    w[r] = (f >= 1u && f <= 512u) ? (uint32_t)bin_capacity(bin_of(f)) : 0u;
}
// flag[r] = 1 when row r opens a tile (lp = exclusive scan of the weights)
__global__ void k_tile_flags(const int64_t* __restrict__ lp, int64_t m, uint32_t* __restrict__ flag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    flag[r] = (r == 0 || (r % TILE_MAX_ROWS) == 0 || lp[r] / TILE_CUT != lp[r - 1] / TILE_CUT) ? 1u : 0u;
}
__global__ void k_tile_starts(const uint32_t* __restrict__ flag, const int64_t* __restrict__ tidx, int64_t m,
                              uint32_t* __restrict__ tile_start) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    if (flag[r]) tile_start[tidx[r]] = (uint32_t)r;
    if (r == m - 1) tile_start[tidx[m]] = (uint32_t)m;
}

template <typename K>
__global__ void __launch_bounds__(TILE_WARPS * 32)
k_tile_pass(DevCsr a, DevCsr b, int64_t row_begin, int64_t m, const uint32_t* __restrict__ flops,
            const uint32_t* __restrict__ pre_nnz, const int64_t* __restrict__ lp, const int64_t* __restrict__ tidx,
            const uint32_t* __restrict__ tile_start, int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col,
            double* __restrict__ c_val, unsigned long long* tile_state, uint32_t* ticket) {
    constexpr int SBK = 9;   // every key carries (column << 9 | arrival): rows of at most 512 products
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    double* vals = reinterpret_cast<double*>(s_raw + sizeof(K) * TILE_CAP);
    __shared__ uint32_t s_f[TILE_MAX_ROWS], s_roff[TILE_MAX_ROWS], s_nnz[TILE_MAX_ROWS], s_noff[TILE_MAX_ROWS];
    __shared__ uint32_t s_tile;
    __shared__ int s_next;
    __shared__ unsigned long long s_excl;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_tile = atomicAdd(ticket, 1u);   // tile id = start order (see k_fused_light)
        s_next = 0;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if ((int64_t)tile >= tidx[m]) return;
    const int64_t r0 = tile_start[tile];
    const int nr = (int)(tile_start[tile + 1] - (uint32_t)r0);
    if ((int)threadIdx.x < nr) {
        s_f[threadIdx.x] = flops[r0 + threadIdx.x];
        s_roff[threadIdx.x] = (uint32_t)(lp[r0 + threadIdx.x] - lp[r0]);
    }
    __syncthreads();
    // rows go to the warps one at a time
    for (;;) {
        int row = 0;
        if (lane == 0) row = atomicAdd(&s_next, 1);
        row = __shfl_sync(FULL, row, 0);
        if (row >= nr) break;
        const uint32_t f = s_f[row];
        int nnz = 0;
        if (f >= 1u && f <= 512u) {
            const int64_t r = r0 + row;
            K* rk = keys + s_roff[row];
            double* rv = vals + s_roff[row];
            const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
            int seq = 0;
            for (int64_t pb = a_begin; pb < a_end; pb += 32) {
                int bt;
                expand_batch<true, false>(a, b, pb + lane, a_end, lane, seq, bt, [&](int sq, uint32_t c, double av, double bv) {
                    rk[sq] = ((K)c << SBK) | (K)sq;
                    rv[sq] = __dmul_rn(av, bv);
                });
                seq += bt;
            }
            switch (bin_of(f)) {
                case 1: nnz = sort_count<K, 1, SBK>(rk, seq, lane); break;
                case 2: nnz = sort_count<K, 2, SBK>(rk, seq, lane); break;
                case 3: nnz = sort_count<K, 4, SBK>(rk, seq, lane); break;
                case 4: nnz = sort_count<K, 8, SBK>(rk, seq, lane); break;
                default: nnz = sort_count<K, 16, SBK>(rk, seq, lane); break;
            }
        } else if (f > 512u) {
            nnz = (int)pre_nnz[r0 + row];   // computed into a scratch row beforehand
        }
        if (lane == 0) s_nnz[row] = (uint32_t)nnz;
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the tile's row counts (two rows per lane), then the look-back for the tile's base
        const uint32_t n0 = 2 * lane < nr ? s_nnz[2 * lane] : 0u, n1 = 2 * lane + 1 < nr ? s_nnz[2 * lane + 1] : 0u;
        uint32_t x = n0 + n1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (2 * lane < nr) s_noff[2 * lane] = x - n0 - n1;
        if (2 * lane + 1 < nr) s_noff[2 * lane + 1] = x - n1;
        const unsigned long long tile_total = __shfl_sync(FULL, x, 31);
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&tile_state[0], FST_PREFIX | tile_total);
        } else {
            if (lane == 0) atomicExch(&tile_state[tile], FST_AGG | tile_total);
            int64_t pidx = (int64_t)tile - 1;
            while (true) {
                int64_t idx = pidx - lane;
                unsigned long long st;
                do {
                    st = (idx >= 0) ? *((volatile unsigned long long*)&tile_state[idx]) : FST_PREFIX;
                } while (__any_sync(FULL, (st & FST_MASK) == 0));
                unsigned has_prefix = __ballot_sync(FULL, (st & FST_MASK) == FST_PREFIX);
                unsigned long long val = st & ~FST_MASK;
                if (has_prefix) {
                    int first = __ffs(has_prefix) - 1;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(FULL, val, d);
                excl += val;
                if (has_prefix) break;
                pidx -= 32;
            }
            if (lane == 0) atomicExch(&tile_state[tile], FST_PREFIX | (excl + tile_total));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    if ((int)threadIdx.x < nr) {
        const int64_t r = r0 + threadIdx.x;
        const int64_t base = (int64_t)(s_excl + s_noff[threadIdx.x]);
        c_ptr[r] = base;
        if (r == m - 1) c_ptr[m] = base + s_nnz[threadIdx.x];
    }
    for (int row = warp; row < nr; row += TILE_WARPS) {
        const uint32_t f = s_f[row];
        if (f >= 1u && f <= 512u)
            reduce_store<K, SBK>(keys + s_roff[row], vals + s_roff[row], (int)f, lane, (int64_t)(s_excl + s_noff[row]), c_col,
                                 c_val);
    }
}

int tile_pass_cut() { return TILE_CUT; }
size_t tile_pass_bound(int64_t m, uint64_t light_slots) { return (size_t)(light_slots / TILE_CUT + (uint64_t)m / TILE_MAX_ROWS + 4); }

// w / lp / tidx / tile_start / tile_state: workspace (m, m + 1, m + 1, bound + 1, bound + 1 entries); scan_state: the
// look-back state of the two scans
void launch_tile_pass(const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t m, const uint32_t* flops,
                      const uint32_t* pre_nnz, int64_t* c_ptr, int32_t* c_col, double* c_val, uint32_t* w, int64_t* lp,
                      int64_t* tidx, uint32_t* tile_start, uint64_t* tile_state, size_t bound, uint64_t* scan_state,
                      PlanCounters* ctr, cudaStream_t s) {
    if (m <= 0) return;
    const unsigned g = (unsigned)((m + 255) / 256);
    k_tile_weights<<<g, 256, 0, s>>>(flops, m, w);
    launch_scan_u32_i64(w, m, lp, scan_state, ctr, s);
    k_tile_flags<<<g, 256, 0, s>>>(lp, m, w);
    launch_scan_u32_i64(w, m, tidx, scan_state, ctr, s);
    k_tile_starts<<<g, 256, 0, s>>>(w, tidx, m, tile_start);
    cudaMemsetAsync(tile_state, 0, (bound + 1) * sizeof(uint64_t), s);
    uint32_t* ticket = &ctr->scan_ticket;
    cudaMemsetAsync(ticket, 0, sizeof(uint32_t), s);
    const bool narrow = (uint64_t)b.cols < (1ull << 23);
    const size_t smem = (narrow ? 12 : 16) * (size_t)TILE_CAP;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_tile_pass<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * TILE_CAP);
        cudaFuncSetAttribute(k_tile_pass<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * TILE_CAP);
    }
    if (narrow) {
        k_tile_pass<uint32_t><<<(unsigned)bound, TILE_WARPS * 32, smem, s>>>(a, b, row_begin, m, flops, pre_nnz, lp, tidx,
                                                                              tile_start, c_ptr, c_col, c_val,
                                                                              (unsigned long long*)tile_state, ticket);
    } else {
        k_tile_pass<uint64_t><<<(unsigned)bound, TILE_WARPS * 32, smem, s>>>(a, b, row_begin, m, flops, pre_nnz, lp, tidx,
                                                                              tile_start, c_ptr, c_col, c_val,
                                                                              (unsigned long long*)tile_state, ticket);
    }
}

constexpr int fused_rpw(int nmax) { return nmax <= 32 ? 4 : (nmax <= 64 ? 2 : 1); }

template <typename K, int NMAX>
static void fused_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t m, const uint32_t* flops,
                         const uint32_t* pre_nnz, int64_t* c_ptr, int32_t* c_col, double* c_val, uint64_t* tile_state,
                         uint32_t* ticket, cudaStream_t s) {
    constexpr int RPW = fused_rpw(NMAX);
    constexpr int TILE_ROWS = LIGHT_WARPS * RPW;
    size_t smem = (sizeof(K) + sizeof(double)) * NMAX * TILE_ROWS;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_fused_light<K, NMAX, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    size_t tiles = (size_t)((m + TILE_ROWS - 1) / TILE_ROWS);
    cudaMemsetAsync(tile_state, 0, tiles * sizeof(uint64_t), s);
    cudaMemsetAsync(ticket, 0, sizeof(uint32_t), s);
    k_fused_light<K, NMAX, RPW><<<(unsigned)tiles, LIGHT_WARPS * 32, smem, s>>>(
        a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, (unsigned long long*)tile_state, ticket);
}

size_t fused_tile_state_words(int64_t m) { return (size_t)((m + LIGHT_WARPS - 1) / LIGHT_WARPS) + 1; }

template <typename K>
static void fused_dispatch(int max_bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t m,
                           const uint32_t* flops, const uint32_t* pre_nnz, int64_t* c_ptr, int32_t* c_col, double* c_val,
                           uint64_t* tile_state, uint32_t* ticket, cudaStream_t s) {
    switch (max_bin) {
        case 1: fused_launch<K, 32>(a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, tile_state, ticket, s); break;
        case 2: fused_launch<K, 64>(a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, tile_state, ticket, s); break;
        case 3: fused_launch<K, 128>(a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, tile_state, ticket, s); break;
        case 4: fused_launch<K, 256>(a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, tile_state, ticket, s); break;
        default: fused_launch<K, 512>(a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, tile_state, ticket, s); break;
    }
}

// max_bin: the largest warp-per-row bin (1..5) that holds rows; it sizes the shared memory of a tile.
void launch_fused_light(int max_bin, bool tiny_quad, const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t m,
                        const uint32_t* flops, const uint32_t* pre_nnz, int64_t* c_ptr, int32_t* c_col, double* c_val,
                        uint64_t* tile_state, PlanCounters* ctr, cudaStream_t s) {
    if (m <= 0) return;
    if (max_bin < 1) max_bin = 1;
    if (max_bin > 5) max_bin = 5;
    uint32_t* ticket = &ctr->scan_ticket;
    if (max_bin == 1) {
        size_t tiles = (size_t)((m + TINY_TILE - 1) / TINY_TILE);
        cudaMemsetAsync(tile_state, 0, tiles * sizeof(uint64_t), s);
        cudaMemsetAsync(ticket, 0, sizeof(uint32_t), s);
        const bool quad = tiny_quad;   // the window of bin 1: [4, 8] or [1, 32] (engine.cu: window_choice)
        const bool narrow = (uint64_t)b.cols < (1ull << 27);
        if (quad) {
            if (narrow)
                k_fused_tiny4<uint32_t><<<(unsigned)tiles, FUSED_WARPS * 32, 0, s>>>(
                    a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, (unsigned long long*)tile_state, ticket);
            else
                k_fused_tiny4<uint64_t><<<(unsigned)tiles, FUSED_WARPS * 32, 0, s>>>(
                    a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, (unsigned long long*)tile_state, ticket);
        } else if (narrow)
            k_fused_tiny<uint32_t><<<(unsigned)tiles, FUSED_WARPS * 32, 0, s>>>(
                a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, (unsigned long long*)tile_state, ticket);
        else
            k_fused_tiny<uint64_t><<<(unsigned)tiles, FUSED_WARPS * 32, 0, s>>>(
                a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, (unsigned long long*)tile_state, ticket);
        return;
    }
    int sbk = 4 + max_bin;
    bool narrow = (uint64_t)b.cols < (1ull << (32 - sbk));  // strict: a valid key is never the all-ones sentinel
    if (narrow)
        fused_dispatch<uint32_t>(max_bin, a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, tile_state, ticket, s);
    else
        fused_dispatch<uint64_t>(max_bin, a, b, row_begin, m, flops, pre_nnz, c_ptr, c_col, c_val, tile_state, ticket, s);
}

}  // namespace spada
