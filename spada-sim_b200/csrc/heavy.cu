// heavy.cu -- stages 2 and 3 for rows with more than 4096 intermediate products (bin 9): an
// occupancy bitmap over B's column space plus per-word ranks per row, so every product finds
// its slot in the (sorted) output row directly -- no hash probing and no sort.
//
// Reference logic replaced: rows longer than a window are K-tiled into several partial rows
// that the adder trees merge later (scheduler.rs:522-524, merge_task :381-480,
// in_cache_merge_task :820-920; adder_tree.rs:73-83, 145-188).  Here the K tiles are "items"
// (a contiguous slice of the A row worth ~8192 products) spread over the whole grid -- a row
// with half a million products no longer serialises on one CTA -- and the merge happens in
// place: bitmap bit j set <=> some product lands on column j; rank(j) = number of set bits
// below j = position of C[i,j] inside the row.
//
// Summation order inside one C[i,j] is not fixed on this path (partial products arrive through
// atomics) -- like the reference, whose merge order depends on HashMap iteration order
// (scheduler.rs:386, 396, 827); results agree with the oracle to within a few ulp (tested at
// rel 1e-12), structure is exact.
//
// Workspace: one uint2 {bits, rank} per 32 columns of B per heavy row of the current wave.
#include <cstdlib>

#include "common.cuh"

namespace spada {

constexpr int HEAVY_THREADS = 256;
constexpr int HEAVY_WARPS = HEAVY_THREADS / 32;
constexpr uint32_t HEAVY_ITEM_PRODUCTS = 8192;
constexpr int TMA_MIN_LEN = 64;   // B rows of at least this many elements are staged by TMA bulk copies
constexpr int TMA_STAGE = 256;    // elements per staging buffer (1 KB of column ids + 2 KB of values per warp)

// Streams one long B row through the warp's shared-memory staging buffer with cp.async.bulk:
// lane 0 arms the warp's mbarrier with the byte count and issues the bulk copies (16-byte aligned
// windows of the row, over-fetching at most 3 elements on either side), all lanes wait on the
// barrier phase and consume the window.  use(col, b_val) is called per element of the row.
template <bool NUMERIC, typename U>
__device__ __forceinline__ void stream_row_tma(const DevCsr& b, int64_t bs, int len, int lane, uint32_t* st_col,
                                               double* st_val, uint64_t* bar, uint32_t& phase, U&& use) {
    const int64_t row_end = bs + len;
    int64_t e0 = bs;
    while (e0 < row_end) {
        const int64_t start_al = e0 & ~(int64_t)3;
        int64_t end_al = (row_end + 3) & ~(int64_t)3;
        if (end_al - start_al > TMA_STAGE) end_al = start_al + TMA_STAGE;
        const int64_t stop = end_al < row_end ? end_al : row_end;
        if (end_al <= b.nnz) {
            const uint32_t n_al = (uint32_t)(end_al - start_al);
            if (lane == 0) {
                mbar_expect_tx(bar, n_al * (NUMERIC ? 12u : 4u));
                tma_load_1d(st_col, b.col + start_al, n_al * 4u, bar);
                if (NUMERIC) tma_load_1d(st_val, b.val + start_al, n_al * 8u, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            for (int64_t e = e0 + lane; e < stop; e += 32) use(st_col[e - start_al], NUMERIC ? st_val[e - start_al] : 0.0);
            __syncwarp();
        } else {
            // the aligned window would run past the end of B's arrays: plain loads for the tail
            for (int64_t e = e0 + lane; e < stop; e += 32)
                use((uint32_t)ldg_i32(b.col + e), NUMERIC ? ldg_f64(b.val + e) : 0.0);
        }
        e0 = stop;
    }
}

__host__ __device__ inline uint32_t heavy_words(int64_t b_cols) { return (uint32_t)((b_cols + 31) / 32); }

// items of one heavy row: n = min(ceil(p / 8192), ceil(len / 32)) slices of `chunk` A entries each
__device__ __forceinline__ void item_shape(uint32_t p, int64_t a_len, uint32_t& n_items, int64_t& chunk) {
    uint32_t by_p = (p + HEAVY_ITEM_PRODUCTS - 1) / HEAVY_ITEM_PRODUCTS;
    int64_t by_len = (a_len + 31) / 32;
    n_items = (uint32_t)(by_p < by_len ? by_p : by_len);
    if (n_items < 1) n_items = 1;
    chunk = ((a_len + n_items - 1) / n_items + 31) & ~(int64_t)31;
    n_items = (uint32_t)((a_len + chunk - 1) / chunk);
    if (n_items < 1) n_items = 1;
}

__global__ void k_heavy_plan(DevCsr a, int64_t row_begin, const uint32_t* __restrict__ rows_list, uint32_t n_rows,
                             const uint32_t* __restrict__ flops, uint32_t* __restrict__ items_per_row) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    uint32_t r = rows_list ? rows_list[i] : i;
    int64_t len = a.ptr[row_begin + r + 1] - a.ptr[row_begin + r];
    uint32_t n;
    int64_t chunk;
    item_shape(flops[r], len, n, chunk);
    items_per_row[i] = n;
}

__global__ void k_heavy_fill(const int64_t* __restrict__ item_off, uint32_t n_rows, uint32_t* __restrict__ item_row) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    for (int64_t t = item_off[i]; t < item_off[i + 1]; ++t) item_row[t] = i;
}

// the slice of A row entries an item covers
struct ItemRange {
    uint32_t hrow;  // index into the heavy list
    uint32_t row;   // row of A (relative to row_begin)
    int64_t a0, a1;
};
__device__ __forceinline__ ItemRange item_range(const DevCsr& a, int64_t row_begin, const uint32_t* rows_list,
                                                const uint32_t* flops, const int64_t* item_off,
                                                const uint32_t* item_row, int64_t it) {
    ItemRange R;
    R.hrow = item_row[it];
    R.row = rows_list ? rows_list[R.hrow] : R.hrow;
    int64_t s = a.ptr[row_begin + R.row], e = a.ptr[row_begin + R.row + 1];
    uint32_t n;
    int64_t chunk;
    item_shape(flops[R.row], e - s, n, chunk);
    int64_t j = it - item_off[R.hrow];
    R.a0 = s + j * chunk;
    R.a1 = R.a0 + chunk < e ? R.a0 + chunk : e;
    return R;
}

// A entries one warp takes per turn inside an item: the whole batch of 32 when the item is long enough to keep all
// eight warps busy, fewer otherwise.  Items are cut by product count, so an item whose B rows are long (R-MAT hubs,
// cari) has only 32 A entries: with 32 per warp seven warps of eight sat idle (ncu on R-MAT: warps active 9-15 %,
// issue active 3 %); the products of a few entries are still dealt over all 32 lanes.
__device__ __forceinline__ int item_sub_batch(const ItemRange& R) {
    const int64_t per_warp = (R.a1 - R.a0 + HEAVY_WARPS - 1) / HEAVY_WARPS;
    return per_warp >= 32 ? 32 : (per_warp < 1 ? 1 : (int)per_warp);
}

// ---- bits: every product sets its column's bit in the row's bitmap ------------------------------
__global__ void __launch_bounds__(HEAVY_THREADS)
k_heavy_bits(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ rows_list,
             const uint32_t* __restrict__ flops, const int64_t* __restrict__ item_off,
             const uint32_t* __restrict__ item_row, uint32_t wave_lo, uint32_t wave_hi, uint2* ws, uint32_t words) {
    __shared__ __align__(16) uint32_t s_col[HEAVY_WARPS][TMA_STAGE];
    __shared__ uint64_t s_bar[HEAVY_WARPS];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (lane == 0) mbar_init(&s_bar[warp], 1);
    __syncwarp();
    uint32_t phase = 0;
    const bool tma_ok = (reinterpret_cast<uintptr_t>(b.col) & 15) == 0;
    const int long_len = tma_ok ? TMA_MIN_LEN : EXPAND_BIG_LEN + 1;
    const int64_t it_end = item_off[wave_hi];
    for (int64_t it = item_off[wave_lo] + blockIdx.x; it < it_end; it += gridDim.x) {
        ItemRange R = item_range(a, row_begin, rows_list, flops, item_off, item_row, it);
        uint2* w = ws + (size_t)(R.hrow - wave_lo) * words;
        // Test the word with a plain load first: the load brings the sector into L2 with full
        // memory-level parallelism (an atomic that misses L2 is served far more slowly), and bits
        // that are already set need no atomic at all (the OR is idempotent).
        const int sub = item_sub_batch(R);
        for (int64_t pb = R.a0 + (int64_t)warp * sub; pb < R.a1; pb += (int64_t)HEAVY_WARPS * sub) {
            const int64_t pe = pb + sub < R.a1 ? pb + sub : R.a1;   // this warp's entries of the turn
            int bt;
            auto set_bit = [&](uint32_t c) {
                uint32_t bit = 1u << (c & 31);
                if (!(__ldcg(&w[c >> 5].x) & bit)) atomicOr(&w[c >> 5].x, bit);
            };
            expand_batch_long<false, true, true>(
                a, b, pb + lane, pe, lane, 0, bt, long_len, [&](int, uint32_t c, double, double) { set_bit(c); },
                [&](int64_t bsj, int lj, double) {
                    if (tma_ok)
                        stream_row_tma<false>(b, bsj, lj, lane, s_col[warp], nullptr, &s_bar[warp], phase,
                                              [&](uint32_t c, double) { set_bit(c); });
                    else
                        for (int t = lane; t < lj; t += 32) set_bit((uint32_t)ldg_i32(b.col + bsj + t));
                });
        }
    }
}

// ---- rank: per row, exclusive prefix of the words' popcounts; total = nnz of the row -----------
__global__ void __launch_bounds__(HEAVY_THREADS)
k_heavy_rank(const uint32_t* __restrict__ rows_list, uint32_t wave_lo, uint2* ws, uint32_t words,
             uint32_t* __restrict__ row_nnz) {
    __shared__ uint32_t s_wtot[HEAVY_WARPS];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t hrow = wave_lo + blockIdx.x;
    uint2* w = ws + (size_t)blockIdx.x * words;
    uint32_t run = 0;
    for (uint32_t wb = 0; wb < words; wb += HEAVY_THREADS) {
        uint32_t i = wb + threadIdx.x;
        uint32_t bits = (i < words) ? __ldcg(&w[i].x) : 0u;
        uint32_t pc = __popc(bits);
        uint32_t x = pc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_wtot[warp] = x;
        __syncthreads();
        uint32_t wbase = run, all = 0;
#pragma unroll
        for (int q = 0; q < HEAVY_WARPS; ++q) {
            uint32_t t = s_wtot[q];
            if (q < warp) wbase += t;
            all += t;
        }
        if (bits) w[i].y = wbase + x - pc;
        run += all;
        __syncthreads();
    }
    if (row_nnz && threadIdx.x == 0) row_nnz[rows_list ? rows_list[hrow] : hrow] = run;
}

// ---- emit: column ids of the row from the bitmap, values zeroed for the accumulation -----------
__global__ void __launch_bounds__(HEAVY_THREADS)
k_heavy_emit(const uint32_t* __restrict__ rows_list, uint32_t wave_lo, const uint2* ws, uint32_t words,
             const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
             const uint32_t* __restrict__ row_nnz) {
    const uint32_t hrow = wave_lo + blockIdx.x;
    const uint32_t r = rows_list ? rows_list[hrow] : hrow;
    const uint2* w = ws + (size_t)blockIdx.x * words;
    // row_nnz given: c_ptr addresses scratch rows of capacity >= nnz (one-shot mode), only nnz slots are used
    const int64_t cbase = c_ptr[r], z = row_nnz ? (int64_t)row_nnz[r] : c_ptr[r + 1] - cbase;
    for (int64_t i = threadIdx.x; i < z; i += HEAVY_THREADS) c_val[cbase + i] = 0.0;
    for (uint32_t i = threadIdx.x; i < words; i += HEAVY_THREADS) {
        uint2 e = __ldcg(&w[i]);
        uint32_t bb = e.x;
        int64_t o = cbase + e.y;
        while (bb) {
            int bit = __ffs(bb) - 1;
            bb &= bb - 1;
            c_col[o++] = (int32_t)(i * 32u + bit);
        }
    }
}

// ---- accumulate: products -> slots ---------------------------------------------------------------
__global__ void __launch_bounds__(HEAVY_THREADS)
k_heavy_accum(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ rows_list,
              const uint32_t* __restrict__ flops, const int64_t* __restrict__ item_off,
              const uint32_t* __restrict__ item_row, uint32_t wave_lo, uint32_t wave_hi, const uint2* ws,
              uint32_t words, const int64_t* __restrict__ c_ptr, double* __restrict__ c_val) {
    __shared__ __align__(16) uint32_t s_col[HEAVY_WARPS][TMA_STAGE];
    __shared__ __align__(16) double s_val[HEAVY_WARPS][TMA_STAGE];
    __shared__ uint64_t s_bar[HEAVY_WARPS];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (lane == 0) mbar_init(&s_bar[warp], 1);
    __syncwarp();
    uint32_t phase = 0;
    const bool tma_ok = ((reinterpret_cast<uintptr_t>(b.col) | reinterpret_cast<uintptr_t>(b.val)) & 15) == 0;
    const int long_len = tma_ok ? TMA_MIN_LEN : EXPAND_BIG_LEN + 1;
    const int64_t it_end = item_off[wave_hi];
    for (int64_t it = item_off[wave_lo] + blockIdx.x; it < it_end; it += gridDim.x) {
        ItemRange R = item_range(a, row_begin, rows_list, flops, item_off, item_row, it);
        const uint2* w = ws + (size_t)(R.hrow - wave_lo) * words;
        double* out = c_val + c_ptr[R.row];
        // The add of one step is issued one step late: its slot was prefetched into L2 a full
        // memory round trip earlier, so the atomic hits L2 instead of waiting on an HBM fill.
        double pend_v = 0.0;
        uint32_t pend_pos = 0;
        bool pend = false;
        const int sub = item_sub_batch(R);
        for (int64_t pb = R.a0 + (int64_t)warp * sub; pb < R.a1; pb += (int64_t)HEAVY_WARPS * sub) {
            const int64_t pe = pb + sub < R.a1 ? pb + sub : R.a1;   // this warp's entries of the turn
            int bt;
            auto add = [&](uint32_t c, double prod) {
                uint2 e = __ldcg(&w[c >> 5]);
                uint32_t pos = e.y + __popc(e.x & ((1u << (c & 31)) - 1u));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(out + pos));
                if (pend) atomicAdd(out + pend_pos, pend_v);
                pend = true;
                pend_pos = pos;
                pend_v = prod;
            };
            expand_batch_long<true, true, true>(
                a, b, pb + lane, pe, lane, 0, bt, long_len,
                [&](int, uint32_t c, double av, double bv) { add(c, __dmul_rn(av, bv)); },
                [&](int64_t bsj, int lj, double aj) {
                    if (tma_ok)
                        stream_row_tma<true>(b, bsj, lj, lane, s_col[warp], s_val[warp], &s_bar[warp], phase,
                                             [&](uint32_t c, double bv) { add(c, __dmul_rn(aj, bv)); });
                    else
                        for (int t = lane; t < lj; t += 32)
                            add((uint32_t)ldg_i32(b.col + bsj + t), __dmul_rn(aj, ldg_f64(b.val + bsj + t)));
                });
        }
        if (pend) atomicAdd(out + pend_pos, pend_v);
    }
}

// ---- host side -------------------------------------------------------------------------------------
HeavyPlan heavy_plan_sizes(uint32_t n_rows, uint64_t products, int64_t b_cols, size_t ws_budget_bytes) {
    HeavyPlan P;
    P.words = heavy_words(b_cols);
    P.max_items = products / HEAVY_ITEM_PRODUCTS + (uint64_t)n_rows + 1;
    size_t per_row = (size_t)P.words * sizeof(uint2);
    size_t rows_fit = per_row ? ws_budget_bytes / per_row : n_rows;
    if (rows_fit < 1) rows_fit = 1;
    P.wave_rows = (uint32_t)(rows_fit < n_rows ? rows_fit : n_rows);
    P.n_waves = (n_rows + P.wave_rows - 1) / P.wave_rows;
    P.ws_words = (size_t)P.wave_rows * P.words;
    return P;
}

void launch_heavy_items(const DevCsr& a, int64_t row_begin, const uint32_t* rows_list, uint32_t n_rows,
                        const uint32_t* flops, uint32_t* items_per_row, int64_t* item_off, uint32_t* item_row,
                        uint64_t* tile_state, PlanCounters* ctr, cudaStream_t s) {
    unsigned g = (n_rows + 255) / 256;
    k_heavy_plan<<<g, 256, 0, s>>>(a, row_begin, rows_list, n_rows, flops, items_per_row);
    launch_scan_u32_i64(items_per_row, n_rows, item_off, tile_state, ctr, s);
    k_heavy_fill<<<g, 256, 0, s>>>(item_off, n_rows, item_row);
}

// resident CTAs of the item kernels: few enough that the bitmaps of the rows in flight stay in L2
static int persistent_grid(int sm_count) {
    static int per_sm = 0;
    if (!per_sm) {
        const char* e = getenv("SPADA_B200_HEAVY_CTAS_PER_SM");
        per_sm = e ? atoi(e) : 8;
        if (per_sm < 1 || per_sm > 8) per_sm = 8;
    }
    return sm_count * per_sm;
}

void launch_heavy_bits(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list,
                       const uint32_t* flops, const int64_t* item_off, const uint32_t* item_row, uint32_t wave_lo,
                       uint32_t wave_hi, uint2* ws, const HeavyPlan& P, int sm_count, cudaStream_t s) {
    k_heavy_bits<<<persistent_grid(sm_count), HEAVY_THREADS, 0, s>>>(a, b, row_begin, rows_list, flops, item_off,
                                                                    item_row, wave_lo, wave_hi, ws, P.words);
}
void launch_heavy_rank(const uint32_t* rows_list, uint32_t wave_lo, uint32_t wave_hi, uint2* ws, const HeavyPlan& P,
                       uint32_t* row_nnz, cudaStream_t s) {
    k_heavy_rank<<<wave_hi - wave_lo, HEAVY_THREADS, 0, s>>>(rows_list, wave_lo, ws, P.words, row_nnz);
}
void launch_heavy_emit(const uint32_t* rows_list, uint32_t wave_lo, uint32_t wave_hi, const uint2* ws,
                       const HeavyPlan& P, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                       const uint32_t* row_nnz) {
    k_heavy_emit<<<wave_hi - wave_lo, HEAVY_THREADS, 0, s>>>(rows_list, wave_lo, ws, P.words, c_ptr, c_col, c_val, row_nnz);
}
void launch_heavy_accum(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list,
                        const uint32_t* flops, const int64_t* item_off, const uint32_t* item_row, uint32_t wave_lo,
                        uint32_t wave_hi, const uint2* ws, const HeavyPlan& P, const int64_t* c_ptr, double* c_val,
                        int sm_count, cudaStream_t s) {
    k_heavy_accum<<<persistent_grid(sm_count), HEAVY_THREADS, 0, s>>>(a, b, row_begin, rows_list, flops, item_off,
                                                                     item_row, wave_lo, wave_hi, ws, P.words, c_ptr,
                                                                     c_val);
}

}  // namespace spada
