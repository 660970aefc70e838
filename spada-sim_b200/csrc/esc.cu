// esc.cu -- stages 2 and 3 for rows with at most 4096 intermediate products (bins 1..8):
// expand the row's products into shared memory, sort them by (column, arrival order),
// sum equal columns left to right, store the canonical row.
//
// This is the GPU restatement of one PE pass of the reference:
//   MultiplierArray::multiply   simulator.rs:86-111   one rounded f64 multiply per product
//   SortingNetwork::pop_elements simulator.rs:143-171  stable sort by column
//   MergeTree::pop_elements     simulator.rs:199-230   equal columns summed left to right
//   write_psums                 simulator.rs:955-983   append to the output row
// The sort key is (column << log2 N | arrival index), arrival index ascending in k then in
// B's stored order, so the summation order of every C[i,j] is the pure ascending-k order the
// CPU oracle fixes (oracle/spgemm_oracle.c): values come out bit-identical, not just within
// 1e-12.  Products use __dmul_rn / __dadd_rn: never contracted into an FMA.
//
// Window shape (scheduler.rs:729-753): bins 1..5 give each A row one warp (32 lanes x E keys
// per lane, E = N/32, bitonic network in registers), four rows share a CTA; bins 6..8 give each
// A row a whole CTA (esc_cta_bitonic.cu).
#include "common.cuh"
#include "sort.cuh"

namespace spada {

#ifndef SPADA_ESC_WARPS
#define SPADA_ESC_WARPS 4
#endif
constexpr int ESC_WARPS = SPADA_ESC_WARPS;  // rows per CTA in the warp-per-row bins

// =============================================================================================
// warp-per-row kernels, N = 32 * E products at most
// =============================================================================================
template <int N>
__global__ void __launch_bounds__(ESC_WARPS * 32)
k_esc_symbolic_warp(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                    uint32_t* __restrict__ row_nnz) {
    constexpr int E = N / 32;
    __shared__ __align__(16) uint32_t s_keys[ESC_WARPS][N];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * ESC_WARPS + warp;
    if (w >= rows) return;
    const uint32_t r = perm ? perm[w] : w;
    uint32_t* keys = s_keys[warp];
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    int seq = 0;
    for (int64_t pb = a_begin; pb < a_end; pb += 32) {
        int bt;
        expand_batch<false, false>(a, b, pb + lane, a_end, lane, seq, bt,
                                   [&](int sq, uint32_t c, double, double) { keys[sq] = c; });
        seq += bt;
    }
    for (int t = seq + lane; t < N; t += 32) keys[t] = 0xffffffffu;
    __syncwarp();
    uint32_t x[E];
    load_blocked<uint32_t, E>(x, keys, lane);
    warp_sort<uint32_t, E>(x, lane);
    uint32_t prev = __shfl_up_sync(FULL, x[E - 1], 1);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        bool first = (lane == 0 && i == 0);
        uint32_t pv = (i == 0) ? prev : x[i - 1];
        if (x[i] != 0xffffffffu && (first || x[i] != pv)) ++cnt;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
    if (lane == 0) row_nnz[r] = (uint32_t)cnt;
}

template <typename K, int N>
__global__ void __launch_bounds__(ESC_WARPS * 32)
k_esc_numeric_warp(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                   const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
                   uint32_t* __restrict__ row_nnz_out) {
    constexpr int E = N / 32;
    constexpr int SB = Log2<N>::v;
    __shared__ __align__(16) K s_keys[ESC_WARPS][N];
    __shared__ __align__(16) double s_vals[ESC_WARPS][N];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * ESC_WARPS + warp;
    if (w >= rows) return;
    const uint32_t r = perm ? perm[w] : w;
    K* keys = s_keys[warp];
    double* vals = s_vals[warp];
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    int seq = 0;
    for (int64_t pb = a_begin; pb < a_end; pb += 32) {
        int bt;
        expand_batch<true, false>(a, b, pb + lane, a_end, lane, seq, bt, [&](int sq, uint32_t c, double av, double bv) {
            keys[sq] = ((K)c << SB) | (K)sq;
            vals[sq] = __dmul_rn(av, bv);
        });
        seq += bt;
    }
    const int p = seq;
    for (int t = p + lane; t < N; t += 32) keys[t] = KeyTraits<K>::sentinel;
    __syncwarp();
    K x[E];
    load_blocked<K, E>(x, keys, lane);
    warp_sort<K, E>(x, lane);
    __syncwarp();
    store_blocked<K, E>(x, keys, lane);
    __syncwarp();
    // segmented left-to-right sums over equal columns, compacted to the row's slot in C
    const int64_t cbase = c_ptr[r];
    int out_base = 0;
    for (int base = 0; base < p; base += 32) {
        int i = base + lane;
        bool valid = i < p;
        K ki = valid ? keys[i] : KeyTraits<K>::sentinel;
        uint32_t col = (uint32_t)(ki >> SB);
        bool head = valid && (i == 0 || (uint32_t)(keys[i - 1] >> SB) != col);
        unsigned hm = __ballot_sync(FULL, head);
        if (head) {
            double sum = vals[(int)(ki & (K)(N - 1))];
            for (int j = i + 1; j < p; ++j) {
                K kj = keys[j];
                if ((uint32_t)(kj >> SB) != col) break;
                sum = __dadd_rn(sum, vals[(int)(kj & (K)(N - 1))]);
            }
            int o = out_base + __popc(hm & ((1u << lane) - 1u));
            st_out(c_col + (cbase + o), (int32_t)col);
            st_out(c_val + (cbase + o), sum);
        }
        out_base += __popc(hm);
    }
    if (row_nnz_out && lane == 0) row_nnz_out[r] = (uint32_t)out_base;
}

// ---- kept-keys variants (two-phase mode) ----------------------------------------------------------
// Symbolic sorts the same (column << log2 N | arrival) keys numeric needs, counts the distinct
// columns and leaves the sorted keys in HBM (4 or 8 B per product); numeric then only expands the
// products' values, reloads the keys with coalesced loads and reduces -- the row is sorted once.
template <typename K, int N>
__global__ void __launch_bounds__(ESC_WARPS * 32)
k_esc_symbolic_keep_warp(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                         uint32_t* __restrict__ row_nnz, const int64_t* __restrict__ prod_ptr, K* __restrict__ kstore) {
    constexpr int E = N / 32;
    constexpr int SB = Log2<N>::v;
    __shared__ __align__(16) K s_keys[ESC_WARPS][N];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * ESC_WARPS + warp;
    if (w >= rows) return;
    const uint32_t r = perm ? perm[w] : w;
    K* keys = s_keys[warp];
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    int seq = 0;
    for (int64_t pb = a_begin; pb < a_end; pb += 32) {
        int bt;
        expand_batch<false, false>(a, b, pb + lane, a_end, lane, seq, bt,
                                   [&](int sq, uint32_t c, double, double) { keys[sq] = ((K)c << SB) | (K)sq; });
        seq += bt;
    }
    const int p = seq;
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
        if (lane * E + i < p && (first || (uint32_t)(x[i] >> SB) != (uint32_t)(pv >> SB))) ++cnt;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
    if (lane == 0) row_nnz[r] = (uint32_t)cnt;
    __syncwarp();
    K* dst = kstore + prod_ptr[r];
    for (int t = lane; t < p; t += 32) dst[t] = keys[t];
}

template <typename K, int N>
__global__ void __launch_bounds__(ESC_WARPS * 32)
k_esc_numeric_presorted_warp(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                             const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
                             const int64_t* __restrict__ prod_ptr, const K* __restrict__ kstore) {
    constexpr int SB = Log2<N>::v;
    __shared__ __align__(16) double s_vals[ESC_WARPS][N];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * ESC_WARPS + warp;
    if (w >= rows) return;
    const uint32_t r = perm ? perm[w] : w;
    double* vals = s_vals[warp];
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    int seq = 0;
    for (int64_t pb = a_begin; pb < a_end; pb += 32) {
        int bt;
        expand_batch<true, false, false>(a, b, pb + lane, a_end, lane, seq, bt,
                                         [&](int sq, uint32_t, double av, double bv) { vals[sq] = __dmul_rn(av, bv); });
        seq += bt;
    }
    const int p = seq;
    __syncwarp();
    const K* keys = kstore + prod_ptr[r];
    const int64_t cbase = c_ptr[r];
    int out_base = 0;
    uint32_t prev_last = 0;
    for (int base = 0; base < p; base += 32) {
        const int i = base + lane;
        const bool valid = i < p;
        const K ki = valid ? keys[i] : KeyTraits<K>::sentinel;
        const uint32_t col = (uint32_t)(ki >> SB);
        uint32_t col_prev = __shfl_up_sync(FULL, col, 1);
        if (lane == 0) col_prev = prev_last;
        const bool head = valid && (i == 0 || col_prev != col);
        const unsigned hm = __ballot_sync(FULL, head);
        prev_last = __shfl_sync(FULL, col, 31);
        if (head) {
            double sum = vals[(int)(ki & (K)(N - 1))];
            for (int j = i + 1; j < p; ++j) {
                const K kj = keys[j];
                if ((uint32_t)(kj >> SB) != col) break;
                sum = __dadd_rn(sum, vals[(int)(kj & (K)(N - 1))]);
            }
            const int o = out_base + __popc(hm & ((1u << lane) - 1u));
            st_out(c_col + (cbase + o), (int32_t)col);
            st_out(c_val + (cbase + o), sum);
        }
        out_base += __popc(hm);
    }
}

// ---- launchers --------------------------------------------------------------------------------
// bins 1..5: warp per row (ESC_WARPS rows per CTA); bins 6..8: CTA per row (esc_cta_bitonic.cu)
int esc_grid(int bin, uint32_t rows) {
    if (bin <= 5) return (int)((rows + ESC_WARPS - 1) / ESC_WARPS);
    return (int)rows;
}

void setup_kernel_attributes() {}

void launch_esc_symbolic(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                         uint32_t rows, uint32_t* row_nnz, cudaStream_t s) {
    if (rows == 0) return;
    unsigned g = (unsigned)esc_grid(bin, rows);
    switch (bin) {
        case 1: k_esc_symbolic_warp<32><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        case 2: k_esc_symbolic_warp<64><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        case 3: k_esc_symbolic_warp<128><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        case 4: k_esc_symbolic_warp<256><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        case 5: k_esc_symbolic_warp<512><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        default: launch_bitonic_cta_symbolic(bin, a, b, row_begin, perm, rows, row_nnz, s); break;
    }
}

template <typename K>
static void numeric_warp_dispatch(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                  uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                  uint32_t* nnz_out) {
    unsigned g = (unsigned)esc_grid(bin, rows);
    switch (bin) {
        case 1: k_esc_numeric_warp<K, 32><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, nnz_out); break;
        case 2: k_esc_numeric_warp<K, 64><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, nnz_out); break;
        case 3: k_esc_numeric_warp<K, 128><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, nnz_out); break;
        case 4: k_esc_numeric_warp<K, 256><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, nnz_out); break;
        default: k_esc_numeric_warp<K, 512><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, nnz_out); break;
    }
}

void launch_esc_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                        uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                        uint32_t* row_nnz_out) {
    if (rows == 0) return;
    if (bin <= 5) {
        // 32-bit keys whenever (column << log2 N | arrival) fits: b.cols <= 2^(32 - log2 N)
        int sb = 4 + bin;  // log2(N): bin 1 -> 32 = 2^5
        bool narrow = (uint64_t)b.cols <= (1ull << (32 - sb));
        if (narrow)
            numeric_warp_dispatch<uint32_t>(bin, a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, row_nnz_out);
        else
            numeric_warp_dispatch<uint64_t>(bin, a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, row_nnz_out);
        return;
    }
    launch_bitonic_cta_numeric(bin, a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, row_nnz_out);
}

bool esc_needs_wide_keys(int bin, int64_t b_cols) {
    int sb = 4 + bin;  // log2 of the bin capacity
    return !((uint64_t)b_cols <= (1ull << (32 - sb)));
}

template <typename K>
static void sym_keep_dispatch(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                              uint32_t rows, uint32_t* row_nnz, const int64_t* prod_ptr, void* kstore, cudaStream_t s) {
    unsigned g = (unsigned)esc_grid(bin, rows);
    K* ks = reinterpret_cast<K*>(kstore);
    switch (bin) {
        case 1: k_esc_symbolic_keep_warp<K, 32><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz, prod_ptr, ks); break;
        case 2: k_esc_symbolic_keep_warp<K, 64><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz, prod_ptr, ks); break;
        case 3: k_esc_symbolic_keep_warp<K, 128><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz, prod_ptr, ks); break;
        case 4: k_esc_symbolic_keep_warp<K, 256><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz, prod_ptr, ks); break;
        default: k_esc_symbolic_keep_warp<K, 512><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, row_nnz, prod_ptr, ks); break;
    }
}
template <typename K>
static void num_presorted_dispatch(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                   uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val,
                                   const int64_t* prod_ptr, const void* kstore, cudaStream_t s) {
    unsigned g = (unsigned)esc_grid(bin, rows);
    const K* ks = reinterpret_cast<const K*>(kstore);
    switch (bin) {
        case 1: k_esc_numeric_presorted_warp<K, 32><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, ks); break;
        case 2: k_esc_numeric_presorted_warp<K, 64><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, ks); break;
        case 3: k_esc_numeric_presorted_warp<K, 128><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, ks); break;
        case 4: k_esc_numeric_presorted_warp<K, 256><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, ks); break;
        default: k_esc_numeric_presorted_warp<K, 512><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, ks); break;
    }
}

void launch_esc_symbolic_keep(int bin, bool wide, const DevCsr& a, const DevCsr& b, int64_t row_begin,
                              const uint32_t* perm, uint32_t rows, uint32_t* row_nnz, const int64_t* prod_ptr,
                              void* kstore, cudaStream_t s) {
    if (rows == 0) return;
    if (wide) sym_keep_dispatch<uint64_t>(bin, a, b, row_begin, perm, rows, row_nnz, prod_ptr, kstore, s);
    else sym_keep_dispatch<uint32_t>(bin, a, b, row_begin, perm, rows, row_nnz, prod_ptr, kstore, s);
}
void launch_esc_numeric_presorted(int bin, bool wide, const DevCsr& a, const DevCsr& b, int64_t row_begin,
                                  const uint32_t* perm, uint32_t rows, const int64_t* c_ptr, int32_t* c_col,
                                  double* c_val, const int64_t* prod_ptr, const void* kstore, cudaStream_t s) {
    if (rows == 0) return;
    if (wide) num_presorted_dispatch<uint64_t>(bin, a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, kstore, s);
    else num_presorted_dispatch<uint32_t>(bin, a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, kstore, s);
}

}  // namespace spada
