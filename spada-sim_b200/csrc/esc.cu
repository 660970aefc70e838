// esc.cu -- stages 2 and 3 in one pass for rows with at most 4096 intermediate products (bins 1..8):
// expand the row's products into shared memory, sort them by (column, arrival order),
// sum equal columns left to right, store the canonical row and record its nnz.
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

// ---- launchers --------------------------------------------------------------------------------
// bins 1..5: warp per row (ESC_WARPS rows per CTA); bins 6..8: CTA per row (esc_cta_bitonic.cu)
int esc_grid(int bin, uint32_t rows) {
    if (bin <= 5) return (int)((rows + ESC_WARPS - 1) / ESC_WARPS);
    return (int)rows;
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

}  // namespace spada
