// esc_cta_bitonic.cu -- CTA-per-row ESC kernels (bins 6..8: 1024 / 2048 / 4096 products) with the
// hybrid bitonic sort: every warp sorts its chunk in registers, chunks are merged through shared
// memory.  (A stable 4-bit LSD radix sort in shared memory was tried for these bins and for an
// 8192 bin: 30+ barriers per row made it 1.3-1.5x slower than this network at these sizes.)
#include "cta_common.cuh"

namespace spada {

// SPLIT: column ids one bit wider than the 32-bit key holds -- sorted without their top bit, then split by it
// (cta_split_top)
template <typename K, int N, bool SPLIT = false>
// resident CTAs per SM the shared memory allows (12 N bytes + the staging): 4 / 5-6 / 8 -- registers capped to match
__global__ void __launch_bounds__(ESC_CTA_THREADS, N == 4096 ? 4 : N == 2048 ? (sizeof(K) == 4 ? 6 : 5) : 8)
k_bitonic_numeric_cta(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                      const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
                      uint32_t* __restrict__ row_nnz_out) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    double* vals = reinterpret_cast<double*>(s_raw + sizeof(K) * N);
    __shared__ CtaStage st;
    __shared__ uint32_t top[SPLIT ? N / 32 + 1 : 1];
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    if constexpr (SPLIT)
        for (int t = threadIdx.x; t < N / 32 + 1; t += ESC_CTA_THREADS) top[t] = 0u;   // the expansion syncs before it writes
    const int p = bitonic_cta_expand<K, N, SPLIT>(a, b, a_begin, a_end, keys, vals, st, top);
    for (int t = p + threadIdx.x; t < N; t += ESC_CTA_THREADS) keys[KeySlot<K, N>::at(t)] = KeyTraits<K>::sentinel;
    __syncthreads();
    bitonic_cta_sort<K, N>(keys);
    int n0 = 0x7fffffff;
    if constexpr (SPLIT) n0 = cta_split_top<N>(keys, p, top, st);
    const int total = cta_reduce_store<K, N>(keys, vals, p, c_ptr[r], c_col, c_val, st, n0);
    if (row_nnz_out && threadIdx.x == 0) row_nnz_out[r] = (uint32_t)total;
}

template <typename K, int N, bool SPLIT = false>
static void bitonic_numeric_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                   uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                   uint32_t* nnz_out) {
    size_t smem = (sizeof(K) + sizeof(double)) * N;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_bitonic_numeric_cta<K, N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    k_bitonic_numeric_cta<K, N, SPLIT><<<rows, ESC_CTA_THREADS, smem, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val,
                                                                    nnz_out);
}

void launch_bitonic_cta_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                uint32_t* nnz_out) {
    if (rows == 0) return;
    int sb = 4 + bin;
    bool narrow = (uint64_t)b.cols <= (1ull << (32 - sb));
    if (!narrow && (uint64_t)b.cols <= (1ull << (33 - sb))) {   // one bit short: sorted without the top bit, then split by it
        switch (bin) {
            case 6: bitonic_numeric_launch<uint32_t, 1024, true>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break;
            case 7: bitonic_numeric_launch<uint32_t, 2048, true>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break;
            default: bitonic_numeric_launch<uint32_t, 4096, true>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break;
        }
        return;
    }
#define BITONIC_CASE(K)                                                                                      \
    switch (bin) {                                                                                            \
        case 6: bitonic_numeric_launch<K, 1024>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break;  \
        case 7: bitonic_numeric_launch<K, 2048>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break;  \
        default: bitonic_numeric_launch<K, 4096>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break; \
    }
    if (narrow) { BITONIC_CASE(uint32_t) } else { BITONIC_CASE(uint64_t) }
#undef BITONIC_CASE
}

}  // namespace spada
