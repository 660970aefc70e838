// esc_cta_bitonic.cu -- CTA-per-row ESC kernels (bins 6..8: 1024 / 2048 / 4096 products) with the
// hybrid bitonic sort: every warp sorts its chunk in registers, chunks are merged through shared
// memory.  (A stable 4-bit LSD radix sort in shared memory was tried for these bins and for an
// 8192 bin: 30+ barriers per row made it 1.3-1.5x slower than this network at these sizes.)
#include "cta_common.cuh"

namespace spada {

// G = 2: two groups of N / 2 sorted with 32-bit keys, then merged (cta_merge_groups2) -- one more bit of column id
// before the 64-bit network is needed
template <typename K, int N, int G = 1>
__global__ void __launch_bounds__(ESC_CTA_THREADS)
k_bitonic_numeric_cta(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                      const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
                      uint32_t* __restrict__ row_nnz_out) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    double* vals = reinterpret_cast<double*>(s_raw + sizeof(K) * N);
    __shared__ CtaStage st;
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    const int p = bitonic_cta_expand<K, N, true, true, true, G>(a, b, a_begin, a_end, keys, vals, st);
    for (int t = p + threadIdx.x; t < N; t += ESC_CTA_THREADS) keys[t] = KeyTraits<K>::sentinel;
    __syncthreads();
    bitonic_cta_sort<K, N, G>(keys);
    int total;
    if constexpr (G == 2) {
        cta_merge_groups2<N>(keys, vals, p);
        total = cta_reduce_store<K, N, true>(keys, vals, p, c_ptr[r], c_col, c_val, st);
    } else {
        total = cta_reduce_store<K, N>(keys, vals, p, c_ptr[r], c_col, c_val, st);
    }
    if (row_nnz_out && threadIdx.x == 0) row_nnz_out[r] = (uint32_t)total;
}

template <typename K, int N, int G = 1>
static void bitonic_numeric_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                   uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                   uint32_t* nnz_out) {
    size_t smem = (sizeof(K) + sizeof(double)) * N;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_bitonic_numeric_cta<K, N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    k_bitonic_numeric_cta<K, N, G><<<rows, ESC_CTA_THREADS, smem, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val,
                                                                    nnz_out);
}

void launch_bitonic_cta_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                uint32_t* nnz_out) {
    if (rows == 0) return;
    int sb = 4 + bin;
    bool narrow = (uint64_t)b.cols <= (1ull << (32 - sb));
    if (!narrow && (uint64_t)b.cols <= (1ull << (33 - sb))) {   // one bit short: two 32-bit groups + a merge
        switch (bin) {
            case 6: bitonic_numeric_launch<uint32_t, 1024, 2>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break;
            case 7: bitonic_numeric_launch<uint32_t, 2048, 2>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break;
            default: bitonic_numeric_launch<uint32_t, 4096, 2>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s, nnz_out); break;
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
