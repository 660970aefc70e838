// esc_cta_bitonic.cu -- CTA-per-row ESC kernels (bins 6..8: 1024 / 2048 / 4096 products) with the
// hybrid bitonic sort: every warp sorts its chunk in registers, chunks are merged through shared
// memory.  (A stable 4-bit LSD radix sort in shared memory was tried for these bins and for an
// 8192 bin: 30+ barriers per row made it 1.3-1.5x slower than this network at these sizes.)
#include "cta_common.cuh"

namespace spada {

template <int N>
__global__ void __launch_bounds__(ESC_CTA_THREADS)
k_bitonic_symbolic_cta(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                   uint32_t* __restrict__ row_nnz) {
    __shared__ __align__(16) uint32_t s_keys[N];
    __shared__ CtaStage st;
    int* s_wtot = st.wtot;
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    const int p = bitonic_cta_expand<uint32_t, N, false>(a, b, a_begin, a_end, s_keys, nullptr, st);
    for (int t = p + threadIdx.x; t < N; t += ESC_CTA_THREADS) s_keys[t] = 0xffffffffu;
    __syncthreads();
    bitonic_cta_sort<uint32_t, N>(s_keys);
    int cnt = 0;
    for (int i = threadIdx.x; i < p; i += ESC_CTA_THREADS)
        if (i == 0 || s_keys[i] != s_keys[i - 1]) ++cnt;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
    if (lane_id() == 0) s_wtot[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < ESC_CTA_THREADS / 32; ++w) t += s_wtot[w];
        row_nnz[r] = (uint32_t)t;
    }
}

template <typename K, int N>
__global__ void __launch_bounds__(ESC_CTA_THREADS)
k_bitonic_numeric_cta(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                      const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
                      uint32_t* __restrict__ row_nnz_out) {
    constexpr int SB = Log2<N>::v;
    constexpr int ITEMS = N / ESC_CTA_THREADS;
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    double* vals = reinterpret_cast<double*>(s_raw + sizeof(K) * N);
    __shared__ CtaStage st;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    const int p = bitonic_cta_expand<K, N, true>(a, b, a_begin, a_end, keys, vals, st);
    for (int t = p + threadIdx.x; t < N; t += ESC_CTA_THREADS) keys[t] = KeyTraits<K>::sentinel;
    __syncthreads();
    bitonic_cta_sort<K, N>(keys);
    const int total = cta_reduce_store<K, N>(keys, vals, p, c_ptr[r], c_col, c_val, st);
    if (row_nnz_out && threadIdx.x == 0) row_nnz_out[r] = (uint32_t)total;
}

// ---- kept-keys variants: symbolic leaves the sorted packed keys in HBM, numeric reloads them ----------
template <typename K, int N>
__global__ void __launch_bounds__(ESC_CTA_THREADS)
k_bitonic_symbolic_keep_cta(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                            uint32_t* __restrict__ row_nnz, const int64_t* __restrict__ prod_ptr, K* __restrict__ kstore) {
    constexpr int SB = Log2<N>::v;
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    __shared__ CtaStage st;
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    const int p = bitonic_cta_expand<K, N, false, true>(a, b, a_begin, a_end, keys, nullptr, st);
    for (int t = p + threadIdx.x; t < N; t += ESC_CTA_THREADS) keys[t] = KeyTraits<K>::sentinel;
    __syncthreads();
    bitonic_cta_sort<K, N>(keys);
    int cnt = 0;
    K* dst = kstore + prod_ptr[r];
    for (int i = threadIdx.x; i < p; i += ESC_CTA_THREADS) {
        const K ki = keys[i];
        dst[i] = ki;
        if (i == 0 || (uint32_t)(keys[i - 1] >> SB) != (uint32_t)(ki >> SB)) ++cnt;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
    __syncthreads();
    if (lane_id() == 0) st.wtot[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < ESC_CTA_THREADS / 32; ++w) t += st.wtot[w];
        row_nnz[r] = (uint32_t)t;
    }
}

template <typename K, int N>
__global__ void __launch_bounds__(ESC_CTA_THREADS)
k_bitonic_numeric_presorted_cta(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                                const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col,
                                double* __restrict__ c_val, const int64_t* __restrict__ prod_ptr,
                                const K* __restrict__ kstore) {
    constexpr int SB = Log2<N>::v;
    constexpr int ITEMS = N / ESC_CTA_THREADS;
    extern __shared__ __align__(16) unsigned char s_raw[];
    double* vals = reinterpret_cast<double*>(s_raw);
    K* keys = reinterpret_cast<K*>(s_raw + sizeof(double) * N);
    __shared__ CtaStage st;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    const int p = bitonic_cta_expand<K, N, true, true, false>(a, b, a_begin, a_end, keys, vals, st);
    const K* src = kstore + prod_ptr[r];
    for (int t = threadIdx.x; t < p; t += ESC_CTA_THREADS) keys[t] = src[t];
    __syncthreads();
    cta_reduce_store<K, N>(keys, vals, p, c_ptr[r], c_col, c_val, st);
}

template <typename K, int N>
static void keep_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm, uint32_t rows,
                        uint32_t* row_nnz, const int64_t* prod_ptr, void* kstore, cudaStream_t s) {
    size_t smem = sizeof(K) * N;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_bitonic_symbolic_keep_cta<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    k_bitonic_symbolic_keep_cta<K, N><<<rows, ESC_CTA_THREADS, smem, s>>>(a, b, row_begin, perm, rows, row_nnz, prod_ptr,
                                                                      reinterpret_cast<K*>(kstore));
}
template <typename K, int N>
static void presorted_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm, uint32_t rows,
                             const int64_t* c_ptr, int32_t* c_col, double* c_val, const int64_t* prod_ptr,
                             const void* kstore, cudaStream_t s) {
    size_t smem = (sizeof(K) + sizeof(double)) * N;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_bitonic_numeric_presorted_cta<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    }
    k_bitonic_numeric_presorted_cta<K, N><<<rows, ESC_CTA_THREADS, smem, s>>>(
        a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, reinterpret_cast<const K*>(kstore));
}

void launch_cta_symbolic_keep(int bin, bool wide, const DevCsr& a, const DevCsr& b, int64_t row_begin,
                              const uint32_t* perm, uint32_t rows, uint32_t* row_nnz, const int64_t* prod_ptr,
                              void* kstore, cudaStream_t s) {
    if (rows == 0) return;
#define KEEP_CASE(K)                                                                                       \
    switch (bin) {                                                                                          \
        case 6: keep_launch<K, 1024>(a, b, row_begin, perm, rows, row_nnz, prod_ptr, kstore, s); break;     \
        case 7: keep_launch<K, 2048>(a, b, row_begin, perm, rows, row_nnz, prod_ptr, kstore, s); break;     \
        default: keep_launch<K, 4096>(a, b, row_begin, perm, rows, row_nnz, prod_ptr, kstore, s); break;    \
    }
    if (wide) { KEEP_CASE(uint64_t) } else { KEEP_CASE(uint32_t) }
#undef KEEP_CASE
}
void launch_cta_numeric_presorted(int bin, bool wide, const DevCsr& a, const DevCsr& b, int64_t row_begin,
                                  const uint32_t* perm, uint32_t rows, const int64_t* c_ptr, int32_t* c_col,
                                  double* c_val, const int64_t* prod_ptr, const void* kstore, cudaStream_t s) {
    if (rows == 0) return;
#define PRE_CASE(K)                                                                                                   \
    switch (bin) {                                                                                                     \
        case 6: presorted_launch<K, 1024>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, kstore, s); break;  \
        case 7: presorted_launch<K, 2048>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, kstore, s); break;  \
        default: presorted_launch<K, 4096>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, prod_ptr, kstore, s); break; \
    }
    if (wide) { PRE_CASE(uint64_t) } else { PRE_CASE(uint32_t) }
#undef PRE_CASE
}

template <typename K, int N>
static void bitonic_numeric_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                   uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                   uint32_t* nnz_out) {
    size_t smem = (sizeof(K) + sizeof(double)) * N;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_bitonic_numeric_cta<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    k_bitonic_numeric_cta<K, N><<<rows, ESC_CTA_THREADS, smem, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val,
                                                                    nnz_out);
}

void launch_bitonic_cta_symbolic(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                 uint32_t rows, uint32_t* row_nnz, cudaStream_t s) {
    if (rows == 0) return;
    switch (bin) {
        case 6: k_bitonic_symbolic_cta<1024><<<rows, ESC_CTA_THREADS, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        case 7: k_bitonic_symbolic_cta<2048><<<rows, ESC_CTA_THREADS, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        default: k_bitonic_symbolic_cta<4096><<<rows, ESC_CTA_THREADS, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
    }
}

void launch_bitonic_cta_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                uint32_t* nnz_out) {
    if (rows == 0) return;
    int sb = 4 + bin;
    bool narrow = (uint64_t)b.cols <= (1ull << (32 - sb));
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
