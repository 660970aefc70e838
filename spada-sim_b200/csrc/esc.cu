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
// per lane, E = N/32), four rows share a CTA; bins 6..8 give each A row a whole CTA.
#include "common.cuh"
#include "sort.cuh"

namespace spada {

constexpr int ESC_WARPS = 4;          // rows per CTA in the warp-per-row bins
constexpr int ESC_CTA_THREADS = 256;  // CTA-per-row bins

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
                            [&](int sq, int64_t q, double) { keys[sq] = (uint32_t)ldg_i32(b.col + q); });
        seq += bt;
    }
    for (int t = seq + lane; t < N; t += 32) keys[t] = 0xffffffffu;
    __syncwarp();
    uint32_t x[E];
    load_blocked<uint32_t, E>(x, keys, lane);
    warp_sort<uint32_t, E>(x, lane, false);
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
                   const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val) {
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
        expand_batch<true, false>(a, b, pb + lane, a_end, lane, seq, bt, [&](int sq, int64_t q, double av) {
            uint32_t c = (uint32_t)ldg_i32(b.col + q);
            keys[sq] = ((K)c << SB) | (K)sq;
            vals[sq] = __dmul_rn(av, ldg_f64(b.val + q));
        });
        seq += bt;
    }
    const int p = seq;
    for (int t = p + lane; t < N; t += 32) keys[t] = KeyTraits<K>::sentinel;
    __syncwarp();
    K x[E];
    load_blocked<K, E>(x, keys, lane);
    warp_sort<K, E>(x, lane, false);
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
            c_col[cbase + o] = (int32_t)col;
            c_val[cbase + o] = sum;
        }
        out_base += __popc(hm);
    }
}

// =============================================================================================
// CTA-per-row kernels, N = 1024 / 2048 / 4096 products at most; 8 warps, chunk = N/8 keys per warp
// =============================================================================================
template <typename K, int N, bool NUMERIC>
__device__ __forceinline__ int cta_expand(const DevCsr& a, const DevCsr& b, int64_t a_begin, int64_t a_end,
                                          K* keys, double* vals, int* s_wtot) {
    constexpr int SB = Log2<N>::v;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int seq_base = 0;
    for (int64_t pb = a_begin; pb < a_end; pb += ESC_CTA_THREADS) {
        // pre-pass: per-warp product totals of this batch so every warp knows its arrival offset
        int64_t p = pb + threadIdx.x;
        int len = 0;
        if (p < a_end) {
            int32_t k = ldg_i32(a.col + p);
            len = (int)(ldg_i64(b.ptr + k + 1) - ldg_i64(b.ptr + k));
        }
        int wt = len;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) wt += __shfl_xor_sync(FULL, wt, d);
        if (lane == 0) s_wtot[warp] = wt;
        __syncthreads();
        int my_base = seq_base, all = 0;
#pragma unroll
        for (int w = 0; w < ESC_CTA_THREADS / 32; ++w) {
            int t = s_wtot[w];
            if (w < warp) my_base += t;
            all += t;
        }
        int bt;
        expand_batch<NUMERIC, false>(a, b, p, a_end, lane, my_base, bt, [&](int sq, int64_t q, double av) {
            uint32_t c = (uint32_t)ldg_i32(b.col + q);
            if (NUMERIC) {
                keys[sq] = ((K)c << SB) | (K)sq;
                vals[sq] = __dmul_rn(av, ldg_f64(b.val + q));
            } else {
                keys[sq] = (K)c;
            }
        });
        seq_base += all;
        __syncthreads();
    }
    return seq_base;
}

template <typename K, int N>
__device__ __forceinline__ void cta_sort(K* keys) {
    constexpr int WARPS = ESC_CTA_THREADS / 32;
    constexpr int CH = N / WARPS;  // keys per warp chunk
    constexpr int E = CH / 32;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    K x[E];
    load_blocked<K, E>(x, keys + warp * CH, lane);
    warp_sort<K, E>(x, lane, (warp & 1) != 0);
    store_blocked<K, E>(x, keys + warp * CH, lane);
    __syncthreads();
#pragma unroll 1
    for (int k = 2 * CH; k <= N; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j >= CH; j >>= 1) {
            for (int t = threadIdx.x; t < N / 2; t += ESC_CTA_THREADS) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool up = (i & k) == 0;
                K ka = keys[i], kb = keys[l];
                if ((ka > kb) == up) {
                    keys[i] = kb;
                    keys[l] = ka;
                }
            }
            __syncthreads();
        }
        load_blocked<K, E>(x, keys + warp * CH, lane);
        warp_merge_tail<K, E>(x, lane, ((warp * CH) & k) == 0);
        store_blocked<K, E>(x, keys + warp * CH, lane);
        __syncthreads();
    }
}

template <int N>
__global__ void __launch_bounds__(ESC_CTA_THREADS)
k_esc_symbolic_cta(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                   uint32_t* __restrict__ row_nnz) {
    __shared__ __align__(16) uint32_t s_keys[N];
    __shared__ int s_wtot[ESC_CTA_THREADS / 32];
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    const int p = cta_expand<uint32_t, N, false>(a, b, a_begin, a_end, s_keys, nullptr, s_wtot);
    for (int t = p + threadIdx.x; t < N; t += ESC_CTA_THREADS) s_keys[t] = 0xffffffffu;
    __syncthreads();
    cta_sort<uint32_t, N>(s_keys);
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
k_esc_numeric_cta(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                  const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val) {
    constexpr int SB = Log2<N>::v;
    extern __shared__ __align__(16) unsigned char s_raw[];
    K* keys = reinterpret_cast<K*>(s_raw);
    double* vals = reinterpret_cast<double*>(s_raw + sizeof(K) * N);
    __shared__ int s_wtot[ESC_CTA_THREADS / 32];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t r = perm ? perm[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    const int p = cta_expand<K, N, true>(a, b, a_begin, a_end, keys, vals, s_wtot);
    for (int t = p + threadIdx.x; t < N; t += ESC_CTA_THREADS) keys[t] = KeyTraits<K>::sentinel;
    __syncthreads();
    cta_sort<K, N>(keys);
    const int64_t cbase = c_ptr[r];
    int out_base = 0;
    for (int base = 0; base < p; base += ESC_CTA_THREADS) {
        int i = base + threadIdx.x;
        bool valid = i < p;
        K ki = valid ? keys[i] : KeyTraits<K>::sentinel;
        uint32_t col = (uint32_t)(ki >> SB);
        bool head = valid && (i == 0 || (uint32_t)(keys[i - 1] >> SB) != col);
        unsigned hm = __ballot_sync(FULL, head);
        if (lane == 0) s_wtot[warp] = __popc(hm);
        __syncthreads();
        int wbase = out_base, all = 0;
#pragma unroll
        for (int w = 0; w < ESC_CTA_THREADS / 32; ++w) {
            int t = s_wtot[w];
            if (w < warp) wbase += t;
            all += t;
        }
        if (head) {
            double sum = vals[(int)(ki & (K)(N - 1))];
            for (int j = i + 1; j < p; ++j) {
                K kj = keys[j];
                if ((uint32_t)(kj >> SB) != col) break;
                sum = __dadd_rn(sum, vals[(int)(kj & (K)(N - 1))]);
            }
            int o = wbase + __popc(hm & ((1u << lane) - 1u));
            c_col[cbase + o] = (int32_t)col;
            c_val[cbase + o] = sum;
        }
        out_base += all;
        __syncthreads();
    }
}

// ---- launchers --------------------------------------------------------------------------------
int esc_grid(int bin, uint32_t rows) {
    if (bin <= 5) return (int)((rows + ESC_WARPS - 1) / ESC_WARPS);
    return (int)rows;
}

template <typename K, int N>
static void numeric_cta_launch(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                               uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val,
                               cudaStream_t s) {
    size_t smem = (sizeof(K) + sizeof(double)) * N;
    k_esc_numeric_cta<K, N><<<rows, ESC_CTA_THREADS, smem, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val);
}

void setup_kernel_attributes() {
    cudaFuncSetAttribute(k_esc_numeric_cta<uint32_t, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 4096);
    cudaFuncSetAttribute(k_esc_numeric_cta<uint64_t, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 4096);
    cudaFuncSetAttribute(k_esc_numeric_cta<uint64_t, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2048);
    cudaFuncSetAttribute(k_esc_numeric_cta<uint32_t, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 2048);
    cudaFuncSetAttribute(k_esc_numeric_cta<uint32_t, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 1024);
    cudaFuncSetAttribute(k_esc_numeric_cta<uint64_t, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
}

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
        case 6: k_esc_symbolic_cta<1024><<<g, ESC_CTA_THREADS, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        case 7: k_esc_symbolic_cta<2048><<<g, ESC_CTA_THREADS, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        case 8: k_esc_symbolic_cta<4096><<<g, ESC_CTA_THREADS, 0, s>>>(a, b, row_begin, perm, rows, row_nnz); break;
        default: break;
    }
}

template <typename K>
static void numeric_dispatch(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                             uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s) {
    unsigned g = (unsigned)esc_grid(bin, rows);
    switch (bin) {
        case 1: k_esc_numeric_warp<K, 32><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val); break;
        case 2: k_esc_numeric_warp<K, 64><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val); break;
        case 3: k_esc_numeric_warp<K, 128><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val); break;
        case 4: k_esc_numeric_warp<K, 256><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val); break;
        case 5: k_esc_numeric_warp<K, 512><<<g, ESC_WARPS * 32, 0, s>>>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val); break;
        case 6: numeric_cta_launch<K, 1024>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s); break;
        case 7: numeric_cta_launch<K, 2048>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s); break;
        case 8: numeric_cta_launch<K, 4096>(a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s); break;
        default: break;
    }
}

void launch_esc_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                        uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s) {
    if (rows == 0) return;
    // 32-bit keys whenever (column << log2 N | arrival) fits: b.cols <= 2^(32 - log2 N)
    int sb = 4 + bin;  // log2(N): bin 1 -> 32 = 2^5
    bool narrow = (uint64_t)b.cols <= (1ull << (32 - sb));
    if (narrow)
        numeric_dispatch<uint32_t>(bin, a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s);
    else
        numeric_dispatch<uint64_t>(bin, a, b, row_begin, perm, rows, c_ptr, c_col, c_val, s);
}

}  // namespace spada
