// heavy.cu -- stages 2 and 3 for rows with more than 4096 intermediate products (bin 9): one
// CTA per row, an occupancy bitmap over B's column space plus per-word ranks, so every product
// finds its slot in the (sorted) output row directly -- no hash probing and no sort.
//
// Reference logic replaced: rows longer than a window are K-tiled into several partial rows
// that the adder trees merge later (scheduler.rs:522-524, merge_task :381-480,
// in_cache_merge_task :820-920; adder_tree.rs:73-83, 145-188).  Here the merge happens in
// place: bitmap bit j set <=> some product lands on column j; rank(j) = number of set bits
// below j = position of C[i,j] inside the row.
//
// Summation order inside one C[i,j] is not fixed on this path (partial products arrive through
// atomics) -- like the reference, whose merge order depends on HashMap iteration order
// (scheduler.rs:386, 396, 827); results agree with the oracle to within a few ulp (tested at
// rel 1e-12), structure is exact.
//
// Workspace: per resident CTA one uint2 {bits, rank} per 32 columns of B, kept all-zero between
// rows (each row clears what it set).
#include "common.cuh"

namespace spada {

constexpr int HEAVY_THREADS = 512;
constexpr int HEAVY_WARPS = HEAVY_THREADS / 32;
constexpr int HEAVY_ACC = 12288;  // f64 accumulators in shared memory (96 KB): rows up to this many nnz
constexpr size_t HEAVY_WS_BUDGET_WORDS = (size_t)1 << 27;  // 1 GiB of uint2

static size_t words_per_cta(int64_t b_cols) { return (size_t)((b_cols + 31) / 32) + 1; }
int heavy_grid(uint32_t rows, int sm_count, int64_t b_cols) {
    size_t g = 2 * (size_t)sm_count;
    if (g > rows) g = rows;
    size_t maxg = HEAVY_WS_BUDGET_WORDS / words_per_cta(b_cols);
    if (g > maxg) g = maxg;
    return g < 1 ? 1 : (int)g;
}
size_t heavy_workspace_words(int grid, int64_t b_cols) { return (size_t)grid * words_per_cta(b_cols); }

__device__ __forceinline__ void set_bits(const DevCsr& a, const DevCsr& b, int64_t a_begin, int64_t a_end,
                                         uint2* ws) {
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    for (int64_t pb = a_begin + warp * 32; pb < a_end; pb += HEAVY_THREADS) {
        int bt;
        expand_batch<false, true>(a, b, pb + lane, a_end, lane, 0, bt, [&](int, int64_t q, double) {
            uint32_t c = (uint32_t)ldg_i32(b.col + q);
            atomicOr(&ws[c >> 5].x, 1u << (c & 31));
        });
    }
}

__global__ void __launch_bounds__(HEAVY_THREADS)
k_heavy_symbolic(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                 uint32_t* __restrict__ row_nnz, uint2* ws_all, uint32_t words) {
    __shared__ int s_cnt[HEAVY_WARPS];
    uint2* ws = ws_all + (size_t)blockIdx.x * (words + 1);
    for (uint32_t idx = blockIdx.x; idx < rows; idx += gridDim.x) {
        const uint32_t r = perm ? perm[idx] : idx;
        set_bits(a, b, a.ptr[row_begin + r], a.ptr[row_begin + r + 1], ws);
        __syncthreads();
        int cnt = 0;
        for (uint32_t w = threadIdx.x; w < words; w += HEAVY_THREADS) {
            uint32_t bits = __ldcg(&ws[w].x);
            if (bits) {
                cnt += __popc(bits);
                ws[w].x = 0u;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
        if (lane_id() == 0) s_cnt[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < HEAVY_WARPS; ++w) t += s_cnt[w];
            row_nnz[r] = (uint32_t)t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(HEAVY_THREADS)
k_heavy_numeric(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ perm, uint32_t rows,
                const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
                uint2* ws_all, uint32_t words) {
    extern __shared__ __align__(16) double s_acc[];
    __shared__ uint32_t s_wtot[HEAVY_WARPS];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    uint2* ws = ws_all + (size_t)blockIdx.x * (words + 1);
    for (uint32_t idx = blockIdx.x; idx < rows; idx += gridDim.x) {
        const uint32_t r = perm ? perm[idx] : idx;
        const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
        const int64_t cbase = c_ptr[r];
        const int64_t z = c_ptr[r + 1] - cbase;
        const bool in_smem = z <= HEAVY_ACC;
        // 1. occupancy bitmap
        set_bits(a, b, a_begin, a_end, ws);
        if (in_smem) {
            for (int i = threadIdx.x; i < (int)z; i += HEAVY_THREADS) s_acc[i] = 0.0;
        } else {
            for (int64_t i = threadIdx.x; i < z; i += HEAVY_THREADS) c_val[cbase + i] = 0.0;
        }
        __syncthreads();
        // 2. per-word ranks (exclusive prefix of popcounts) and the row's column ids
        uint32_t run = 0;
        for (uint32_t wb = 0; wb < words; wb += HEAVY_THREADS) {
            uint32_t w = wb + threadIdx.x;
            uint32_t bits = (w < words) ? __ldcg(&ws[w].x) : 0u;
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
            if (bits) {
                uint32_t rank = wbase + x - pc;
                ws[w].y = rank;
                uint32_t bb = bits;
                int64_t o = cbase + rank;
                while (bb) {
                    int bit = __ffs(bb) - 1;
                    bb &= bb - 1;
                    c_col[o++] = (int32_t)(w * 32u + bit);
                }
            }
            run += all;
            __syncthreads();
        }
        // 3. products -> slots
        for (int64_t pb = a_begin + warp * 32; pb < a_end; pb += HEAVY_THREADS) {
            int bt;
            expand_batch<true, true>(a, b, pb + lane, a_end, lane, 0, bt, [&](int, int64_t q, double av) {
                uint32_t c = (uint32_t)ldg_i32(b.col + q);
                double prod = __dmul_rn(av, ldg_f64(b.val + q));
                uint2 e = __ldcg(&ws[c >> 5]);
                uint32_t pos = e.y + __popc(e.x & ((1u << (c & 31)) - 1u));
                if (in_smem)
                    atomicAdd(&s_acc[pos], prod);
                else
                    atomicAdd(&c_val[cbase + pos], prod);
            });
        }
        __syncthreads();
        // 4. store values, clear the bitmap for the next row
        if (in_smem)
            for (int i = threadIdx.x; i < (int)z; i += HEAVY_THREADS) c_val[cbase + i] = s_acc[i];
        for (uint32_t w = threadIdx.x; w < words; w += HEAVY_THREADS)
            if (__ldcg(&ws[w].x)) ws[w] = make_uint2(0u, 0u);
        __syncthreads();
    }
}

void launch_heavy_symbolic(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                           uint32_t rows, uint32_t* row_nnz, uint2* ws, int grid, cudaStream_t s) {
    if (rows == 0) return;
    uint32_t words = (uint32_t)((b.cols + 31) / 32);
    k_heavy_symbolic<<<grid, HEAVY_THREADS, 0, s>>>(a, b, row_begin, perm, rows, row_nnz, ws, words);
}

void launch_heavy_numeric(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                          uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, uint2* ws, int grid,
                          cudaStream_t s) {
    if (rows == 0) return;
    uint32_t words = (uint32_t)((b.cols + 31) / 32);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_heavy_numeric, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(HEAVY_ACC * sizeof(double)));
        attr = true;
    }
    k_heavy_numeric<<<grid, HEAVY_THREADS, HEAVY_ACC * sizeof(double), s>>>(a, b, row_begin, perm, rows, c_ptr,
                                                                          c_col, c_val, ws, words);
}

}  // namespace spada
