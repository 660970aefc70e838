// heavy_smem.cu -- stages 2 and 3 for rows with 4097 .. 65536 intermediate products (bin 9):
// one CTA (1024 threads) per row, the row's occupancy bitmap lives in SHARED memory.
//
// The global-bitmap path of heavy.cu touches one cold 32-byte sector per product in each of
// its passes (measured: 160-210 B of DRAM traffic per product, 27-40 G atomics/s).  Here the
// bitmap of up to 2^20 columns (128 KB) plus one 16-bit rank per word (64 KB) fit the 227 KB of
// shared memory a B200 CTA can have, so setting bits and looking ranks up never leaves the SM;
// only the value accumulation uses global atomics, into the row's own slice of C (a few hundred
// KB per row, <= 148 rows in flight: L2 resident).  B matrices wider than 2^20 columns are
// processed in column-range passes of 2^20 columns, each pass streaming the row's products again.
//
// Reference logic replaced: the K-tiled partial rows and their adder-tree merges
// (scheduler.rs:381-480, 820-920; adder_tree.rs:73-83, 145-188) -- the merge happens in place,
// rank(j) in the bitmap = slot of C[i,j] in the sorted row.  Summation order of one C[i,j] is not
// fixed (atomics), as in the reference (scheduler.rs:386, 396, 827); tested at relative 1e-12.
#include "common.cuh"

namespace spada {

constexpr int HS_THREADS = 1024;
constexpr int HS_WARPS = HS_THREADS / 32;
constexpr int HS_WORDS = 32768;                 // bitmap words per pass: 2^20 columns
constexpr int HS_GROUP = 8;                     // words per rank group
// bitmap (128 KB) + one 16-bit rank per word (64 KB): 192 KB of the 227 KB a CTA can have
constexpr size_t HS_SMEM = sizeof(uint32_t) * HS_WORDS + sizeof(uint16_t) * HS_WORDS;

__device__ __forceinline__ uint32_t hs_block_sum(uint32_t v, uint32_t* s_warp) {
    const int lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < HS_WARPS; ++w) t += s_warp[w];
    return t;
}

// exclusive scan over the block (value per thread), returns the exclusive prefix; total in `total`
__device__ __forceinline__ uint32_t hs_block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < HS_WARPS; ++w) {
        uint32_t t = s_warp[w];
        if (w < warp) base += t;
        all += t;
    }
    total = all;
    return base + x - v;
}

template <bool NUMERIC, typename F>
__device__ __forceinline__ void hs_stream(const DevCsr& a, const DevCsr& b, int64_t a_begin, int64_t a_end, F&& f) {
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    // A entries one warp takes per turn: 32 when the row keeps all 32 warps busy that way, fewer otherwise -- a row of
    // 60 entries with long B rows (R-MAT) would occupy two warps of the CTA's 32; the products of a few entries are
    // still dealt over all lanes
    const int64_t per_warp = (a_end - a_begin + HS_WARPS - 1) / HS_WARPS;
    const int sub = per_warp >= 32 ? 32 : (per_warp < 1 ? 1 : (int)per_warp);
    for (int64_t pb = a_begin + (int64_t)warp * sub; pb < a_end; pb += (int64_t)HS_WARPS * sub) {
        const int64_t pe = pb + sub < a_end ? pb + sub : a_end;
        int bt;
        expand_batch<NUMERIC, true>(a, b, pb + lane, pe, lane, 0, bt,
                                    [&](int, uint32_t c, double av, double bv) { f(c, av, bv); });
    }
}

__device__ __forceinline__ void hs_set_bits(const DevCsr& a, const DevCsr& b, int64_t a_begin, int64_t a_end,
                                            uint32_t* bm, uint32_t words, uint32_t c0, uint32_t c1) {
    for (uint32_t w = threadIdx.x; w < words; w += HS_THREADS) bm[w] = 0u;
    __syncthreads();
    hs_stream<false>(a, b, a_begin, a_end, [&](uint32_t c, double, double) {
        if (c >= c0 && c < c1) {
            const uint32_t d = c - c0;
            const uint32_t bit = 1u << (d & 31);
            if (!(bm[d >> 5] & bit)) atomicOr(&bm[d >> 5], bit);
        }
    });
    __syncthreads();
}

__global__ void __launch_bounds__(HS_THREADS)
k_heavy_smem_symbolic(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ rows_list, uint32_t n_rows,
                      uint32_t* __restrict__ row_nnz) {
    extern __shared__ __align__(16) uint32_t s_u32[];
    uint32_t* bm = s_u32;
    __shared__ uint32_t s_warp[HS_WARPS];
    const uint32_t r = rows_list ? rows_list[blockIdx.x] : blockIdx.x;
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    uint32_t nnz = 0;
    for (int64_t c0 = 0; c0 < b.cols; c0 += (int64_t)HS_WORDS * 32) {
        const int64_t c1 = (c0 + (int64_t)HS_WORDS * 32 < b.cols) ? c0 + (int64_t)HS_WORDS * 32 : b.cols;
        const uint32_t words = (uint32_t)((c1 - c0 + 31) / 32);
        hs_set_bits(a, b, a_begin, a_end, bm, words, (uint32_t)c0, (uint32_t)c1);
        uint32_t cnt = 0;
        for (uint32_t w = threadIdx.x; w < words; w += HS_THREADS) cnt += __popc(bm[w]);
        nnz += hs_block_sum(cnt, s_warp);
        __syncthreads();
    }
    if (threadIdx.x == 0) row_nnz[r] = nnz;
}

// bitmap -> ranks -> column ids -> values of one row; the row's slice of C starts at cbase (its values
// are zeroed here, range by range, once the ranks are known).  Returns nnz of the row.
__device__ __forceinline__ uint32_t hs_numeric_row(const DevCsr& a, const DevCsr& b, int64_t row_begin, uint32_t r,
                                                   int64_t cbase, int32_t* __restrict__ c_col,
                                                   double* __restrict__ c_val, uint32_t* bm, uint16_t* rank16,
                                                   uint32_t* s_warp) {
    const int64_t a_begin = a.ptr[row_begin + r], a_end = a.ptr[row_begin + r + 1];
    uint32_t pass_base = 0;  // outputs of the column ranges already done
    for (int64_t c0 = 0; c0 < b.cols; c0 += (int64_t)HS_WORDS * 32) {
        const int64_t c1 = (c0 + (int64_t)HS_WORDS * 32 < b.cols) ? c0 + (int64_t)HS_WORDS * 32 : b.cols;
        const uint32_t words = (uint32_t)((c1 - c0 + 31) / 32);
        const uint32_t groups = (words + HS_GROUP - 1) / HS_GROUP;
        hs_set_bits(a, b, a_begin, a_end, bm, words, (uint32_t)c0, (uint32_t)c1);
        // ranks: rank16[w] = outputs of this column range before word w (block scan over groups of 8 words)
        uint32_t run = pass_base;
        for (uint32_t gb = 0; gb < groups; gb += HS_THREADS) {
            const uint32_t g = gb + threadIdx.x;
            uint32_t s = 0;
            if (g < groups) {
#pragma unroll
                for (int j = 0; j < HS_GROUP; ++j) {
                    const uint32_t w = g * HS_GROUP + j;
                    if (w < words) s += __popc(bm[w]);
                }
            }
            uint32_t total;
            const uint32_t ex = hs_block_excl_scan(s, s_warp, total);
            if (g < groups) {
                uint32_t before = run + ex - pass_base;   // < 65536: a row of this bin has at most 65536 products
#pragma unroll
                for (int j = 0; j < HS_GROUP; ++j) {
                    const uint32_t w = g * HS_GROUP + j;
                    if (w < words) {
                        rank16[w] = (uint16_t)before;
                        before += __popc(bm[w]);
                    }
                }
            }
            run += total;
        }
        for (uint32_t i = pass_base + threadIdx.x; i < run; i += HS_THREADS) c_val[cbase + i] = 0.0;
        __syncthreads();
        // column ids of this range, in order
        for (uint32_t w = threadIdx.x; w < words; w += HS_THREADS) {
            uint32_t bits = bm[w];
            if (bits) {
                int64_t o = cbase + pass_base + rank16[w];
                while (bits) {
                    const int bit = __ffs(bits) - 1;
                    bits &= bits - 1;
                    c_col[o++] = (int32_t)((uint32_t)c0 + w * 32u + bit);
                }
            }
        }
        __syncthreads();  // also orders the zeroing of c_val before the adds below (same CTA)
        hs_stream<true>(a, b, a_begin, a_end, [&](uint32_t c, double av, double bv) {
            if (c >= (uint32_t)c0 && c < (uint32_t)c1) {
                const uint32_t d = c - (uint32_t)c0;
                const uint32_t w = d >> 5;
                const uint32_t pos = pass_base + rank16[w] + __popc(bm[w] & ((1u << (d & 31)) - 1u));
                atomicAdd(&c_val[cbase + pos], __dmul_rn(av, bv));
            }
        });
        pass_base = run;
        __syncthreads();
    }
    return pass_base;
}

__global__ void __launch_bounds__(HS_THREADS)
k_heavy_smem_numeric(DevCsr a, DevCsr b, int64_t row_begin, const uint32_t* __restrict__ rows_list, uint32_t n_rows,
                     const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_col, double* __restrict__ c_val,
                     uint32_t* __restrict__ row_nnz_out) {
    extern __shared__ __align__(16) uint32_t s_u32[];
    __shared__ uint32_t s_warp[HS_WARPS];
    const uint32_t r = rows_list ? rows_list[blockIdx.x] : blockIdx.x;
    const uint32_t nnz = hs_numeric_row(a, b, row_begin, r, c_ptr[r], c_col, c_val, s_u32, reinterpret_cast<uint16_t*>(s_u32 + HS_WORDS), s_warp);
    if (row_nnz_out && threadIdx.x == 0) row_nnz_out[r] = nnz;   // one-shot mode: the slice at c_ptr[r] is a scratch row
}

void launch_heavy_smem_symbolic(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list,
                                uint32_t n_rows, uint32_t* row_nnz, cudaStream_t s) {
    if (n_rows == 0) return;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_heavy_smem_symbolic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HS_SMEM);
        cudaFuncSetAttribute(k_heavy_smem_numeric, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HS_SMEM);
    }
    k_heavy_smem_symbolic<<<n_rows, HS_THREADS, HS_SMEM, s>>>(a, b, row_begin, rows_list, n_rows, row_nnz);
}

void launch_heavy_smem_numeric(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list,
                               uint32_t n_rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                               uint32_t* row_nnz_out) {
    if (n_rows == 0) return;
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(k_heavy_smem_symbolic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HS_SMEM);
        cudaFuncSetAttribute(k_heavy_smem_numeric, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HS_SMEM);
    }
    k_heavy_smem_numeric<<<n_rows, HS_THREADS, HS_SMEM, s>>>(a, b, row_begin, rows_list, n_rows, c_ptr, c_col, c_val,
                                                             row_nnz_out);
}

}  // namespace spada
