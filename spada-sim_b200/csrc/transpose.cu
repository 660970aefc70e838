// transpose.cu -- structural transpose of a device CSR operand: B = A^T as canonical CSR.
//
// The step before the hot path for every non-square SS workload: GEMM::from_mat (gemm.rs:41-53) builds
// B = A.transpose_into().to_csr() on the host (sprs counting transpose, single thread); here the same
// result -- rows of A^T in ascending order of A's row index, values moved, nothing computed -- is produced on
// the device from the resident A, bit-identical to scipy's a.T.tocsr() with sorted indices.
//
// Method: a stable LSD radix sort of the entry indices by column id, 8 bits per pass (ceil(log2(cols)/8)
// passes), hand written:
//   k_radix_hist     one warp per tile of 2048 entries: digit histogram, written digit-major so that one
//                    exclusive scan (plan.cu) turns all (digit, tile) counts into global offsets
//   k_radix_scatter  the same warp walks its tile 32 entries at a time in order; lanes with the same digit
//                    are found with match.any, ranked with a popcount, and the group leader advances the
//                    warp's running offset of that digit in shared memory -- warp-synchronous, no barriers,
//                    stable by construction (entries of a digit keep their arrival order)
// Entries are visited in CSR order (row-major), so a stable sort by column leaves every column's entries in
// ascending row order: exactly the transposed row.  Then one gather builds (row id, value) of the sorted
// entries, and the column counts (one atomic per entry) are scanned into the transposed row_ptr.
// HBM traffic per pass: 4 B (histogram) + 8 B read + 8 B written per nonzero; scattered 4-byte writes.
#include "common.cuh"

namespace spada {

constexpr int TR_WARPS = 8;
constexpr int TR_TILE = 2048;   // entries per warp tile
constexpr int TR_BINS = 256;

__global__ void k_entry_rows(DevCsr a, uint32_t* __restrict__ erow, uint32_t* __restrict__ col_count) {
    // 16 lanes per row: erow[e] = row of entry e; col_count[c] += 1 for every entry of column c
    const int64_t r = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 4;
    const int sub = threadIdx.x & 15;
    if (r >= a.rows) return;
    const int64_t s0 = a.ptr[r], s1 = a.ptr[r + 1];
    for (int64_t e = s0 + sub; e < s1; e += 16) {
        erow[e] = (uint32_t)r;
        atomicAdd(&col_count[a.col[e]], 1u);
    }
}

// digit of entry e in this pass: pass 0 reads A's column ids in place, later passes the ping-pong keys
__device__ __forceinline__ uint32_t tr_digit(const int32_t* __restrict__ keys, int64_t e, int shift) {
    return ((uint32_t)keys[e] >> shift) & (TR_BINS - 1);
}

__global__ void __launch_bounds__(TR_WARPS * 32)
k_radix_hist(const int32_t* __restrict__ keys, int64_t n, int shift, int64_t n_tiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[TR_WARPS][TR_BINS];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * TR_WARPS + warp;
    for (int d = lane; d < TR_BINS; d += 32) s_hist[warp][d] = 0u;
    __syncwarp();
    if (tile < n_tiles) {
        const int64_t e0 = tile * TR_TILE;
        for (int i = 0; i < TR_TILE; i += 32) {
            const int64_t e = e0 + i + lane;
            const uint32_t d = e < n ? tr_digit(keys, e, shift) : 0xffffffffu;   // tail lanes form their own group
            const unsigned grp = __match_any_sync(FULL, d);
            if (e < n && lane == __ffs(grp) - 1) s_hist[warp][d] += (uint32_t)__popc(grp);
            __syncwarp();
        }
        for (int d = lane; d < TR_BINS; d += 32) hist[(int64_t)d * n_tiles + tile] = s_hist[warp][d];
    }
}

__global__ void __launch_bounds__(TR_WARPS * 32)
k_radix_scatter(const int32_t* __restrict__ keys, const uint32_t* __restrict__ pay, int64_t n, int shift, int64_t n_tiles,
                const int64_t* __restrict__ offs, int32_t* __restrict__ keys_out, uint32_t* __restrict__ pay_out) {
    __shared__ int64_t s_off[TR_WARPS][TR_BINS];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * TR_WARPS + warp;
    if (tile >= n_tiles) return;
    for (int d = lane; d < TR_BINS; d += 32) s_off[warp][d] = offs[(int64_t)d * n_tiles + tile];
    __syncwarp();
    const int64_t e0 = tile * TR_TILE;
    for (int i = 0; i < TR_TILE; i += 32) {
        const int64_t e = e0 + i + lane;
        const bool valid = e < n;
        int32_t key = 0;
        uint32_t d = 0xffffffffu;
        if (valid) {
            key = keys[e];
            d = ((uint32_t)key >> shift) & (TR_BINS - 1);
        }
        const unsigned grp = __match_any_sync(FULL, d);
        const int leader = __ffs(grp) - 1;
        int64_t base = 0;
        if (valid && lane == leader) {
            base = s_off[warp][d];
            s_off[warp][d] = base + __popc(grp);
        }
        base = shfl_i64(base, leader);
        if (valid) {
            const int64_t pos = base + __popc(grp & ((1u << lane) - 1u));
            keys_out[pos] = key;
            pay_out[pos] = pay ? pay[e] : (uint32_t)e;   // pass 0: the payload is the entry index itself
        }
        __syncwarp();
    }
}

__global__ void k_transpose_gather(const uint32_t* __restrict__ pay, const uint32_t* __restrict__ erow,
                                   const double* __restrict__ val, int64_t n, int32_t* __restrict__ t_col,
                                   double* __restrict__ t_val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t e = pay ? pay[i] : (uint32_t)i;
        t_col[i] = (int32_t)erow[e];
        t_val[i] = val[e];
    }
}

int transpose_passes(int64_t cols) {
    int bits = 0;
    while (bits < 31 && (1ll << bits) < cols) ++bits;
    return (bits + 7) / 8;
}
int64_t transpose_tiles(int64_t nnz) { return (nnz + TR_TILE - 1) / TR_TILE; }

void launch_entry_rows(const DevCsr& a, uint32_t* erow, uint32_t* col_count, cudaStream_t s) {
    if (a.rows > 0) k_entry_rows<<<(unsigned)((a.rows * 16 + 255) / 256), 256, 0, s>>>(a, erow, col_count);
}
void launch_radix_hist(const int32_t* keys, int64_t n, int shift, uint32_t* hist, cudaStream_t s) {
    const int64_t tiles = transpose_tiles(n);
    if (tiles > 0)
        k_radix_hist<<<(unsigned)((tiles + TR_WARPS - 1) / TR_WARPS), TR_WARPS * 32, 0, s>>>(keys, n, shift, tiles, hist);
}
void launch_radix_scatter(const int32_t* keys, const uint32_t* pay, int64_t n, int shift, const int64_t* offs,
                          int32_t* keys_out, uint32_t* pay_out, cudaStream_t s) {
    const int64_t tiles = transpose_tiles(n);
    if (tiles > 0)
        k_radix_scatter<<<(unsigned)((tiles + TR_WARPS - 1) / TR_WARPS), TR_WARPS * 32, 0, s>>>(keys, pay, n, shift, tiles,
                                                                                                offs, keys_out, pay_out);
}
void launch_transpose_gather(const uint32_t* pay, const uint32_t* erow, const double* val, int64_t n, int32_t* t_col,
                             double* t_val, cudaStream_t s) {
    if (n > 0) k_transpose_gather<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pay, erow, val, n, t_col, t_val);
}

}  // namespace spada
