// plan.cu -- stage 1 (flop count + binning = the window choice) and stage 4 (row_ptr scan),
// plus operand conversion / validation kernels.
//
// Reference logic replaced:
//   * row-length tables of the scheduler (scheduler.rs:197-202) and the per-window
//     product counts they imply (b_row_lens lookups, scheduler.rs:654),
//   * the adaptive window shape [R, lane_num/R] of rowwise_perf_adjust.rs:121-252: there R is
//     picked per row group from simulated latency; here each row's intermediate-product
//     count picks the bin, and the bin fixes how many lanes cooperate on a row,
//   * CsrMatStorage::write's indptr maintenance (storage.rs:196-210) -> exclusive scan.
#include "common.cuh"

namespace spada {

constexpr int FLOPS_THREADS = 256;
constexpr int LONG_ROW = 256;  // A rows longer than this are counted by a whole CTA

// ---------------------------------------------------------------------------------------
// K1a: one thread per A row.  HBM traffic: A.row_ptr + A.col streamed once; the length of B row k comes
// from a compact u32 table (4 B per B row, built by k_row_lengths: one gather per A nonzero instead of two
// 8-byte row_ptr reads, and 4(k) <= 17 MB of L2 instead of 34 MB for the benchmark sizes).
__global__ void k_row_lengths(const int64_t* __restrict__ ptr, int64_t rows, uint32_t* __restrict__ len) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) len[i] = (uint32_t)(ptr[i + 1] - ptr[i]);
}

// LPR lanes per row: 1 for short rows (stencils: consecutive threads read consecutive rows), 8 when rows average eight
// entries or more (eight lanes read consecutive entries of one row: a thread per row then touches a different
// 32-byte sector with every load, 8x the column-id traffic)
template <int LPR>
__global__ void __launch_bounds__(FLOPS_THREADS)
k_flops(DevCsr a, const uint32_t* __restrict__ b_len, int64_t row_begin, int64_t m,
        uint32_t* __restrict__ flops, uint32_t* __restrict__ long_list, PlanCounters* ctr) {
    __shared__ uint32_t s_rows[NUM_BINS];
    __shared__ unsigned long long s_prod[NUM_BINS];
    __shared__ uint32_t s_max, s_fit;
    if (threadIdx.x < NUM_BINS) {
        s_rows[threadIdx.x] = 0;
        s_prod[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) {
        s_max = 0;
        s_fit = 0;
    }
    __syncthreads();
    const int64_t i = ((int64_t)blockIdx.x * FLOPS_THREADS + threadIdx.x) / LPR;
    const int sub = threadIdx.x % LPR;
    int b = -1;   // bin of this thread's row, -1: nothing to count here
    bool fits = false;   // a row of bin 1 with at most 8 A entries
    unsigned long long f = 0;
    int64_t s = 0, e = 0;
    if (i < m) {
        s = a.ptr[row_begin + i];
        e = a.ptr[row_begin + i + 1];
        if (e - s <= LONG_ROW)
            for (int64_t p = s + sub; p < e; p += LPR) f += (unsigned long long)__ldg(b_len + ldg_i32(a.col + p));
    }
    if (LPR > 1) {
#pragma unroll
        for (int d = LPR / 2; d > 0; d >>= 1) f += __shfl_xor_sync(FULL, f, d);
    }
    if (i < m && sub == 0) {
        if (e - s > LONG_ROW) {
            uint32_t slot = atomicAdd(&ctr->long_rows, 1u);
            long_list[slot] = (uint32_t)i;
        } else {
            uint32_t f32 = f > 0xffffffffull ? 0xffffffffu : (uint32_t)f;
            flops[i] = f32;
            b = bin_of(f32);
            if (f32 > ESC_MAX_PRODUCTS) atomicMax(&s_max, f32);
            fits = b == 1 && e - s <= 8;
            if (f >> 32) {   // cannot happen below 2^32 products per row; counted directly
                atomicAdd(&s_rows[b], 1u);
                atomicAdd(&s_prod[b], f);
                b = -1;
            }
        }
    }
    // one shared-memory update per (warp, bin): the lanes of a bin are found with match.any, their counts summed
    // with redux (16-bit halves, so 32 lanes cannot overflow) -- per-thread 64-bit shared atomics on one address
    // serialised the whole CTA when every row falls in the same bin (stencils)
    const unsigned fm = __ballot_sync(FULL, fits);
    if (fm && lane_id() == 0) atomicAdd(&s_fit, (uint32_t)__popc(fm));
    const unsigned grp = __match_any_sync(FULL, b);
    if (b >= 0) {
        const unsigned lo = __reduce_add_sync(grp, (unsigned)(f & 0xffffull));
        const unsigned hi = __reduce_add_sync(grp, (unsigned)(f >> 16));
        if (lane_id() == __ffs(grp) - 1) {
            atomicAdd(&s_rows[b], (uint32_t)__popc(grp));
            atomicAdd(&s_prod[b], ((unsigned long long)hi << 16) + lo);
        }
    }
    __syncthreads();
    if (threadIdx.x < NUM_BINS && s_rows[threadIdx.x]) {
        atomicAdd(&ctr->bin_rows[threadIdx.x], s_rows[threadIdx.x]);
        atomicAdd(&ctr->bin_products[threadIdx.x], s_prod[threadIdx.x]);
        atomicAdd(&ctr->total_products, s_prod[threadIdx.x]);
    }
    if (threadIdx.x == 0 && s_max) atomicMax(&ctr->max_flops, s_max);
    if (threadIdx.x == 0 && s_fit) atomicAdd(&ctr->tiny_fit, s_fit);
}

// K1b: long A rows, one CTA of 1024 threads per row (persistent over the deferred list); four
// independent gathers per thread in flight.
constexpr int FLOPS_LONG_THREADS = 1024;
__global__ void __launch_bounds__(FLOPS_LONG_THREADS)
k_flops_long(DevCsr a, const uint32_t* __restrict__ b_len, int64_t row_begin,
             uint32_t* __restrict__ flops, const uint32_t* __restrict__ long_list, PlanCounters* ctr) {
    __shared__ unsigned long long s_warp[FLOPS_LONG_THREADS / 32];
    uint32_t n_long = ctr->long_rows;
    for (uint32_t idx = blockIdx.x; idx < n_long; idx += gridDim.x) {
        uint32_t i = long_list[idx];
        int64_t s = a.ptr[row_begin + i], e = a.ptr[row_begin + i + 1];
        unsigned long long f = 0;
        for (int64_t p0 = s + threadIdx.x; p0 < e; p0 += 4 * FLOPS_LONG_THREADS) {
            int32_t k[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int64_t p = p0 + (int64_t)u * FLOPS_LONG_THREADS;
                k[u] = p < e ? ldg_i32(a.col + p) : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k[u] >= 0) f += (unsigned long long)__ldg(b_len + k[u]);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) f += __shfl_xor_sync(FULL, f, d);
        if (lane_id() == 0) s_warp[threadIdx.x >> 5] = f;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int w = 0; w < FLOPS_LONG_THREADS / 32; ++w) t += s_warp[w];
            uint32_t f32 = t > 0xffffffffull ? 0xffffffffu : (uint32_t)t;
            flops[i] = f32;
            int b = bin_of(f32);
            atomicAdd(&ctr->bin_rows[b], 1u);
            atomicAdd(&ctr->bin_products[b], t);
            atomicAdd(&ctr->total_products, t);
            atomicMax(&ctr->max_flops, f32);
        }
        __syncthreads();
    }
}

// b_len: workspace of b_rows u32 (filled here with the lengths of B's rows)
void launch_flops(const DevCsr& a, const int64_t* b_ptr, int64_t b_rows, uint32_t* b_len, int64_t row_begin, int64_t m,
                  uint32_t* flops, uint32_t* long_list, PlanCounters* ctr, cudaStream_t s) {
    if (m <= 0) return;
    if (b_rows > 0) k_row_lengths<<<(unsigned)((b_rows + 255) / 256), 256, 0, s>>>(b_ptr, b_rows, b_len);
    // one thread per row.  (Eight lanes per row for operands that average >= 8 entries per row were measured: the
    // column ids coalesce, but the kernel is bound by the gathers of the B-row lengths and the extra threads cost more
    // than they save -- rect 0.27 -> 0.33 ms, R-MAT 0.26 -> 0.61 ms.)
    unsigned grid = (unsigned)((m + FLOPS_THREADS - 1) / FLOPS_THREADS);
    k_flops<1><<<grid, FLOPS_THREADS, 0, s>>>(a, b_len, row_begin, m, flops, long_list, ctr);
    k_flops_long<<<148 * 2, FLOPS_LONG_THREADS, 0, s>>>(a, b_len, row_begin, flops, long_list, ctr);
}

// a few bytes from device memory into mapped pinned host memory, written by the SM (no copy engine: engine.cu)
__global__ void k_publish(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int words) {
    for (int i = threadIdx.x; i < words; i += 32) dst[i] = src[i];
    __threadfence_system();
}
}  // namespace spada
void launch_publish(const void* d_src, void* h_dst, size_t bytes, cudaStream_t s) {
    spada::k_publish<<<1, 32, 0, s>>>(reinterpret_cast<const uint32_t*>(d_src), reinterpret_cast<uint32_t*>(h_dst),
                                      (int)(bytes / 4));
}
namespace spada {

// ---------------------------------------------------------------------------------------
// K1c: scatter row ids into per-bin lists.  One smem histogram per CTA, one global
// reservation per (CTA, bin).
__global__ void __launch_bounds__(256)
k_bin_scatter(const uint32_t* __restrict__ flops, int64_t m, BinTable tbl, uint32_t* __restrict__ perm,
              PlanCounters* ctr) {
    __shared__ uint32_t s_cnt[NUM_BINS];
    __shared__ uint32_t s_base[NUM_BINS];
    if (threadIdx.x < NUM_BINS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    int b = 0;
    uint32_t rank = 0;
    if (i < m) {
        b = bin_of(flops[i]);
        if (b != BIN_EMPTY) rank = atomicAdd(&s_cnt[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NUM_BINS && threadIdx.x != BIN_EMPTY && s_cnt[threadIdx.x])
        s_base[threadIdx.x] = tbl.offset[threadIdx.x] + atomicAdd(&ctr->bin_cursor[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (i < m && b != BIN_EMPTY) perm[s_base[b] + rank] = (uint32_t)i;
}

void launch_bin_scatter(const uint32_t* flops, int64_t m, const BinTable& tbl, uint32_t* perm,
                        PlanCounters* ctr, cudaStream_t s) {
    if (m <= 0) return;
    unsigned grid = (unsigned)((m + 255) / 256);
    k_bin_scatter<<<grid, 256, 0, s>>>(flops, m, tbl, perm, ctr);
}

// ---------------------------------------------------------------------------------------
// K4: single-pass exclusive scan (decoupled look-back), u32 per-row nnz -> i64 row_ptr.
// HBM traffic: 4 B read + 8 B written per row.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
#define ST_AGG (1ull << 62)
#define ST_PREFIX (2ull << 62)
#define ST_MASK (3ull << 62)

size_t scan_tile_state_words(int64_t n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1; }

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_u32_i64(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out,
               unsigned long long* tile_state, uint32_t* ticket) {
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;

    uint32_t v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
        const uint4* p4 = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS / 4; ++j) {
            uint4 q = p4[j];
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) v[j] = (base + j < n) ? in[base + j] : 0u;
    }
    unsigned long long tsum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) tsum += v[j];
    // warp inclusive scan of the per-thread sums
    unsigned long long x = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    unsigned long long warp_off = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        unsigned long long t = s_warp[w];
        if (w < warp) warp_off += t;
        tile_total += t;
    }
    // look-back by warp 0
    if (warp == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&tile_state[0], ST_PREFIX | tile_total);
        } else {
            if (lane == 0) atomicExch(&tile_state[tile], ST_AGG | tile_total);
            int64_t p = (int64_t)tile - 1;
            while (true) {
                int64_t idx = p - lane;
                unsigned long long s;
                do {
                    s = (idx >= 0) ? *((volatile unsigned long long*)&tile_state[idx]) : ST_PREFIX;
                } while (__any_sync(FULL, (s & ST_MASK) == 0));
                unsigned has_prefix = __ballot_sync(FULL, (s & ST_MASK) == ST_PREFIX);
                unsigned long long val = s & ~ST_MASK;
                if (has_prefix) {
                    int first = __ffs(has_prefix) - 1;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(FULL, val, d);
                excl += val;
                if (has_prefix) break;
                p -= 32;
            }
            if (lane == 0) atomicExch(&tile_state[tile], ST_PREFIX | (excl + tile_total));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    unsigned long long run = s_excl + warp_off + (x - tsum);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) {
            out[base + j] = (int64_t)run;
            run += v[j];
            if (base + j == n - 1) out[n] = (int64_t)run;
        }
    }
}

void launch_scan_u32_i64(const uint32_t* in, int64_t n, int64_t* out, uint64_t* tile_state,
                         PlanCounters* ctr, cudaStream_t s) {
    if (n <= 0) {
        cudaMemsetAsync(out, 0, sizeof(int64_t), s);
        return;
    }
    size_t tiles = (size_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    cudaMemsetAsync(tile_state, 0, tiles * sizeof(uint64_t), s);
    cudaMemsetAsync(&ctr->scan_ticket, 0, sizeof(uint32_t), s);
    k_scan_u32_i64<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, n, out, (unsigned long long*)tile_state,
                                                            &ctr->scan_ticket);
}

// ---------------------------------------------------------------------------------------
// The first pass writes finished rows to a scratch CSR laid out by product count (an upper bound of every
// row's nnz); after the row_ptr scan they are copied to their final place.  k_mask_sorted yields the
// per-row scratch sizes, k_copy_rows moves the rows.
__global__ void k_mask_sorted(const uint32_t* __restrict__ flops, int64_t m, uint32_t lo, uint32_t hi,
                              uint32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        uint32_t f = flops[i];
        out[i] = (f > lo && f <= hi) ? f : 0u;
    }
}
// rows with lo < products <= hi go through the scratch CSR
void launch_mask_sorted(const uint32_t* flops, int64_t m, uint32_t lo, uint32_t hi, uint32_t* out, cudaStream_t s) {
    if (m > 0) k_mask_sorted<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(flops, m, lo, hi, out);
}

// The copy out of the scratch CSR is also where the C shards of a multi-GPU run are exchanged: dst.col[d] / dst.val[d]
// are the C buffers of every GPU (this one first, the peers through NVLink peer mappings), so one read of the scratch
// row feeds the local store and the stores into every peer -- the all-gather of C fused into the kernel that writes C
// (SURVEY.md 8e "fusion candidate").  With a single destination it is a plain copy.
constexpr int COPY_WARPS = 8;
// one element of a row into every destination
template <int ND>
__device__ __forceinline__ void copy_store(const CopyDst& dst, int64_t o, int32_t c, double v) {
    if (ND == 1) {
        st_out(dst.col[0] + o, c);
        st_out(dst.val[0] + o, v);
    } else {
        // the peers first: their stores cross NVLink and take longest to drain
#pragma unroll 1
        for (int d = dst.n - 1; d >= 0; --d) {
            st_out(dst.col[d] + o, c);
            st_out(dst.val[d] + o, v);
        }
    }
}
// A row of n entries from scratch (src) to its place (o) by `width` threads (lane = this thread's index among them).
// The stores are what crosses NVLink in a sharded run, so they are the aligned side: a short head brings the
// destination to a multiple of 32 entries (128 bytes of column ids, 256 of values), after which every warp store is
// whole aligned lines instead of two partial ones.
template <int ND>
__device__ __forceinline__ void copy_row(const CopyDst& dst, int64_t o, const int32_t* __restrict__ t_col,
                                         const double* __restrict__ t_val, int64_t src, int64_t n, int lane, int width) {
    const int64_t head = (32 - (o & 31)) & 31;
    if (lane < head && lane < n) copy_store<ND>(dst, o + lane, t_col[src + lane], t_val[src + lane]);
    int64_t j = head + lane;
    for (; j + width < n; j += 2 * width) {   // two elements per thread in flight
        const int32_t c0 = t_col[src + j], c1 = t_col[src + j + width];
        const double v0 = t_val[src + j], v1 = t_val[src + j + width];
        copy_store<ND>(dst, o + j, c0, v0);
        copy_store<ND>(dst, o + j + width, c1, v1);
    }
    if (j < n) copy_store<ND>(dst, o + j, t_col[src + j], t_val[src + j]);
}

template <int ND>
__global__ void __launch_bounds__(COPY_WARPS * 32)
k_copy_rows(const uint32_t* __restrict__ flops, int64_t m, uint32_t lo, uint32_t hi, const int64_t* __restrict__ t_ptr,
            const int32_t* __restrict__ t_col, const double* __restrict__ t_val, const int64_t* __restrict__ c_ptr,
            CopyDst dst) {
    const int lane = lane_id();
    const int64_t r = (int64_t)blockIdx.x * COPY_WARPS + (threadIdx.x >> 5);
    if (r >= m) return;
    const uint32_t f = flops[r];
    if (f <= lo || f > hi) return;  // rows outside (lo, hi] are written by their own kernels
    const int64_t src = t_ptr[r], d0 = c_ptr[r];
    const int64_t n = c_ptr[r + 1] - d0;
    copy_row<ND>(dst, d0 + shard_offset(dst.off, dst.shard_nnz, dst.shard_idx), t_col, t_val, src, n, lane, 32);
}
void launch_copy_rows(const uint32_t* flops, int64_t m, uint32_t lo, uint32_t hi, const int64_t* t_ptr,
                      const int32_t* t_col, const double* t_val, const int64_t* c_ptr, const CopyDst& dst,
                      cudaStream_t s) {
    if (m <= 0) return;
    const unsigned grid = (unsigned)((m + COPY_WARPS - 1) / COPY_WARPS);
    if (dst.n == 1) k_copy_rows<1><<<grid, COPY_WARPS * 32, 0, s>>>(flops, m, lo, hi, t_ptr, t_col, t_val, c_ptr, dst);
    else k_copy_rows<8><<<grid, COPY_WARPS * 32, 0, s>>>(flops, m, lo, hi, t_ptr, t_col, t_val, c_ptr, dst);
}

// long scratch rows: one CTA per row of the list, a warp per row would leave a tail of 300-iteration warps behind
// the rest of the copy
template <int ND>
__global__ void __launch_bounds__(256)
k_copy_rows_list(const uint32_t* __restrict__ rows_list, const int64_t* __restrict__ t_ptr,
                 const int32_t* __restrict__ t_col, const double* __restrict__ t_val, const int64_t* __restrict__ c_ptr,
                 CopyDst dst) {
    const uint32_t r = rows_list ? rows_list[blockIdx.x] : blockIdx.x;
    const int64_t src = t_ptr[r], d0 = c_ptr[r];
    const int64_t n = c_ptr[r + 1] - d0;
    copy_row<ND>(dst, d0 + shard_offset(dst.off, dst.shard_nnz, dst.shard_idx), t_col, t_val, src, n, (int)threadIdx.x, 256);
}
void launch_copy_rows_list(const uint32_t* rows_list, uint32_t n_rows, const int64_t* t_ptr, const int32_t* t_col,
                           const double* t_val, const int64_t* c_ptr, const CopyDst& dst, cudaStream_t s) {
    if (n_rows == 0) return;
    if (dst.n == 1) k_copy_rows_list<1><<<n_rows, 256, 0, s>>>(rows_list, t_ptr, t_col, t_val, c_ptr, dst);
    else k_copy_rows_list<8><<<n_rows, 256, 0, s>>>(rows_list, t_ptr, t_col, t_val, c_ptr, dst);
}

// row pointers of a shard, shifted by the shard's global nnz offset, into the row_ptr of every GPU
__global__ void k_shift_row_ptr(const int64_t* __restrict__ ptr, int64_t m, int64_t off, const int64_t* shard_nnz,
                                int shard_idx, RowPtrDst dst, int64_t row_off) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > m) return;
    const int64_t v = ptr[r] + shard_offset(off, shard_nnz, shard_idx);
    for (int d = 0; d < dst.n; ++d) dst.ptr[d][row_off + r] = v;
}
void launch_shift_row_ptr(const int64_t* ptr, int64_t m, int64_t off, const int64_t* shard_nnz, int shard_idx,
                          const RowPtrDst& dst, int64_t row_off, cudaStream_t s) {
    k_shift_row_ptr<<<(unsigned)((m + 1 + 255) / 256), 256, 0, s>>>(ptr, m, off, shard_nnz, shard_idx, dst, row_off);
}

// ---------------------------------------------------------------------------------------
// Fiber store of a B operand (DevCsr::desc): rows re-laid on FIBER_PAD-element boundaries + one packed
// (start, length) descriptor per row.  HBM traffic: 12 B read + 12 B written per nonzero, once per operand.
__global__ void k_fiber_lengths(const int64_t* __restrict__ ptr, int64_t rows, uint32_t pad, uint32_t* __restrict__ out,
                                PlanCounters* ctr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) {
        const int64_t len = ptr[i + 1] - ptr[i];
        if (len >= (1ll << FIBER_LEN_BITS)) atomicAdd(&ctr->invalid_rows, 1u);   // no descriptor can hold it
        out[i] = (uint32_t)((len + pad - 1) / pad * pad);
    }
}
void launch_fiber_lengths(const int64_t* ptr, int64_t rows, uint32_t pad, uint32_t* padded_len, PlanCounters* ctr,
                          cudaStream_t s) {
    if (rows > 0) k_fiber_lengths<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(ptr, rows, pad, padded_len, ctr);
}

// 16 lanes per row; start == nullptr: descriptors into the canonical arrays (no copy)
__global__ void __launch_bounds__(256)
k_fiber_fill(DevCsr m, const int64_t* __restrict__ start, unsigned long long* __restrict__ desc,
             int32_t* __restrict__ gcol, double* __restrict__ gval) {
    const int64_t r = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 4;
    const int sub = threadIdx.x & 15;
    if (r >= m.rows) return;
    const int64_t s0 = m.ptr[r], len = m.ptr[r + 1] - s0;
    const int64_t d0 = start ? start[r] : s0;
    if (sub == 0) desc[r] = ((unsigned long long)d0 << FIBER_LEN_BITS) | (unsigned long long)len;
    if (start)
        for (int64_t j = sub; j < len; j += 16) {
            gcol[d0 + j] = m.col[s0 + j];
            gval[d0 + j] = m.val[s0 + j];
        }
}
void launch_fiber_fill(const DevCsr& m, const int64_t* start, unsigned long long* desc, int32_t* gcol, double* gval,
                       cudaStream_t s) {
    if (m.rows > 0) k_fiber_fill<<<(unsigned)((m.rows * 16 + 255) / 256), 256, 0, s>>>(m, start, desc, gcol, gval);
}

// ---------------------------------------------------------------------------------------
// operand conversion: the reference holds usize (u64) indices (storage.rs:150-160); the device
// layout is i64 row_ptr / i32 col (12 B per nonzero instead of 16).
__global__ void k_narrow_u64_i32(const uint64_t* __restrict__ src, int64_t n, int32_t* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (int32_t)src[i];
}
__global__ void k_widen_i32_i64(const int32_t* __restrict__ src, int64_t n, int64_t* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (int64_t)src[i];
}
__global__ void k_widen_i32_u64(const int32_t* __restrict__ src, int64_t n, uint64_t* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (uint64_t)(uint32_t)src[i];
}
static unsigned grid_for(int64_t n) {
    int64_t g = (n + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    if (g < 1) g = 1;
    return (unsigned)g;
}
void launch_widen_u64(const uint64_t*, int64_t, int64_t*, const uint64_t* src_idx, int64_t nnz,
                      int32_t* dst_idx, cudaStream_t s) {
    if (nnz > 0) k_narrow_u64_i32<<<grid_for(nnz), 256, 0, s>>>(src_idx, nnz, dst_idx);
}
void launch_widen_i32(const int32_t* src_ptr, int64_t n_ptr, int64_t* dst_ptr, cudaStream_t s) {
    if (n_ptr > 0) k_widen_i32_i64<<<grid_for(n_ptr), 256, 0, s>>>(src_ptr, n_ptr, dst_ptr);
}
void launch_narrow_result(const int64_t*, int64_t, uint64_t*, const int32_t* idx, int64_t nnz,
                          uint64_t* out_idx, cudaStream_t s) {
    if (nnz > 0) k_widen_i32_u64<<<grid_for(nnz), 256, 0, s>>>(idx, nnz, out_idx);
}

// ---------------------------------------------------------------------------------------
// canonical-CSR validation (SURVEY.md 8a edge cases: the reference's sort/merge units assume
// ascending unique columns, simulator.rs:28, adder_tree.rs:182).  nnz-parallel: every
// descent col[p] <= col[p-1] must coincide with the first stored element of a non-empty row.
__global__ void k_validate_nnz(DevCsr a, PlanCounters* ctr) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t bad = 0, descents = 0;
    for (; p < a.nnz; p += stride) {
        int32_t c = a.col[p];
        if (c < 0 || c >= a.cols) ++bad;
        if (p > 0 && c <= a.col[p - 1]) ++descents;
    }
    if (bad) atomicAdd(&ctr->invalid_rows, bad);
    if (descents) atomicAdd(&ctr->long_rows, descents);  // long_rows reused as the descent counter
}
__global__ void k_validate_rows(DevCsr a, PlanCounters* ctr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t bad = 0, starts = 0;
    for (; i < a.rows; i += stride) {
        int64_t s = a.ptr[i], e = a.ptr[i + 1];
        if (e < s || s < 0 || e > a.nnz) { ++bad; continue; }
        if (i == 0 && s != 0) ++bad;
        if (i == a.rows - 1 && e != a.nnz) ++bad;
        if (e > s && s > 0 && a.col[s] <= a.col[s - 1]) ++starts;
    }
    if (bad) atomicAdd(&ctr->invalid_rows, bad);
    if (starts) atomicAdd(&ctr->scan_ticket, starts);  // scan_ticket reused as the allowed-descent counter
}
void launch_validate(const DevCsr& a, PlanCounters* ctr, cudaStream_t s) {
    // caller zeroes *ctr first and afterwards checks invalid_rows == 0 && long_rows == scan_ticket
    if (a.nnz > 0) k_validate_nnz<<<grid_for(a.nnz), 256, 0, s>>>(a, ctr);
    if (a.rows > 0) k_validate_rows<<<grid_for(a.rows), 256, 0, s>>>(a, ctr);
}

}  // namespace spada
