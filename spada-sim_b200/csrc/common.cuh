// common.cuh -- shared device helpers and the launcher interface between the kernels (*.cu)
// and the host engine (engine.cu).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>

namespace spada {

constexpr unsigned FULL = 0xffffffffu;

// ---- bins ---------------------------------------------------------------------------
// Stage 1 classifies every A row by its intermediate-product count p (Spada's window
// choice restated for a GPU, rowwise_perf_adjust.rs:121-252 / scheduler.rs:729-753: R rows
// share a window of L lanes; here a bin fixes how many lanes cooperate on one row).
//   bin 0          p == 0                 nothing to do, nnz = 0
//   bin 1          p <= 32                four rows per warp (8 lanes x 4 keys each), products sorted in registers
//   bin 2..5       p <= 64,128,256,512    one warp per row, E = N/32 keys per lane
//   bin 6..8       p <= 1024,2048,4096    one CTA per row, register chunk sorts merged in shared memory
//   bin 8+L        p <= 4096 * 2^L        LONG rows (L = 1..20): K-tiled into chunks of 4096 products, every chunk sorted
//                                         by one CTA, the chunks merged pairwise in L levels (longrow.cu) -- the
//                                         reference's partial rows + merge tree (scheduler.rs:381-480, 820-920)
constexpr int NUM_BINS = 29;
constexpr int BIN_EMPTY = 0;
constexpr int BIN_LONG0 = 9;            // first long bin (two chunks, one merge level)
constexpr uint32_t ESC_MAX_PRODUCTS = 4096;
constexpr int LONG_UNIT_LOG = 12;
constexpr int LONG_UNIT = 1 << LONG_UNIT_LOG;   // products per chunk of a long row = outputs per merge tile

__host__ __device__ inline int bin_of(uint32_t p) {
    if (p == 0) return 0;
    if (p <= 32) return 1;
    if (p <= 64) return 2;
    if (p <= 128) return 3;
    if (p <= 256) return 4;
    if (p <= 512) return 5;
    if (p <= 1024) return 6;
    if (p <= 2048) return 7;
    if (p <= 4096) return 8;
    // L = ceil(log2(ceil(p / 4096))) merge levels
    uint32_t x = (p - 1) >> LONG_UNIT_LOG;   // >= 1
#ifdef __CUDA_ARCH__
    return 8 + (32 - __clz((int)x));
#else
    int L = 0;
    while (x) { ++L; x >>= 1; }
    return 8 + L;
#endif
}
__host__ __device__ inline int long_levels(uint32_t p) { return bin_of(p) - 8; }
__host__ __device__ inline uint64_t bin_capacity(int b) { return b == 0 ? 0ull : (32ull << (b - 1)); }

struct BinTable {
    uint32_t offset[NUM_BINS + 1];  // start of each bin inside perm[]
};

// counters written by stage 1, read back by the host in one copy
struct PlanCounters {
    unsigned long long total_products;
    unsigned long long bin_products[NUM_BINS];
    uint32_t bin_rows[NUM_BINS];
    uint32_t bin_cursor[NUM_BINS];
    uint32_t long_rows;      // rows deferred to the CTA-per-row flop counter
    uint32_t invalid_rows;   // validation failures
    uint32_t scan_ticket;    // dynamic tile id of the look-back scan
    uint32_t max_flops;      // largest per-row product count (saturates at 2^32-1)
    uint32_t tiny_fit;       // rows of bin 1 with at most 8 A entries: they fit an 8-lane group (window [4, 8])
    uint32_t pad;
};

// ---- device-side CSR view ------------------------------------------------------------
struct DevCsr {
    const int64_t* ptr;   // rows + 1
    const int32_t* col;   // nnz
    const double* val;    // nnz
    int64_t rows, cols, nnz;
    // B side only (nullable): the operand's FIBER STORE.  desc[k] = (start << 24) | length of row k inside col/val,
    // which then point at a copy whose rows start on 16-element boundaries (64-byte DRAM atoms for the column ids,
    // 128 bytes for the values): one 8-byte descriptor instead of two row_ptr reads per visited B row, and a short
    // B row costs one DRAM atom per array instead of the two an unaligned row straddles.  ptr stays the canonical
    // row_ptr (lengths only).  Built by launch_fiber_* (plan.cu) at upload time -- the device-side counterpart of
    // CsrMatStorage::init_with_gemm (storage.rs:214-239) laying B out for the fiber cache (storage.rs:460-).
    const unsigned long long* desc;
};
constexpr int FIBER_LEN_BITS = 24;   // rows of 2^24 or more elements: no fiber store (desc == nullptr)
constexpr int FIBER_PAD = 16;

#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}

__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
    int lo = __shfl_sync(FULL, (int)(v & 0xffffffffll), src);
    int hi = __shfl_sync(FULL, (int)(v >> 32), src);
    return ((int64_t)hi << 32) | (uint32_t)lo;
}
__device__ __forceinline__ double shfl_f64(double v, int src) {
    long long b = __double_as_longlong(v);
    return __longlong_as_double(shfl_i64(b, src));
}

// read-only, L1-bypassing streaming loads for operand arrays that are touched once per row
__device__ __forceinline__ int32_t ldg_i32(const int32_t* p) { return __ldg(p); }
__device__ __forceinline__ double ldg_f64(const double* p) { return __ldg(p); }
__device__ __forceinline__ int64_t ldg_i64(const int64_t* p) { return __ldg(p); }

// stores of finished rows (C and the scratch CSR are written once and not read again by the kernel that writes them):
// cache-streaming (evict-first) so that 3+ GB of output do not push B's rows and descriptors out of L2
#ifndef SPADA_STREAM_STORES
#define SPADA_STREAM_STORES 1
#endif
__device__ __forceinline__ void st_out(int32_t* p, int32_t v) {
    if (SPADA_STREAM_STORES) __stcs(p, v); else *p = v;
}
__device__ __forceinline__ void st_out(double* p, double v) {
    if (SPADA_STREAM_STORES) __stcs(p, v); else *p = v;
}

// where row k of B starts inside b.col / b.val and how long it is
__device__ __forceinline__ void b_row(const DevCsr& b, int32_t k, int64_t& bs, int& len) {
    if (b.desc) {
        const unsigned long long d = __ldg(b.desc + k);
        bs = (int64_t)(d >> FIBER_LEN_BITS);
        len = (int)(d & ((1ull << FIBER_LEN_BITS) - 1ull));
    } else {
        bs = ldg_i64(b.ptr + k);
        len = (int)(ldg_i64(b.ptr + k + 1) - bs);
    }
}

// ---- product expansion ---------------------------------------------------------------------
// A warp walks 32 A nonzeros (one per lane: lane l owns position p, if p < a_end), scans the
// B-row lengths and then deals the products to lanes round-robin, so B rows are read with
// consecutive lanes on consecutive elements (coalesced) whatever their length.  This is the
// reference's "window" of A scalars fanned out over the lanes (scheduler.rs:551-556,
// simulator.rs:728-757) with L = 32.
// seq = arrival index of the product inside the C row (ascending k, then B's stored order).
// BIG: B rows longer than 2^24 are streamed one at a time by the whole warp (keeps the 32-bit
// scan from overflowing); their seq is unused.
constexpr int EXPAND_BIG_LEN = 1 << 24;
#ifndef SPADA_EXPAND_UNROLL
#define SPADA_EXPAND_UNROLL 2
#endif
constexpr int EXPAND_UNROLL = SPADA_EXPAND_UNROLL;  // independent B gathers in flight per lane
// emit(seq, col, a_val, b_val): the B column id (and value when NUMERIC) are already loaded.  The loads
// of EXPAND_UNROLL consecutive steps are issued back to back before any of them is consumed, so a
// warp keeps several HBM/L2 round trips in flight instead of one.
// expand_batch_long: B rows of at least `long_len` elements are taken out of the dealt stream and handed,
// one at a time and warp-uniformly, to on_long(b_row_start, length, a_val) (rows too long for the 32-bit scan).
template <bool NUMERIC, bool BIG, bool LOAD_COL, typename F, typename L>
__device__ __forceinline__ void expand_batch_long(const DevCsr& a, const DevCsr& b, int64_t p, int64_t a_end,
                                                  int lane, int seq_base, int& batch_total, int long_len, F&& emit,
                                                  L&& on_long) {
    int64_t bs = 0;
    int len = 0;
    double av = 0.0;
    if (p < a_end) {
        int32_t k = ldg_i32(a.col + p);
        if (NUMERIC) av = ldg_f64(a.val + p);
        b_row(b, k, bs, len);
    }
    int big_len = 0;
    unsigned big = 0;
    if (BIG) {
        big = __ballot_sync(FULL, len >= long_len);
        if (len >= long_len) {
            big_len = len;
            len = 0;
        }
    }
    int total;
    int off = warp_excl_scan(len, lane, total);
    batch_total = total;
    if (total <= 32) {
        // short batch (tiny rows): one step, no unrolling overhead
        int j = 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            int o = __shfl_sync(FULL, off, j + s);
            if (o <= lane) j += s;
        }
        int oj = __shfl_sync(FULL, off, j);
        int64_t bsj = shfl_i64(bs, j);
        double aj = 0.0;
        if (NUMERIC) aj = shfl_f64(av, j);
        if (lane < total) {
            int64_t q = bsj + (lane - oj);
            emit(seq_base + lane, LOAD_COL ? (uint32_t)ldg_i32(b.col + q) : 0u, aj, NUMERIC ? ldg_f64(b.val + q) : 0.0);
        }
    } else
    for (int base = 0; base < total; base += 32 * EXPAND_UNROLL) {
        int64_t q[EXPAND_UNROLL];
        double aj[EXPAND_UNROLL];
#pragma unroll
        for (int u = 0; u < EXPAND_UNROLL; ++u) {
            q[u] = 0;
            aj[u] = 0.0;
            if (base + u * 32 < total) {  // warp-uniform
                int t = base + u * 32 + lane;
                int j = 0;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    int o = __shfl_sync(FULL, off, j + s);
                    if (o <= t) j += s;
                }
                int oj = __shfl_sync(FULL, off, j);
                int64_t bsj = shfl_i64(bs, j);
                if (NUMERIC) aj[u] = shfl_f64(av, j);
                q[u] = bsj + (t - oj);
            }
        }
        uint32_t c[EXPAND_UNROLL];
        double bv[EXPAND_UNROLL];
#pragma unroll
        for (int u = 0; u < EXPAND_UNROLL; ++u) {
            c[u] = 0;
            bv[u] = 0.0;
            if (base + u * 32 + lane < total) {
                if (LOAD_COL) c[u] = (uint32_t)ldg_i32(b.col + q[u]);
                if (NUMERIC) bv[u] = ldg_f64(b.val + q[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < EXPAND_UNROLL; ++u) {
            int t = base + u * 32 + lane;
            if (t < total) emit(seq_base + t, c[u], aj[u], bv[u]);
        }
    }
    if (BIG) {
        while (big) {
            int j = __ffs(big) - 1;
            big &= big - 1;
            int64_t bsj = shfl_i64(bs, j);
            int lj = __shfl_sync(FULL, big_len, j);
            double aj = 0.0;
            if (NUMERIC) aj = shfl_f64(av, j);
            on_long(bsj, lj, aj);
        }
    }
}

template <bool NUMERIC, bool BIG, bool LOAD_COL = true, typename F>
__device__ __forceinline__ void expand_batch(const DevCsr& a, const DevCsr& b, int64_t p, int64_t a_end,
                                             int lane, int seq_base, int& batch_total, F&& emit) {
    expand_batch_long<NUMERIC, BIG, LOAD_COL>(a, b, p, a_end, lane, seq_base, batch_total, EXPAND_BIG_LEN + 1, emit,
                                              [&](int64_t bsj, int lj, double aj) {
                                                  for (int t = lane; t < lj; t += 32)
                                                      emit(-1, LOAD_COL ? (uint32_t)ldg_i32(b.col + bsj + t) : 0u, aj,
                                                           NUMERIC ? ldg_f64(b.val + bsj + t) : 0.0);
                                              });
}

// (A variant that deals the products of a padded fiber store four at a time -- one LDG.E.128 for the column ids, two for
// the values, one entry search per four products -- was measured on B200 and removed: ER fused<256> 5.19 -> 5.53 ms,
// Poisson 1.34 -> 1.70 ms, rect's sort passes +-2 %: the kernels are bound by the sort network's issue slots and the wider
// loads cost registers, i.e. resident warps.)

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier, 1-D, global -> shared -------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#endif

// cudaFuncSetAttribute is per device: every launcher that needs more than 48 KB of dynamic shared memory raises the
// limit the first time it runs on a device (one handle per device, but several devices per process are allowed)
struct PerDeviceOnce {
    bool done[64] = {};
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

// ---- launchers (implemented in the .cu files, called by engine.cu) ------------------------
// stage 1
void launch_flops(const DevCsr& a, const int64_t* b_ptr, int64_t b_rows, uint32_t* b_len, int64_t row_begin, int64_t m,
                  uint32_t* flops, uint32_t* long_list, PlanCounters* ctr, cudaStream_t s);
void launch_bin_scatter(const uint32_t* flops, int64_t m, const BinTable& tbl, uint32_t* perm,
                        PlanCounters* ctr, cudaStream_t s);
// scratch CSR of the first pass: rows with lo < products <= hi get `products` slots, the others none
void launch_mask_sorted(const uint32_t* flops, int64_t m, uint32_t lo, uint32_t hi, uint32_t* out, cudaStream_t s);
// Scratch rows -> their final place.  dst[0..n_dst) are the C buffers of every GPU that holds a copy of C (this one
// first; peers are written through NVLink peer mappings: the all-gather of C fused into the store), dst_off = where
// this shard's first entry goes inside them.
struct CopyDst {
    int32_t* col[8];
    double* val[8];
    int n;
    int64_t off;                 // host-known part of the shard's offset
    const int64_t* shard_nnz;    // nullable, device: nnz of every shard (all-gathered); the shard's offset then also
    int shard_idx;               // includes shard_nnz[0 .. shard_idx) -- no host round trip between the two halves
};
#ifdef __CUDACC__
__device__ __forceinline__ int64_t shard_offset(int64_t off, const int64_t* shard_nnz, int shard_idx) {
    if (shard_nnz)
        for (int i = 0; i < shard_idx; ++i) off += shard_nnz[i];
    return off;
}
#endif
void launch_copy_rows(const uint32_t* flops, int64_t m, uint32_t lo, uint32_t hi, const int64_t* t_ptr,
                      const int32_t* t_col, const double* t_val, const int64_t* c_ptr, const CopyDst& dst,
                      cudaStream_t s);
void launch_copy_rows_list(const uint32_t* rows_list, uint32_t n_rows, const int64_t* t_ptr, const int32_t* t_col,
                           const double* t_val, const int64_t* c_ptr, const CopyDst& dst, cudaStream_t s);
// fiber store of a B operand (see DevCsr::desc): padded row lengths -> (scan) -> starts -> descriptors + aligned copy
void launch_fiber_lengths(const int64_t* ptr, int64_t rows, uint32_t pad, uint32_t* padded_len, PlanCounters* ctr,
                          cudaStream_t s);
void launch_fiber_fill(const DevCsr& m, const int64_t* start, unsigned long long* desc, int32_t* gcol, double* gval,
                       cudaStream_t s);
// device transpose (transpose.cu): stable LSD radix sort of the entry indices by column id
int transpose_passes(int64_t cols);
int64_t transpose_tiles(int64_t nnz);
void launch_entry_rows(const DevCsr& a, uint32_t* erow, uint32_t* col_count, cudaStream_t s);
void launch_radix_hist(const int32_t* keys, int64_t n, int shift, uint32_t* hist, cudaStream_t s);
void launch_radix_scatter(const int32_t* keys, const uint32_t* pay, int64_t n, int shift, const int64_t* offs,
                          int32_t* keys_out, uint32_t* pay_out, cudaStream_t s);
void launch_transpose_gather(const uint32_t* pay, const uint32_t* erow, const double* val, int64_t n, int32_t* t_col,
                             double* t_val, cudaStream_t s);
// stage 4
void launch_scan_u32_i64(const uint32_t* in, int64_t n, int64_t* out /* n+1 */, uint64_t* tile_state,
                         PlanCounters* ctr, cudaStream_t s);
size_t scan_tile_state_words(int64_t n);
// conversions / validation
void launch_widen_u64(const uint64_t* src_ptr, int64_t n_ptr, int64_t* dst_ptr, const uint64_t* src_idx,
                      int64_t nnz, int32_t* dst_idx, cudaStream_t s);
void launch_widen_i32(const int32_t* src_ptr, int64_t n_ptr, int64_t* dst_ptr, cudaStream_t s);
void launch_narrow_result(const int64_t* ptr, int64_t n_ptr, uint64_t* out_ptr, const int32_t* idx,
                          int64_t nnz, uint64_t* out_idx, cudaStream_t s);
void launch_validate(const DevCsr& a, PlanCounters* ctr, cudaStream_t s);
// stages 2+3 in one pass per row, sort bins (1..8): the finished row goes to c_ptr[row] (a scratch row of capacity >=
// nnz, or the row's final place) and its nnz is recorded in row_nnz_out (nullable)
int esc_grid(int bin, uint32_t rows);
void launch_esc_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                        uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                        uint32_t* row_nnz_out = nullptr);
// CTA-per-row bins 6..8 (esc_cta_bitonic.cu)
void launch_bitonic_cta_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                uint32_t* row_nnz_out = nullptr);
// long rows (bins >= BIN_LONG0), longrow.cu: K-tiled chunk sorts + merge levels + left-to-right sums, all in the
// oracle's ascending-k order (no atomics).  LongPlan = the tables of the whole long-row list; a WAVE = a contiguous slice
// of the list whose products fit the pong buffer.
struct LongPlan {
    const uint32_t* rows_list;   // perm[] slice of the long rows, ascending merge-level count
    uint32_t n_rows;
    uint64_t unit_bound;         // upper bound of all chunks
    uint32_t* p;                 // [n_rows]   products per row
    uint32_t* u;                 // [n_rows]   chunks per row
    int64_t* prod_off;           // [n_rows+1] prefix of p (origin of a row inside its wave's pong buffer)
    int64_t* unit_off;           // [n_rows+1] first chunk of the row
    uint32_t* unit_row;          // [unit_bound] row (index inside the list) of every chunk / tile
    uint32_t* unit_heads;        // [unit_bound] distinct columns that start inside the tile
    int64_t* unit_hoff;          // [unit_bound+1]
    const int64_t* t_ptr;        // scratch CSR: a long row's scratch row holds its sorted products (buffer 0)
    int32_t* s_col;
    double* s_val;
    int32_t* pong_col;           // buffer 1: one wave
    double* pong_val;
    void* tiles;                 // [wave units] x 32 B: merge-tile descriptors of the level in flight
    uint64_t* tile_state;        // look-back state of the scans
};
struct LongWaveRange {
    uint32_t lo, hi;             // slice of the list
    uint32_t level_lo[24];       // first list index that takes part in merge level l (1-based)
    uint32_t level_grid[24];     // upper bound of the merge tiles of level l
    uint64_t level_products[24]; // upper bound of the products that take part in level l
    int max_level;
    uint64_t unit_bound;         // upper bound of the wave's chunks
    uint64_t products_bound;     // upper bound of the wave's products
};
// aseq[e] (one u32 per nonzero of A, written for long rows only) = arrival number of the first product of entry e
void launch_long_prefix(const DevCsr& a, int64_t row_begin, const uint32_t* rows_list, uint32_t n_rows,
                        const uint32_t* b_len, uint32_t* aseq, cudaStream_t s);
// stages (nullable): called around the sort, the merge levels and the count so the engine can time them.  All return
// the number of kernels launched.
struct LongStages {
    std::function<void(const char*, uint32_t, uint64_t)> on;   // stage name, grid, products it touches
    std::function<void()> off;
};
struct CopyDst;
uint32_t launch_long_setup(const LongPlan& P, const uint32_t* flops, PlanCounters* ctr, cudaStream_t s);
uint32_t launch_long_wave(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* aseq, const LongPlan& P,
                          const LongWaveRange& w, uint32_t* row_nnz, cudaStream_t s, const LongStages* stages = nullptr);
uint32_t launch_long_heads_scan(const LongPlan& P, PlanCounters* ctr, cudaStream_t s);
uint32_t launch_long_reduce(const LongPlan& P, const int64_t* c_ptr, const CopyDst& dst, cudaStream_t s);
// long rows of a product with few columns (B.cols <= DENSE_MAX_COLS): dense accumulator in shared memory, B rows applied
// one after the other in ascending k (longrow.cu)
constexpr int64_t DENSE_MAX_COLS = 16384;
bool dense_rows_fit(int64_t b_cols);
void launch_dense_rows(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list, uint32_t n_rows,
                       const int64_t* t_ptr, int32_t* t_col, double* t_val, uint32_t* row_nnz, cudaStream_t s);
// shard row pointers shifted by the shard's global offset, written into every GPU's row_ptr (sharded runs)
struct RowPtrDst {
    int64_t* ptr[8];
    int n;
};
void launch_shift_row_ptr(const int64_t* ptr, int64_t m, int64_t off, const int64_t* shard_nnz, int shard_idx,
                          const RowPtrDst& dst, int64_t row_off, cudaStream_t s);
// stages 2+3+4 fused for the warp-per-row bins (fused.cu)
size_t fused_tile_state_words(int64_t m);
// single pass for mixed row lengths: tiles cut by sort-slot budget (fused.cu: k_tile_pass)
size_t tile_pass_bound(int64_t m, uint64_t light_slots);
int tile_pass_cut();
void launch_tile_pass(const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t m, const uint32_t* flops,
                      const uint32_t* pre_nnz, int64_t* c_ptr, int32_t* c_col, double* c_val, uint32_t* w, int64_t* lp,
                      int64_t* tidx, uint32_t* tile_start, uint64_t* tile_state, size_t bound, uint64_t* scan_state,
                      PlanCounters* ctr, cudaStream_t s);
// tiny_quad: rows of bin 1 run four to a warp (window [4, 8]) instead of one per warp ([1, 32])
void launch_fused_light(int max_bin, bool tiny_quad, const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t m,
                        const uint32_t* flops, const uint32_t* pre_nnz, int64_t* c_ptr, int32_t* c_col, double* c_val,
                        uint64_t* tile_state, PlanCounters* ctr, cudaStream_t s);

}  // namespace spada
