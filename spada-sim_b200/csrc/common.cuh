// common.cuh -- shared device helpers and the launcher interface between the kernels (*.cu)
// and the host engine (engine.cu).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
//   bin 9          p <= 65536             one CTA (1024 threads) per row, bitmap + ranks in shared memory (B up to 2^21
//                                         columns wide; wider: joins bin 10)
//   bin 10         p  > 65536             items of ~8192 products over the grid, bitmap + ranks in HBM/L2
constexpr int NUM_BINS = 11;
constexpr int BIN_EMPTY = 0;
constexpr int BIN_HEAVY = 9;
constexpr int BIN_HUGE = 10;
constexpr uint32_t ESC_MAX_PRODUCTS = 4096;
constexpr uint32_t HEAVY_MAX_PRODUCTS = 65536;

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
    if (p <= 65536) return 9;
    return 10;
}
__host__ __device__ inline uint32_t bin_capacity(int b) { return b == 0 ? 0u : (b >= BIN_HEAVY ? 0u : (32u << (b - 1))); }

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
// one at a time and warp-uniformly, to on_long(b_row_start, length, a_val) -- the TMA-staged path of the
// huge bin hooks in here.
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
void launch_mask_sorted(const uint32_t* flops, int64_t m, uint32_t limit, uint32_t* out, cudaStream_t s);
void launch_copy_rows(const uint32_t* flops, int64_t m, uint32_t limit, const int64_t* t_ptr, const int32_t* t_col,
                      const double* t_val, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s);
void launch_copy_rows_list(const uint32_t* rows_list, uint32_t n_rows, const int64_t* t_ptr, const int32_t* t_col,
                           const double* t_val, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s);
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
// stage 2 / 3, ESC bins (1..8)
int esc_grid(int bin, uint32_t rows);
void launch_esc_symbolic(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                         uint32_t rows, uint32_t* row_nnz, cudaStream_t s);
// row_nnz_out (nullable): the kernel also records every row's nnz (first pass of the scratch mode)
void launch_esc_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                        uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                        uint32_t* row_nnz_out = nullptr);
// two-phase mode with kept keys: symbolic stores each row's sorted (column, arrival) keys at
// kstore + prod_ptr[row]; numeric reloads them instead of sorting again (bins 1..8)
void launch_esc_symbolic_keep(int bin, bool wide, const DevCsr& a, const DevCsr& b, int64_t row_begin,
                              const uint32_t* perm, uint32_t rows, uint32_t* row_nnz, const int64_t* prod_ptr,
                              void* kstore, cudaStream_t s);
void launch_esc_numeric_presorted(int bin, bool wide, const DevCsr& a, const DevCsr& b, int64_t row_begin,
                                  const uint32_t* perm, uint32_t rows, const int64_t* c_ptr, int32_t* c_col,
                                  double* c_val, const int64_t* prod_ptr, const void* kstore, cudaStream_t s);
void launch_cta_symbolic_keep(int bin, bool wide, const DevCsr& a, const DevCsr& b, int64_t row_begin,
                              const uint32_t* perm, uint32_t rows, uint32_t* row_nnz, const int64_t* prod_ptr,
                              void* kstore, cudaStream_t s);
void launch_cta_numeric_presorted(int bin, bool wide, const DevCsr& a, const DevCsr& b, int64_t row_begin,
                                  const uint32_t* perm, uint32_t rows, const int64_t* c_ptr, int32_t* c_col,
                                  double* c_val, const int64_t* prod_ptr, const void* kstore, cudaStream_t s);
bool esc_needs_wide_keys(int bin, int64_t b_cols);
// stage 2 / 3, heavy bin (9): one CTA per row, bitmap in shared memory (heavy_smem.cu)
void launch_heavy_smem_symbolic(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list,
                                uint32_t n_rows, uint32_t* row_nnz, cudaStream_t s);
// row_nnz_out (nullable): one-shot mode -- c_ptr addresses scratch rows of capacity >= nnz, nnz is recorded
void launch_heavy_smem_numeric(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list,
                               uint32_t n_rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                               uint32_t* row_nnz_out = nullptr);
// stage 2 / 3, huge bin (10): rows are cut into items (~8192 products) spread over the grid (heavy.cu)
struct HeavyPlan {
    uint32_t words;      // bitmap words per row = ceil(B.cols / 32)
    uint32_t wave_rows;  // heavy rows whose bitmaps fit the workspace at once
    uint32_t n_waves;
    uint64_t max_items;  // capacity of the item list
    size_t ws_words;     // uint2 words of workspace
};
HeavyPlan heavy_plan_sizes(uint32_t n_rows, uint64_t products, int64_t b_cols, size_t ws_budget_bytes);
void launch_heavy_items(const DevCsr& a, int64_t row_begin, const uint32_t* rows_list, uint32_t n_rows,
                        const uint32_t* flops, uint32_t* items_per_row, int64_t* item_off, uint32_t* item_row,
                        uint64_t* tile_state, PlanCounters* ctr, cudaStream_t s);
void launch_heavy_bits(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list,
                       const uint32_t* flops, const int64_t* item_off, const uint32_t* item_row, uint32_t wave_lo,
                       uint32_t wave_hi, uint2* ws, const HeavyPlan& P, int sm_count, cudaStream_t s);
void launch_heavy_rank(const uint32_t* rows_list, uint32_t wave_lo, uint32_t wave_hi, uint2* ws, const HeavyPlan& P,
                       uint32_t* row_nnz /* or NULL */, cudaStream_t s);
void launch_heavy_emit(const uint32_t* rows_list, uint32_t wave_lo, uint32_t wave_hi, const uint2* ws,
                       const HeavyPlan& P, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                       const uint32_t* row_nnz = nullptr);
void launch_heavy_accum(const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* rows_list,
                        const uint32_t* flops, const int64_t* item_off, const uint32_t* item_row, uint32_t wave_lo,
                        uint32_t wave_hi, const uint2* ws, const HeavyPlan& P, const int64_t* c_ptr, double* c_val,
                        int sm_count, cudaStream_t s);
// bitonic variant of the CTA-per-row bins 6..8 (esc_cta_bitonic.cu)
void launch_bitonic_cta_symbolic(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                 uint32_t rows, uint32_t* row_nnz, cudaStream_t s);
void launch_bitonic_cta_numeric(int bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, const uint32_t* perm,
                                uint32_t rows, const int64_t* c_ptr, int32_t* c_col, double* c_val, cudaStream_t s,
                                uint32_t* row_nnz_out = nullptr);
// stages 2+3+4 fused for the warp-per-row bins (fused.cu)
size_t fused_tile_state_words(int64_t m);
void launch_fused_light(int max_bin, const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t m,
                        const uint32_t* flops, const uint32_t* pre_nnz, int64_t* c_ptr, int32_t* c_col, double* c_val,
                        uint64_t* tile_state, PlanCounters* ctr, cudaStream_t s);
void setup_kernel_attributes();

}  // namespace spada
