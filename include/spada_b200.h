/*
 * spada_b200.h -- C ABI of the B200-native SpGEMM engine that replaces the functional
 * core of tsinghua-ideal/spada-sim (C = A x B over the simulator's CSR types).
 *
 * This is the drop-in boundary: plain C, pointers and sizes only, no C++/torch types, no
 * exceptions across the boundary.  The reference has no FFI seam of its own (single Rust
 * binary); the seam this header cuts is  src/main.rs:74-100  of the reference:
 *
 *     Simulator::new(.., &mut dram_a, &mut dram_b, &mut dram_psum, ..)   simulator.rs:431-507
 *     cycle_simu.execute()                                                simulator.rs:509-890
 *     cycle_simu.get_exec_result() -> Vec<CsrRow>                         simulator.rs:1034-1062
 *
 * whose inputs are CsrMatStorage{data: Vec<f64>, indptr: Vec<usize>, indices: Vec<usize>}
 * (storage.rs:150-160, built by init_with_gemm, storage.rs:214-239).  The Rust-side binding
 * a maintainer adds is spada-sim_b200/rust/spada-b200-sys.
 *
 * Contract: inputs are canonical CSR (ascending unique column ids per
 * row), f64 values, any m, k, n.  Output is canonical CSR with *structural* nnz (explicit or
 * cancelled zeros are kept: simulator.rs:209-221, adder_tree.rs:73-83), empty rows allowed
 * (simulator.rs:1037-1044).  Every product is one rounded f64 multiply followed by separate
 * rounded adds -- no FMA (simulator.rs:101, :217).
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point returns
 * SPADA_B200_NO_DEVICE.
 *
 * Threading: a handle is not thread-safe (mirrors the reference's `&mut self`); calls are
 * synchronous (they return after the result is resident on the device and the stream is
 * synchronised).  Errors: 0 = OK, otherwise a spada_b200_status; the message of the last
 * failure on the calling thread is available from spada_b200_last_error().
 */
#ifndef SPADA_B200_H
#define SPADA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SPADA_B200_API __attribute__((visibility("default")))
#else
#define SPADA_B200_API
#endif

#define SPADA_B200_ABI_VERSION 2
#define SPADA_B200_MAX_BINS 32
#define SPADA_B200_MAX_LAUNCHES 64

typedef enum spada_b200_status {
    SPADA_B200_OK = 0,
    SPADA_B200_INVALID_ARG = 1,
    SPADA_B200_UNSORTED_INPUT = 2, /* a row with non-ascending / duplicate / out-of-range columns */
    SPADA_B200_DIM_MISMATCH = 3,   /* A.cols != B.rows (the reference panics: scheduler.rs:654) */
    SPADA_B200_CUDA_ERROR = 4,
    SPADA_B200_NCCL_ERROR = 5,
    SPADA_B200_OOM = 6,
    SPADA_B200_NO_DEVICE = 7,
    SPADA_B200_TOO_LARGE = 8       /* a dimension >= 2^31 (device column ids are i32) */
} spada_b200_status;

/* Accelerator argument of the reference CLI (frontend.rs:33-41).  It selects the window
 * policy (scheduler.rs:729-753): how many A rows share one CTA / warp ("R" of the window
 * shape [R, lane_num/R]).  It never changes C. */
typedef enum spada_b200_accelerator {
    SPADA_B200_ACC_IP = 0,       /* row-wise:    [1, L]       -> one row per cooperative group */
    SPADA_B200_ACC_OP = 1,       /* column-wise: [L, 1]       */
    SPADA_B200_ACC_MULTIROW = 2, /* fixed [block_shape[0], L/block_shape[0]] */
    SPADA_B200_ACC_SPADA = 3     /* adaptive: rows binned by intermediate-product count */
} spada_b200_accelerator;

/* Host CSR exactly as the reference holds it: Vec<usize> / Vec<usize> / Vec<f64>
 * (storage.rs:150-160; usize == uint64_t on x86-64).  Borrowed for the duration of a call. */
typedef struct spada_csr_view {
    uint64_t rows, cols, nnz;
    const uint64_t *indptr;  /* rows + 1 */
    const uint64_t *indices; /* nnz column ids */
    const double *data;      /* nnz */
} spada_csr_view;

/* Host CSR as scipy hands it to the reference's loaders before pyo3 widens it
 * (py2rust.rs:77-80, gemm.rs:18-24): int32 indptr / indices, float64 data. */
typedef struct spada_csr_view32 {
    uint64_t rows, cols, nnz;
    const int32_t *indptr;
    const int32_t *indices;
    const double *data;
} spada_csr_view32;

typedef struct spada_b200_opts {
    int32_t device;          /* CUDA device ordinal, -1 = current device */
    int32_t accelerator;     /* spada_b200_accelerator */
    uint32_t lane_num;       /* OmegaConfig.lane_num (frontend.rs:13); 0 = 8 */
    uint32_t block_shape[2]; /* OmegaConfig.block_shape (frontend.rs:16) */
    uint32_t flags;          /* SPADA_B200_FLAG_* */
    void *stream;            /* cudaStream_t to launch on; NULL = engine-owned stream */
} spada_b200_opts;

#define SPADA_B200_FLAG_VALIDATE 1u  /* check canonical CSR on upload (UNSORTED_INPUT) */
#define SPADA_B200_FLAG_TWO_PHASE 2u /* always take the scratch pass: every row sorted and summed once into a scratch CSR,
                                        then placed into an exact-size C */
#define SPADA_B200_FLAG_SINGLE_PASS 4u /* always do rows with <= 512 products in one fused pass (C is then
                                          allocated with capacity = intermediate-product count).  With neither
                                          flag the engine picks: single pass when one bin holds >= 80 % of the rows */

#define SPADA_B200_FLAG_SERIAL 8u     /* every kernel on the one stream (the engine otherwise runs the long rows, > 4096
                                          products, on a side stream beside the sort bins): clean per-launch times */

typedef struct spada_b200 spada_b200_t;               /* engine handle (streams, workspace pool) */
typedef struct spada_b200_csr spada_b200_csr_t;       /* device-resident operand */
typedef struct spada_b200_result spada_b200_result_t; /* device-resident C (engine owned) */
typedef struct spada_b200_shard spada_b200_shard_t;   /* a row shard's product between its two halves */
typedef struct spada_b200_cbuf spada_b200_cbuf_t;     /* full-size C buffers (row_ptr, col, val) on one GPU */
typedef struct spada_b200_group spada_b200_group_t;   /* the GPUs of one process, one engine handle each */
#define SPADA_B200_IPC_HANDLE_BYTES 64                /* sizeof(cudaIpcMemHandle_t) */

typedef struct spada_b200_launch {
    char name[32];     /* kernel family + bin, e.g. "esc_numeric<256>" */
    float ms;          /* CUDA-event duration on the engine stream */
    uint32_t grid;     /* CTAs launched */
    uint64_t rows;     /* rows the launch processed */
    uint64_t products; /* intermediate products of those rows */
    uint64_t nnz;      /* output nnz of those rows (0 for the flop-count pass) */
} spada_b200_launch;

typedef struct spada_b200_stats {
    uint64_t rows, cols, nnz_a, nnz_b, products, nnz_c;
    uint64_t bin_rows[SPADA_B200_MAX_BINS];     /* rows per bin (bin 0 = rows without products) */
    uint64_t bin_products[SPADA_B200_MAX_BINS]; /* products per bin */
    uint32_t bin_window_rows[SPADA_B200_MAX_BINS]; /* R: A rows sharing one CTA in that bin */
    uint32_t bin_window_lanes[SPADA_B200_MAX_BINS]; /* lanes cooperating on one row */
    float ms_total;    /* first kernel start -> last kernel end (device) */
    float ms_flops;    /* stage 1: flop count + binning */
    float ms_symbolic; /* stage 2 */
    float ms_scan;     /* stage 4 (runs between 2 and 3) */
    float ms_numeric;  /* stage 3 */
    float ms_h2d;      /* host-level entry points only */
    float ms_d2h;
    uint32_t n_launches; /* kernels launched by the last spgemm */
    uint32_t n_recorded; /* entries valid in launches[] */
    spada_b200_launch launches[SPADA_B200_MAX_LAUNCHES];
} spada_b200_stats;

/* ---- library ------------------------------------------------------------------------ */
SPADA_B200_API int spada_b200_abi_version(void);
SPADA_B200_API const char *spada_b200_last_error(void);
SPADA_B200_API int spada_b200_device_count(int *count);
/* pinned host memory for callers that want full PCIe speed (bench.py's e2e leg) */
SPADA_B200_API int spada_b200_host_alloc(void **ptr, size_t bytes);
SPADA_B200_API int spada_b200_host_free(void *ptr);

/* ---- handle: replaces Simulator::new (simulator.rs:431-507) --------------------------- */
SPADA_B200_API int spada_b200_create(const spada_b200_opts *opts, spada_b200_t **out);
SPADA_B200_API void spada_b200_destroy(spada_b200_t *h);
SPADA_B200_API int spada_b200_set_stream(spada_b200_t *h, void *cuda_stream);
SPADA_B200_API int spada_b200_synchronize(spada_b200_t *h);
/* return pooled device memory to the driver */
SPADA_B200_API int spada_b200_trim(spada_b200_t *h);

/* ---- operands: replaces CsrMatStorage::init_with_gemm (storage.rs:214-239) ------------ */
SPADA_B200_API int spada_b200_upload(spada_b200_t *h, const spada_csr_view *m, spada_b200_csr_t **out);
SPADA_B200_API int spada_b200_upload32(spada_b200_t *h, const spada_csr_view32 *m, spada_b200_csr_t **out);
/* wrap device arrays the caller owns (i64 row_ptr, i32 col, f64 val); not freed by csr_free */
SPADA_B200_API int spada_b200_csr_wrap_device(spada_b200_t *h, uint64_t rows, uint64_t cols, uint64_t nnz,
                               const int64_t *d_indptr, const int32_t *d_indices,
                               const double *d_data, spada_b200_csr_t **out);
/* Builds the operand's fiber store: one packed (start, length) descriptor per row and, when rows average >= 6
 * nonzeros, a copy whose rows start on 64-byte boundaries -- the layout the kernels gather B rows from (the
 * engine-side counterpart of CsrMatStorage::init_with_gemm laying B out for the fiber cache, storage.rs:214-239,
 * 460-).  Uploaded operands get it automatically the first time they are used as B; call this for wrapped device
 * arrays (it snapshots them: call again after changing them -- free and re-wrap).  ms_or_null: device time. */
SPADA_B200_API int spada_b200_csr_prepare(spada_b200_t *h, spada_b200_csr_t *m, float *ms_or_null);
/* Marks an operand as used once: no fiber store is built for it (re-laying B costs about as much as one product saves:
 * rect 3 ms of build against 0.1 ms of kernel time), the kernels gather its rows through row_ptr. */
SPADA_B200_API int spada_b200_csr_set_one_shot(spada_b200_csr_t *m);
/* B = A^T as a new device operand (canonical CSR: ascending column ids = A's row ids).  Replaces the host-side
 * transpose of GEMM::from_mat for non-square SS workloads (gemm.rs:41-53: `transpose_into().to_csr()`); values are
 * moved, not computed, so the result is bit-identical to the reference's / scipy's. */
SPADA_B200_API int spada_b200_transpose(spada_b200_t *h, const spada_b200_csr_t *a, spada_b200_csr_t **out);
SPADA_B200_API int spada_b200_csr_shape(const spada_b200_csr_t *m, uint64_t *rows, uint64_t *cols, uint64_t *nnz);
SPADA_B200_API int spada_b200_csr_device_ptrs(const spada_b200_csr_t *m, const int64_t **d_indptr,
                               const int32_t **d_indices, const double **d_data);
/* device operand -> caller-allocated host arrays (int64 indptr [rows+1], int32 indices, float64 data) */
SPADA_B200_API int spada_b200_csr_download32(const spada_b200_csr_t *m, int64_t *indptr, int32_t *indices, double *data);
SPADA_B200_API void spada_b200_csr_free(spada_b200_csr_t *m);

/* ---- the hot path: replaces Simulator::execute + get_exec_result ---------------------- */
/* C[row_begin:row_end, :] = A[row_begin:row_end, :] x B with operands resident on the device.
 * row_end == UINT64_MAX means A.rows.  The result's row_ptr starts at 0. */
SPADA_B200_API int spada_b200_spgemm_dev(spada_b200_t *h, const spada_b200_csr_t *a, const spada_b200_csr_t *b,
                          uint64_t row_begin, uint64_t row_end, spada_b200_result_t **out);
/* host in, device result: upload + spgemm_dev (what the Rust wrapper calls) */
SPADA_B200_API int spada_b200_spgemm(spada_b200_t *h, const spada_csr_view *a, const spada_csr_view *b,
                      spada_b200_result_t **out);
SPADA_B200_API int spada_b200_spgemm32(spada_b200_t *h, const spada_csr_view32 *a, const spada_csr_view32 *b,
                        spada_b200_result_t **out);

/* per-row intermediate-product counts (stage 1 alone) and the balanced row split used to
 * shard A over n_shards GPUs: bounds[0..n_shards], bounds[0]=0, bounds[n]=A.rows, equal
 * product count per shard. */
SPADA_B200_API int spada_b200_flops(spada_b200_t *h, const spada_b200_csr_t *a, const spada_b200_csr_t *b,
                     uint64_t *total_products, uint64_t *host_flops_or_null);
SPADA_B200_API int spada_b200_plan_shards(spada_b200_t *h, const spada_b200_csr_t *a, const spada_b200_csr_t *b,
                           uint32_t n_shards, uint64_t *bounds);

/* ---- sharded runs: A row-sharded by equal product count over several GPUs, B replicated,
 * C gathered on every GPU.  The reference has no counterpart (single thread); each C row is an independent unit
 * there too (every window row writes its own psum address, scheduler.rs:548-550), which is what makes rows the
 * shard unit.  Every GPU owns a full-size set of C buffers; the second half of a shard's product stores the shard's
 * rows into ALL of them at the shard's global offset -- its own through HBM, the peers' through NVLink peer
 * mappings -- so the all-gather of C is fused into the kernel that writes C.
 *
 * One process per GPU (torchrun):  cbuf_create -> cbuf_export -> (handles exchanged by the launcher's collective)
 * -> cbuf_import of every peer; per product: shard_begin (first pass into scratch rows, local row_ptr; the shard's
 * nnz is left in d_nnz_local on the device) -> all-gather of the per-shard nnz (8 bytes per rank, NCCL) ->
 * shard_finish with the gathered device array (no host round trip in between) -> a barrier before C is read.
 * One handle carries one shard at a time. */
SPADA_B200_API int spada_b200_cbuf_create(spada_b200_t *h, uint64_t rows, uint64_t cols, uint64_t capacity_nnz,
                           spada_b200_cbuf_t **out);
/* handles: 3 x SPADA_B200_IPC_HANDLE_BYTES (row_ptr, col, val), valid in other processes of this node */
SPADA_B200_API int spada_b200_cbuf_export(const spada_b200_cbuf_t *c, void *handles);
SPADA_B200_API int spada_b200_cbuf_import(spada_b200_t *h, const void *handles, uint64_t rows, uint64_t cols,
                           uint64_t capacity_nnz, spada_b200_cbuf_t **out);
SPADA_B200_API void spada_b200_cbuf_free(spada_b200_cbuf_t *c);
SPADA_B200_API int spada_b200_cbuf_device_ptrs(const spada_b200_cbuf_t *c, const int64_t **d_indptr,
                                const int32_t **d_indices, const double **d_data);
SPADA_B200_API int spada_b200_cbuf_nnz(const spada_b200_cbuf_t *c, uint64_t *nnz); /* row_ptr[rows] */
/* gathered C -> caller-allocated host arrays (int64 indptr [rows+1], int32 indices, float64 data of cbuf_nnz entries) */
SPADA_B200_API int spada_b200_cbuf_copy32(const spada_b200_cbuf_t *c, int64_t *indptr, int32_t *indices, double *data);
/* first half.  d_nnz_local (nullable, device): receives the shard's nnz; nnz_local (nullable, host): the same after
 * a stream synchronisation. */
SPADA_B200_API int spada_b200_shard_begin(spada_b200_t *h, const spada_b200_csr_t *a, const spada_b200_csr_t *b,
                           uint64_t row_begin, uint64_t row_end, int64_t *d_nnz_local, uint64_t *nnz_local,
                           spada_b200_shard_t **out);
/* second half (consumes the shard, also on failure).  bufs[0] = this GPU's buffers, bufs[1..n) = the peers'.  The shard's
 * first entry goes to nnz_offset + sum(d_shard_nnz[0 .. shard_index)) (d_shard_nnz: nullable device array). */
SPADA_B200_API int spada_b200_shard_finish(spada_b200_shard_t *s, spada_b200_cbuf_t *const *bufs, uint32_t n_bufs,
                            uint64_t nnz_offset, const int64_t *d_shard_nnz, uint32_t shard_index,
                            spada_b200_stats *stats_or_null);
SPADA_B200_API void spada_b200_shard_abort(spada_b200_shard_t *s);

/* All GPUs of ONE process -- what the spada-sim CLI (a single process, main.rs:30-121) drives: n_gpus engine handles
 * with peer access between every pair; a product uploads the operands to device 0, replicates them over NVLink,
 * shards A by equal product count, runs the two halves with one host thread per device and returns device 0's
 * gathered C.  Replaces Simulator::new / execute / get_exec_result at main.rs:74-100 for n_gpus > 1. */
SPADA_B200_API int spada_b200_group_create(const spada_b200_opts *opts, uint32_t n_gpus, spada_b200_group_t **out);
SPADA_B200_API int spada_b200_group_spgemm(spada_b200_group_t *g, const spada_csr_view *a, const spada_csr_view *b,
                            spada_b200_result_t **out);
SPADA_B200_API int spada_b200_group_spgemm32(spada_b200_group_t *g, const spada_csr_view32 *a, const spada_csr_view32 *b,
                              spada_b200_result_t **out);
SPADA_B200_API void spada_b200_group_destroy(spada_b200_group_t *g);

/* ---- row-panel streaming: the step after the path (main.rs:113-116 prints ten rows of C and drops the rest) --------
 * C is computed in panels of consecutive rows (at most panel_products intermediate products each; 0 = engine's choice);
 * every finished panel is copied to pinned host memory while the next one is computed and handed to `sink` in row
 * order.  The arrays a sink receives are valid only during the call: indptr has row_end - row_begin + 1 entries holding
 * GLOBAL offsets (indptr[0] == nnz_begin), indices / data hold the panel's entries.  A non-zero return aborts the run.
 * C never has to fit in HBM, nor to be kept at all (a checksum sink stores nothing). */
typedef int (*spada_b200_panel_sink)(void *user, uint64_t row_begin, uint64_t row_end, uint64_t nnz_begin,
                                     const int64_t *indptr, const int32_t *indices, const double *data);
typedef struct spada_b200_stream_stats {
    uint64_t panels, products, nnz_c, max_panel_products;
    float ms_total; /* first panel's first kernel -> last panel handed over */
} spada_b200_stream_stats;
SPADA_B200_API int spada_b200_spgemm_stream(spada_b200_t *h, const spada_b200_csr_t *a, const spada_b200_csr_t *b,
                             uint64_t panel_products, spada_b200_panel_sink sink, void *user,
                             spada_b200_stream_stats *stats_or_null);

/* The same panels copied straight into caller-allocated host arrays at their global offsets (pin them for full PCIe
 * speed): int64 indptr [A.rows + 1], int32 indices / float64 data [capacity_nnz >= nnz(C), e.g. the product count of
 * spada_b200_flops].  What the Rust wrapper calls when it wants the whole of C on the host: the D2H copy of a panel
 * runs beside the computation of the next one. */
SPADA_B200_API int spada_b200_spgemm_to_host(spada_b200_t *h, const spada_b200_csr_t *a, const spada_b200_csr_t *b,
                              uint64_t panel_products, int64_t *indptr, int32_t *indices, double *data,
                              uint64_t capacity_nnz, spada_b200_stream_stats *stats_or_null);

/* Host operands in, whole C in caller-allocated host arrays -- one call for Simulator::new + execute + get_exec_result
 * (main.rs:74-100) that overlaps everything PCIe allows: B is uploaded first (every row of C needs all of it), A's
 * entries follow in row panels while the panels already on the device are computed and their results travel down.
 * Arrays as for spada_b200_spgemm_to_host; pinned host memory on both sides for full PCIe speed.  A non-canonical A is
 * reported (UNSORTED_INPUT) after the run, when all of it is on the device. */
SPADA_B200_API int spada_b200_spgemm32_host_to_host(spada_b200_t *h, const spada_csr_view32 *a, const spada_csr_view32 *b,
                                     int64_t *indptr, int32_t *indices, double *data, uint64_t capacity_nnz,
                                     spada_b200_stream_stats *stats_or_null);

/* ---- results: replaces get_exec_result (simulator.rs:1034-1062) ----------------------- */
SPADA_B200_API int spada_b200_result_shape(const spada_b200_result_t *r, uint64_t *rows, uint64_t *cols,
                            uint64_t *nnz);
/* caller-allocated outputs sized from result_shape; usize layout for CsrRow::new_from_data */
SPADA_B200_API int spada_b200_result_copy(const spada_b200_result_t *r, uint64_t *indptr, uint64_t *indices,
                           double *data);
/* device-native layout: int64 indptr (rows+1), int32 indices, float64 data */
SPADA_B200_API int spada_b200_result_copy32(const spada_b200_result_t *r, int64_t *indptr, int32_t *indices,
                             double *data);
SPADA_B200_API int spada_b200_result_device_ptrs(const spada_b200_result_t *r, const int64_t **d_indptr,
                                  const int32_t **d_indices, const double **d_data);
SPADA_B200_API int spada_b200_result_stats(const spada_b200_result_t *r, spada_b200_stats *out);
SPADA_B200_API void spada_b200_result_free(spada_b200_result_t *r);

#ifdef __cplusplus
}
#endif
#endif /* SPADA_B200_H */
