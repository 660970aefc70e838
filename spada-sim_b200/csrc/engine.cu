// engine.cu -- host side of the engine and the C ABI of include/spada_b200.h.
//
// One handle owns a device, a stream (plus a side stream that is always joined back into it) and a
// caching device-memory pool (freed blocks are kept and handed out again by size, so steady-state calls
// never reach the driver allocator; everything is ordered on the one stream by the time a block is freed,
// which makes immediate reuse safe).  A call to spada_b200_spgemm_dev runs the stages of the path:
//   1. flop count + binning        (plan.cu)   -- one host read-back of ~300 bytes of counters
//   then, picked per operand ("Engine modes"):
//   single pass   rows <= 512 products expanded, sorted, reduced and placed by a look-back scan in ONE kernel
//                 (fused.cu); heavier rows: first pass into scratch rows before it, copied into place after it
//   two phase     2. one pass per bin into a scratch CSR sized by product count (esc.cu, esc_cta_bitonic.cu;
//                    long rows: chunk sorts + merge levels, longrow.cu)
//                 4. exclusive scan -> row_ptr (plan.cu)   -- one host read-back of nnz(C) to size C
//                 3. copy of the scratch rows into C (and, for sharded runs, into every peer's C)
// There is no CPU compute path here: if no CUDA device is present every entry point fails with
// SPADA_B200_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <unordered_map>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/spada_b200.h"
#include "common.cuh"

using namespace spada;

// Small device -> host read-backs (stage-1 counters, nnz(C)) are written by a one-warp kernel straight into pinned host
// memory instead of going through cudaMemcpyAsync: a copy-engine transfer queues behind whatever large D2H copy is in
// flight on another stream (the row-panel pipeline, a caller's own copies) and would stall the product for its whole
// duration -- measured on the rect config: row panels 75 ms with copy-engine read-backs against the 58 ms the overlap
// allows.
void launch_publish(const void* d_src, void* h_dst, size_t bytes, cudaStream_t s);

// ---- errors -------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            int code__ = (e__ == cudaErrorMemoryAllocation) ? SPADA_B200_OOM                      \
                         : (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver)       \
                             ? SPADA_B200_NO_DEVICE                                               \
                             : SPADA_B200_CUDA_ERROR;                                             \
            return fail(code__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                \
        }                                                                                         \
    } while (0)

// ---- objects ------------------------------------------------------------------------------
struct spada_b200 {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // Side stream: the long rows (chunk sorts + merge levels, many small launches) are independent of the sort bins
    // until the row_ptr scan and run beside them, joined back into `stream` by an event.  SPADA_B200_FLAG_SERIAL puts
    // everything on the one stream, which is what the per-launch event times of the stats are meaningful for.
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int fiber_pad = -1;            // SPADA_B200_FIBER_PAD: -1 auto (16 when rows average >= 6 nonzeros, else descriptors
                                   // only), 0 no fiber store, 1 descriptors only, 16 always pad
    size_t long_ws_budget = (size_t)24 << 30;   // ping-pong buffers of one wave of long rows (SPADA_B200_LONG_WS_MB)
    bool tile_pass = false;        // SPADA_B200_TILE_PASS=1: mixed row lengths in one pass with work-cut tiles (k_tile_pass).
                                   // Off by default: measured on the power-law config 4.6-6.2 ms for the tile pass (four
                                   // shapes of tile) against 2.2 ms of sort passes + 0.6 ms of copy for the same rows
    spada_b200_opts opts{};
    PlanCounters* d_ctr = nullptr;
    PlanCounters* d_ctr_side = nullptr;   // scan tickets of the side stream
    PlanCounters* h_ctr = nullptr;  // pinned
    int64_t* h_scalar = nullptr;    // pinned
    std::vector<cudaEvent_t> events;
    size_t ev_used = 0;
    // caching pool: free blocks by capacity, live blocks by address
    std::multimap<size_t, void*> pool_free;
    std::unordered_map<void*, size_t> pool_live;
    size_t pool_bytes = 0;
    size_t dev_total_mem = 0;
};

struct spada_b200_csr {
    spada_b200* h;
    DevCsr d;
    bool owned;
    // fiber store (DevCsr::desc): built on first use as the B operand for owned matrices, or by
    // spada_b200_csr_prepare for wrapped ones; engine-owned pool blocks
    bool fib_ready = false;          // build attempted (desc stays NULL when the operand cannot have one)
    unsigned long long* desc = nullptr;
    int32_t* gcol = nullptr;         // NULL: descriptors address the canonical arrays (no padding)
    double* gval = nullptr;
    int64_t fib_extent = 0;          // elements of gcol / gval (>= nnz)
    float fib_ms = 0.f;
};

struct spada_b200_result {
    spada_b200* h;
    uint64_t rows, cols, nnz;
    int64_t* ptr;
    int32_t* col;
    double* val;
    spada_b200_stats stats;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Pool policy: capacities are rounded up (512 B below 1 MiB, 2 MiB above); a request takes the
// smallest cached block that fits and wastes at most 25 % (or 1 MiB); otherwise cudaMalloc, and
// on out-of-memory the cache is released once and the allocation retried.
size_t pool_round(size_t bytes) {
    if (bytes < (1u << 20)) return (bytes + 511) & ~(size_t)511;
    return (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
}
void pool_release_cached(spada_b200* h) {
    cudaStreamSynchronize(h->stream);
    for (auto& kv : h->pool_free) {
        cudaFree(kv.second);
        h->pool_bytes -= kv.first;
    }
    h->pool_free.clear();
}
int pool_alloc(spada_b200* h, void** p, size_t bytes) {
    *p = nullptr;
    size_t cap = pool_round(bytes ? bytes : 1);
    auto it = h->pool_free.lower_bound(cap);
    if (it != h->pool_free.end() && it->first <= cap + std::max<size_t>(cap / 4, 1u << 20)) {
        *p = it->second;
        h->pool_live[*p] = it->first;
        h->pool_free.erase(it);
        return 0;
    }
    cudaError_t e = cudaMalloc(p, cap);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        pool_release_cached(h);
        e = cudaMalloc(p, cap);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? SPADA_B200_OOM : SPADA_B200_CUDA_ERROR,
                    "cudaMalloc(%zu bytes) failed: %s (device: %zu MiB free of %zu, this handle holds %zu MiB)", cap,
                    cudaGetErrorString(e), free_b >> 20, total_b >> 20, h->pool_bytes >> 20);
    }
    h->pool_live[*p] = cap;
    h->pool_bytes += cap;
    return 0;
}
void pool_free(spada_b200* h, void* p) {
    if (!p) return;
    auto it = h->pool_live.find(p);
    if (it == h->pool_live.end()) return;
    h->pool_free.emplace(it->second, p);
    h->pool_live.erase(it);
}
template <typename T>
int dalloc(spada_b200* h, T** p, size_t count) {
    return pool_alloc(h, (void**)p, count * sizeof(T));
}
template <typename T>
void dfree(spada_b200* h, T* p) {
    pool_free(h, (void*)p);
}

cudaEvent_t next_event(spada_b200* h, cudaStream_t on = nullptr) {
    if (h->ev_used == h->events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->events.push_back(e);
    }
    cudaEvent_t e = h->events[h->ev_used++];
    cudaEventRecord(e, on ? on : h->stream);
    return e;
}

int check_csr_args(uint64_t rows, uint64_t cols, uint64_t nnz, const void* indptr, const void* indices,
                   const void* data) {
    if (!indptr) return fail(SPADA_B200_INVALID_ARG, "indptr is NULL");
    if (nnz && (!indices || !data)) return fail(SPADA_B200_INVALID_ARG, "indices/data is NULL with nnz > 0");
    if (cols >= (1ull << 31)) return fail(SPADA_B200_TOO_LARGE, "cols = %llu >= 2^31", (unsigned long long)cols);
    if (rows >= (1ull << 32) - 2) return fail(SPADA_B200_TOO_LARGE, "rows = %llu >= 2^32-2", (unsigned long long)rows);
    return 0;
}

int validate_device_csr(spada_b200* h, const DevCsr& d) {
    if (d.rows == 0) {
        if (d.nnz != 0) return fail(SPADA_B200_UNSORTED_INPUT, "matrix with 0 rows has nnz != 0");
        return 0;
    }
    CU(cudaMemsetAsync(h->d_ctr, 0, sizeof(PlanCounters), h->stream));
    launch_validate(d, h->d_ctr, h->stream);
    CU(cudaGetLastError());
    launch_publish(h->d_ctr, h->h_ctr, sizeof(PlanCounters), h->stream);
    CU(cudaStreamSynchronize(h->stream));
    if (h->h_ctr->invalid_rows != 0 || h->h_ctr->long_rows != h->h_ctr->scan_ticket)
        return fail(SPADA_B200_UNSORTED_INPUT,
                    "input is not canonical CSR (%u out-of-range entries / bad row pointers, %u unsorted or "
                    "duplicate column ids inside rows)",
                    h->h_ctr->invalid_rows, h->h_ctr->long_rows - h->h_ctr->scan_ticket);
    return 0;
}

int make_csr(spada_b200* h, uint64_t rows, uint64_t cols, uint64_t nnz, spada_b200_csr** out, int64_t** ptr,
             int32_t** col, double** val) {
    int rc;
    *ptr = nullptr;
    *col = nullptr;
    *val = nullptr;
    spada_b200_csr* m = nullptr;
    if ((rc = dalloc(h, ptr, rows + 1)) || (rc = dalloc(h, col, nnz)) || (rc = dalloc(h, val, nnz)) ||
        !(m = new (std::nothrow) spada_b200_csr)) {
        dfree(h, *ptr);
        dfree(h, *col);
        dfree(h, *val);
        return rc ? rc : fail(SPADA_B200_OOM, "host allocation failed");
    }
    m->h = h;
    m->d = DevCsr{*ptr, *col, *val, (int64_t)rows, (int64_t)cols, (int64_t)nnz};
    m->owned = true;
    *out = m;
    return 0;
}

}  // namespace

// ---- library ------------------------------------------------------------------------------
extern "C" int spada_b200_abi_version(void) { return SPADA_B200_ABI_VERSION; }
extern "C" const char* spada_b200_last_error(void) { return g_err; }

extern "C" int spada_b200_device_count(int* count) {
    if (!count) return fail(SPADA_B200_INVALID_ARG, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(SPADA_B200_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return 0;
}

extern "C" int spada_b200_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(SPADA_B200_INVALID_ARG, "ptr is NULL");
    CU(cudaMallocHost(ptr, bytes ? bytes : 1));
    return 0;
}
extern "C" int spada_b200_host_free(void* ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return 0;
}

// ---- handle -------------------------------------------------------------------------------
static void destroy_handle(spada_b200* h) {
    if (!h) return;
    DeviceGuard g(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->side) cudaStreamSynchronize(h->side);
    for (auto& kv : h->pool_free) cudaFree(kv.second);
    h->pool_free.clear();
    for (auto& kv : h->pool_live) cudaFree(kv.first);  // objects the caller never freed
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->d_ctr) cudaFree(h->d_ctr);
    if (h->d_ctr_side) cudaFree(h->d_ctr_side);
    if (h->h_ctr) cudaFreeHost(h->h_ctr);
    if (h->h_scalar) cudaFreeHost(h->h_scalar);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

extern "C" int spada_b200_create(const spada_b200_opts* opts, spada_b200_t** out) {
    if (!out) return fail(SPADA_B200_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SPADA_B200_NO_DEVICE, "no CUDA device (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    spada_b200* h = new (std::nothrow) spada_b200;
    if (!h) return fail(SPADA_B200_OOM, "host allocation failed");
    // every early return below releases what has been acquired so far
    struct Guard {
        spada_b200* h;
        ~Guard() { destroy_handle(h); }
    } guard{h};
    if (opts) h->opts = *opts;
    else {
        h->opts.device = -1;
        h->opts.accelerator = SPADA_B200_ACC_SPADA;
        h->opts.flags = SPADA_B200_FLAG_VALIDATE;
    }
    if (h->opts.lane_num == 0) h->opts.lane_num = 8;
    // The accelerator argument of the reference CLI selects the window policy (main.rs:67-72,
    // scheduler.rs:729-753); it never changes C.  Ip = row-wise [1, L]: every row its own group ->
    // separate passes; Op = column-wise [L, 1] and MultiRow with R > 1: rows share a tile -> single
    // pass; Spada = adaptive (the engine decides per operand).  Explicit flags win.
    if (!(h->opts.flags & (SPADA_B200_FLAG_TWO_PHASE | SPADA_B200_FLAG_SINGLE_PASS))) {
        if (h->opts.accelerator == SPADA_B200_ACC_IP) h->opts.flags |= SPADA_B200_FLAG_TWO_PHASE;
        else if (h->opts.accelerator == SPADA_B200_ACC_OP) h->opts.flags |= SPADA_B200_FLAG_SINGLE_PASS;
        else if (h->opts.accelerator == SPADA_B200_ACC_MULTIROW)
            h->opts.flags |= h->opts.block_shape[0] > 1 ? SPADA_B200_FLAG_SINGLE_PASS : SPADA_B200_FLAG_TWO_PHASE;
    }
    if (h->opts.device < 0) {
        CU(cudaGetDevice(&h->device));
    } else {
        if (h->opts.device >= n) return fail(SPADA_B200_INVALID_ARG, "device %d out of range (count %d)", h->opts.device, n);
        h->device = h->opts.device;
    }
    DeviceGuard g(h->device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10)
        return fail(SPADA_B200_NO_DEVICE, "device %d is sm_%d%d; this build targets sm_100a only", h->device, prop.major,
                    prop.minor);
    h->sm_count = prop.multiProcessorCount;
    h->dev_total_mem = prop.totalGlobalMem;
    if (h->opts.stream) {
        h->stream = (cudaStream_t)h->opts.stream;
    } else {
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    {
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        CU(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi_prio));
        CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        if (const char* e = getenv("SPADA_B200_FIBER_PAD")) h->fiber_pad = atoi(e);
        if (const char* e = getenv("SPADA_B200_TILE_PASS")) h->tile_pass = atoi(e) != 0;
        if (const char* e = getenv("SPADA_B200_LONG_WS_MB")) {
            long mb = atol(e);
            if (mb > 0) h->long_ws_budget = (size_t)mb << 20;
        }
    }
    CU(cudaMalloc((void**)&h->d_ctr, sizeof(PlanCounters)));
    CU(cudaMalloc((void**)&h->d_ctr_side, sizeof(PlanCounters)));
    CU(cudaMemset(h->d_ctr_side, 0, sizeof(PlanCounters)));
    CU(cudaHostAlloc((void**)&h->h_ctr, sizeof(PlanCounters), cudaHostAllocMapped));   // written by launch_publish
    CU(cudaHostAlloc((void**)&h->h_scalar, 64, cudaHostAllocMapped));
    guard.h = nullptr;
    *out = h;
    return 0;
}

extern "C" void spada_b200_destroy(spada_b200_t* h) { destroy_handle(h); }

extern "C" int spada_b200_set_stream(spada_b200_t* h, void* cuda_stream) {
    if (!h) return fail(SPADA_B200_INVALID_ARG, "handle is NULL");
    DeviceGuard g(h->device);
    CU(cudaStreamSynchronize(h->stream));
    if (h->own_stream) {
        cudaStreamDestroy(h->stream);
        h->own_stream = false;
    }
    if (cuda_stream) {
        h->stream = (cudaStream_t)cuda_stream;
    } else {
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    return 0;
}

extern "C" int spada_b200_synchronize(spada_b200_t* h) {
    if (!h) return fail(SPADA_B200_INVALID_ARG, "handle is NULL");
    DeviceGuard g(h->device);
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int spada_b200_trim(spada_b200_t* h) {
    if (!h) return fail(SPADA_B200_INVALID_ARG, "handle is NULL");
    DeviceGuard g(h->device);
    pool_release_cached(h);
    return 0;
}

// ---- operands -----------------------------------------------------------------------------
extern "C" int spada_b200_upload(spada_b200_t* h, const spada_csr_view* m, spada_b200_csr_t** out) {
    if (!h || !m || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    int rc;
    if ((rc = check_csr_args(m->rows, m->cols, m->nnz, m->indptr, m->indices, m->data))) return rc;
    if (m->indptr[m->rows] != m->nnz)
        return fail(SPADA_B200_INVALID_ARG, "indptr[rows] = %llu != nnz = %llu", (unsigned long long)m->indptr[m->rows],
                    (unsigned long long)m->nnz);
    DeviceGuard g(h->device);
    int64_t* ptr;
    int32_t* col;
    double* val;
    spada_b200_csr* c;
    if ((rc = make_csr(h, m->rows, m->cols, m->nnz, &c, &ptr, &col, &val))) return rc;
    // every failure below releases the operand (and the temporary) again
    uint64_t* tmp = nullptr;
    auto body = [&]() -> int {
        // usize row pointers are bit-identical to i64 below 2^63; column ids are narrowed on the device
        CU(cudaMemcpyAsync(ptr, m->indptr, (m->rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
        if (m->nnz) {
            int rc2;
            if ((rc2 = dalloc(h, &tmp, m->nnz))) return rc2;
            CU(cudaMemcpyAsync(tmp, m->indices, m->nnz * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
            launch_widen_u64(nullptr, 0, nullptr, tmp, (int64_t)m->nnz, col, h->stream);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(val, m->data, m->nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        if (h->opts.flags & SPADA_B200_FLAG_VALIDATE) {
            // a u64 column id >= 2^31 would alias after narrowing: check on the host side of the copy
            for (uint64_t i = 0; i < m->nnz; ++i)
                if (m->indices[i] >= m->cols)
                    return fail(SPADA_B200_UNSORTED_INPUT, "column id %llu out of range at position %llu",
                                (unsigned long long)m->indices[i], (unsigned long long)i);
            return validate_device_csr(h, c->d);
        }
        return 0;
    };
    rc = body();
    dfree(h, tmp);
    if (rc) {
        spada_b200_csr_free(c);
        return rc;
    }
    *out = c;
    return 0;
}

extern "C" int spada_b200_upload32(spada_b200_t* h, const spada_csr_view32* m, spada_b200_csr_t** out) {
    if (!h || !m || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    int rc;
    if ((rc = check_csr_args(m->rows, m->cols, m->nnz, m->indptr, m->indices, m->data))) return rc;
    if (m->nnz >= (1ull << 31)) return fail(SPADA_B200_TOO_LARGE, "nnz >= 2^31 needs the 64-bit view");
    if ((uint64_t)(int64_t)m->indptr[m->rows] != m->nnz)
        return fail(SPADA_B200_INVALID_ARG, "indptr[rows] = %d != nnz = %llu", m->indptr[m->rows],
                    (unsigned long long)m->nnz);
    DeviceGuard g(h->device);
    int64_t* ptr;
    int32_t* col;
    double* val;
    spada_b200_csr* c;
    if ((rc = make_csr(h, m->rows, m->cols, m->nnz, &c, &ptr, &col, &val))) return rc;
    int32_t* tmp = nullptr;
    auto body = [&]() -> int {
        int rc2;
        if ((rc2 = dalloc(h, &tmp, m->rows + 1))) return rc2;
        CU(cudaMemcpyAsync(tmp, m->indptr, (m->rows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        launch_widen_i32(tmp, (int64_t)m->rows + 1, ptr, h->stream);
        CU(cudaGetLastError());
        if (m->nnz) {
            CU(cudaMemcpyAsync(col, m->indices, m->nnz * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemcpyAsync(val, m->data, m->nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        if (h->opts.flags & SPADA_B200_FLAG_VALIDATE) return validate_device_csr(h, c->d);
        return 0;
    };
    rc = body();
    dfree(h, tmp);
    if (rc) {
        spada_b200_csr_free(c);
        return rc;
    }
    *out = c;
    return 0;
}

extern "C" int spada_b200_csr_wrap_device(spada_b200_t* h, uint64_t rows, uint64_t cols, uint64_t nnz,
                                          const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
                                          spada_b200_csr_t** out) {
    if (!h || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    int rc;
    if ((rc = check_csr_args(rows, cols, nnz, d_indptr, d_indices, d_data))) return rc;
    spada_b200_csr* c = new (std::nothrow) spada_b200_csr;
    if (!c) return fail(SPADA_B200_OOM, "host allocation failed");
    c->h = h;
    c->d = DevCsr{d_indptr, d_indices, d_data, (int64_t)rows, (int64_t)cols, (int64_t)nnz};
    c->owned = false;
    *out = c;
    return 0;
}

extern "C" int spada_b200_csr_shape(const spada_b200_csr_t* m, uint64_t* rows, uint64_t* cols, uint64_t* nnz) {
    if (!m) return fail(SPADA_B200_INVALID_ARG, "matrix is NULL");
    if (rows) *rows = (uint64_t)m->d.rows;
    if (cols) *cols = (uint64_t)m->d.cols;
    if (nnz) *nnz = (uint64_t)m->d.nnz;
    return 0;
}

extern "C" int spada_b200_csr_device_ptrs(const spada_b200_csr_t* m, const int64_t** d_indptr,
                                          const int32_t** d_indices, const double** d_data) {
    if (!m) return fail(SPADA_B200_INVALID_ARG, "matrix is NULL");
    if (d_indptr) *d_indptr = m->d.ptr;
    if (d_indices) *d_indices = m->d.col;
    if (d_data) *d_data = m->d.val;
    return 0;
}

extern "C" int spada_b200_csr_download32(const spada_b200_csr_t* m, int64_t* indptr, int32_t* indices, double* data) {
    if (!m || !indptr) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    if (m->d.nnz && (!indices || !data)) return fail(SPADA_B200_INVALID_ARG, "NULL output array");
    spada_b200* h = m->h;
    DeviceGuard g(h->device);
    CU(cudaMemcpyAsync(indptr, m->d.ptr, (size_t)(m->d.rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (m->d.nnz) {
        CU(cudaMemcpyAsync(indices, m->d.col, (size_t)m->d.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(data, m->d.val, (size_t)m->d.nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" void spada_b200_csr_free(spada_b200_csr_t* m) {
    if (!m) return;
    if (m->desc || m->gcol || m->gval) {
        DeviceGuard g(m->h->device);
        dfree(m->h, m->desc);
        dfree(m->h, m->gcol);
        dfree(m->h, m->gval);
    }
    if (m->owned) {
        DeviceGuard g(m->h->device);
        dfree(m->h, const_cast<int64_t*>(m->d.ptr));
        dfree(m->h, const_cast<int32_t*>(m->d.col));
        dfree(m->h, const_cast<double*>(m->d.val));
    }
    delete m;
}

// ---- fiber store of a B operand ---------------------------------------------------------------
namespace {
int build_fibers(spada_b200* h, spada_b200_csr* c) {
    if (c->fib_ready) return 0;
    c->fib_ready = true;
    const DevCsr& d = c->d;
    int pad = h->fiber_pad;
    if (pad < 0) pad = (d.rows > 0 && d.nnz >= 6 * d.rows) ? FIBER_PAD : 1;
    if (pad == 0 || d.rows == 0 || d.nnz == 0) return 0;
    if (pad > 1) pad = FIBER_PAD;
    cudaStream_t s = h->stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    uint32_t* d_len = nullptr;
    int64_t* d_start = nullptr;
    uint64_t* d_tiles = nullptr;
    int rc = 0;
    auto done = [&](int code) {
        dfree(h, d_len);
        dfree(h, d_start);
        dfree(h, d_tiles);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return code;
    };
    if ((rc = dalloc(h, &d_len, (size_t)d.rows))) return done(rc);
    CU(cudaMemsetAsync(h->d_ctr, 0, sizeof(PlanCounters), s));
    launch_fiber_lengths(d.ptr, d.rows, (uint32_t)pad, d_len, h->d_ctr, s);
    int64_t total = d.nnz;
    if (pad > 1) {
        if ((rc = dalloc(h, &d_start, (size_t)d.rows + 1))) return done(rc);
        if ((rc = dalloc(h, &d_tiles, scan_tile_state_words(d.rows)))) return done(rc);
        launch_scan_u32_i64(d_len, d.rows, d_start, d_tiles, h->d_ctr, s);
        if (cudaMemcpyAsync(h->h_scalar, d_start + d.rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s) != cudaSuccess)
            return done(fail(SPADA_B200_CUDA_ERROR, "fiber store: copy failed"));
    }
    if (cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(PlanCounters), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return done(fail(SPADA_B200_CUDA_ERROR, "fiber store: %s", cudaGetErrorString(cudaGetLastError())));
    if (h->h_ctr->invalid_rows) return done(0);   // a row of 2^24 or more elements: the kernels use row_ptr
    if (pad > 1) total = h->h_scalar[0];
    if ((rc = dalloc(h, &c->desc, (size_t)d.rows))) return done(rc);
    if (pad > 1) {
        if ((rc = dalloc(h, &c->gcol, (size_t)total)) || (rc = dalloc(h, &c->gval, (size_t)total))) {
            dfree(h, c->desc);
            dfree(h, c->gcol);
            c->desc = nullptr;
            c->gcol = nullptr;
            return done(rc == SPADA_B200_OOM ? 0 : rc);   // no room for the copy: run without it
        }
    }
    c->fib_extent = total;
    launch_fiber_fill(d, d_start, c->desc, c->gcol, c->gval, s);
    cudaEventRecord(e1, s);
    if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return done(fail(SPADA_B200_CUDA_ERROR, "fiber store: fill failed"));
    cudaEventElapsedTime(&c->fib_ms, e0, e1);
    return done(0);
}
}  // namespace

extern "C" int spada_b200_csr_prepare(spada_b200_t* h, spada_b200_csr_t* m, float* ms_or_null) {
    if (!h || !m) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    DeviceGuard g(h->device);
    int rc = build_fibers(h, m);
    if (ms_or_null) *ms_or_null = m->fib_ms;
    return rc;
}

extern "C" int spada_b200_csr_set_one_shot(spada_b200_csr_t* m) {
    if (!m) return fail(SPADA_B200_INVALID_ARG, "matrix is NULL");
    m->fib_ready = true;   // build_fibers() will not run for it: the kernels gather through row_ptr
    return 0;
}

// ---- B = A^T on the device: replaces GEMM::from_mat's transpose_into().to_csr() (gemm.rs:44-46) --------------
extern "C" int spada_b200_transpose(spada_b200_t* h, const spada_b200_csr_t* a, spada_b200_csr_t** out) {
    if (!h || !a || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    const DevCsr& A = a->d;
    if (A.nnz >= (1ll << 32)) return fail(SPADA_B200_TOO_LARGE, "transpose: nnz >= 2^32");
    if (A.rows >= (1ll << 31)) return fail(SPADA_B200_TOO_LARGE, "transpose: rows >= 2^31 (they become column ids)");
    DeviceGuard g(h->device);
    cudaStream_t s = h->stream;
    int rc;
    int64_t* ptr;
    int32_t* col;
    double* val;
    spada_b200_csr* c;
    if ((rc = make_csr(h, (uint64_t)A.cols, (uint64_t)A.rows, (uint64_t)A.nnz, &c, &ptr, &col, &val))) return rc;
    const int64_t n = A.nnz, k = A.cols;
    const int passes = transpose_passes(k);
    const int64_t tiles = transpose_tiles(n);
    uint32_t *d_erow = nullptr, *d_cnt = nullptr, *d_hist = nullptr, *d_pay[2] = {nullptr, nullptr};
    int32_t* d_key[2] = {nullptr, nullptr};
    int64_t* d_offs = nullptr;
    uint64_t* d_tiles = nullptr;
    auto done = [&](int code) {
        dfree(h, d_erow);
        dfree(h, d_cnt);
        dfree(h, d_hist);
        dfree(h, d_pay[0]);
        dfree(h, d_pay[1]);
        dfree(h, d_key[0]);
        dfree(h, d_key[1]);
        dfree(h, d_offs);
        dfree(h, d_tiles);
        if (code) spada_b200_csr_free(c);
        return code;
    };
    if (n == 0 || k == 0) {
        if (cudaMemsetAsync(ptr, 0, (size_t)(k + 1) * sizeof(int64_t), s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess)
            return done(fail(SPADA_B200_CUDA_ERROR, "transpose: %s", cudaGetErrorString(cudaGetLastError())));
        *out = c;
        return done(0);
    }
    const size_t scan_n = (size_t)std::max<int64_t>(k, (int64_t)256 * tiles);
    if ((rc = dalloc(h, &d_erow, (size_t)n)) || (rc = dalloc(h, &d_cnt, (size_t)k)) ||
        (rc = dalloc(h, &d_tiles, scan_tile_state_words((int64_t)scan_n))))
        return done(rc);
    if (passes > 0) {
        if ((rc = dalloc(h, &d_hist, (size_t)256 * tiles)) || (rc = dalloc(h, &d_offs, (size_t)256 * tiles + 1)) ||
            (rc = dalloc(h, &d_key[0], (size_t)n)) || (rc = dalloc(h, &d_pay[0], (size_t)n)))
            return done(rc);
        if (passes > 1 && ((rc = dalloc(h, &d_key[1], (size_t)n)) || (rc = dalloc(h, &d_pay[1], (size_t)n)))) return done(rc);
    }
    cudaMemsetAsync(d_cnt, 0, (size_t)k * sizeof(uint32_t), s);
    launch_entry_rows(A, d_erow, d_cnt, s);
    launch_scan_u32_i64(d_cnt, k, ptr, d_tiles, h->d_ctr, s);   // row_ptr of A^T
    const int32_t* key_in = A.col;
    const uint32_t* pay_in = nullptr;   // pass 0: payload = entry index
    for (int p = 0; p < passes; ++p) {
        launch_radix_hist(key_in, n, 8 * p, d_hist, s);
        launch_scan_u32_i64(d_hist, (int64_t)256 * tiles, d_offs, d_tiles, h->d_ctr, s);
        launch_radix_scatter(key_in, pay_in, n, 8 * p, d_offs, d_key[p & 1], d_pay[p & 1], s);
        key_in = d_key[p & 1];
        pay_in = d_pay[p & 1];
    }
    launch_transpose_gather(pay_in, d_erow, A.val, n, col, val, s);
    if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return done(fail(SPADA_B200_CUDA_ERROR, "transpose: %s", cudaGetErrorString(cudaGetLastError())));
    *out = c;
    return done(0);
}

// ---- stage 1 alone ------------------------------------------------------------------------
namespace {
int run_flops(spada_b200* h, const DevCsr& a, const DevCsr& b, int64_t row_begin, int64_t m, uint32_t* d_flops,
              uint32_t* d_long) {
    CU(cudaMemsetAsync(h->d_ctr, 0, sizeof(PlanCounters), h->stream));
    uint32_t* d_blen = nullptr;   // lengths of B's rows, 4 B each (freed right away: same-stream reuse is ordered)
    int rc = dalloc(h, &d_blen, (size_t)std::max<int64_t>(b.rows, 1));
    if (rc) return rc;
    launch_flops(a, b.ptr, b.rows, d_blen, row_begin, m, d_flops, d_long, h->d_ctr, h->stream);
    dfree(h, d_blen);
    CU(cudaGetLastError());
    launch_publish(h->d_ctr, h->h_ctr, sizeof(PlanCounters), h->stream);
    return 0;
}
}  // namespace

extern "C" int spada_b200_flops(spada_b200_t* h, const spada_b200_csr_t* a, const spada_b200_csr_t* b,
                                uint64_t* total_products, uint64_t* host_flops) {
    if (!h || !a || !b) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    if (a->d.cols != b->d.rows)
        return fail(SPADA_B200_DIM_MISMATCH, "A is %lld x %lld but B has %lld rows", (long long)a->d.rows,
                    (long long)a->d.cols, (long long)b->d.rows);
    DeviceGuard g(h->device);
    int64_t m = a->d.rows;
    uint32_t *d_flops = nullptr, *d_long = nullptr;
    std::vector<uint32_t> tmp;
    auto body = [&]() -> int {
        int rc;
        if ((rc = dalloc(h, &d_flops, (size_t)std::max<int64_t>(m, 1)))) return rc;
        if ((rc = dalloc(h, &d_long, (size_t)(a->d.nnz / 256 + 2)))) return rc;
        if ((rc = run_flops(h, a->d, b->d, 0, m, d_flops, d_long))) return rc;
        if (host_flops && m) {
            tmp.resize((size_t)m);
            CU(cudaMemcpyAsync(tmp.data(), d_flops, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        return 0;
    };
    const int rc = body();
    dfree(h, d_flops);
    dfree(h, d_long);
    if (rc) return rc;
    if (total_products) *total_products = h->h_ctr->total_products;
    if (host_flops)
        for (int64_t i = 0; i < m; ++i) host_flops[i] = tmp[(size_t)i];
    return 0;
}

extern "C" int spada_b200_plan_shards(spada_b200_t* h, const spada_b200_csr_t* a, const spada_b200_csr_t* b,
                                      uint32_t n_shards, uint64_t* bounds) {
    if (!h || !a || !b || !bounds || n_shards == 0) return fail(SPADA_B200_INVALID_ARG, "bad argument");
    uint64_t m = (uint64_t)a->d.rows;
    std::vector<uint64_t> f((size_t)m);
    uint64_t total = 0;
    int rc = spada_b200_flops(h, a, b, &total, f.data());
    if (rc) return rc;
    // contiguous ranges with (as near as rows allow) equal intermediate-product counts; rows that
    // produce nothing still cost their A entries, so weigh every row by flops + 1
    uint64_t weight_total = total + m;
    bounds[0] = 0;
    uint64_t acc = 0, row = 0;
    for (uint32_t s = 1; s < n_shards; ++s) {
        uint64_t target = (uint64_t)((__uint128_t)weight_total * s / n_shards);
        while (row < m && acc + f[(size_t)row] + 1 <= target) {
            acc += f[(size_t)row] + 1;
            ++row;
        }
        bounds[s] = row;
    }
    bounds[n_shards] = m;
    return 0;
}

// ---- window choice ----------------------------------------------------------------------------
// Spada's window [R, L/R] (scheduler.rs:729-753, rowwise_perf_adjust.rs:121-252): R rows share the L lanes of a
// window, every row gets L/R of them.  On the GPU the bin of a row (its intermediate-product count) fixes the lanes that
// cooperate on it; what is still chosen per operand -- and per accelerator argument, main.rs:67-72 -- is how many rows
// share one cooperative group:
//   bin 1 (<= 32 products)   [4, 8]  four rows per warp, 8 lanes x 4 keys each, or [1, 32] one row per warp.  The
//                            4-row window only holds rows with at most 8 A entries (one per lane of the group).
//       Spada (adaptive)     [4, 8] when at least half of the bin's rows fit it (counted by stage 1), else [1, 32]
//       Ip  (row-wise)       [1, L']: one row per group, L' = 32 lanes
//       Op  (column-wise)    [L'/8, 8]: rows share the lanes
//       MultiRow             [R, L'/R] with R = block_shape[0]: R >= 4 -> [4, 8], else [1, 32]; lane_num > 8 widens
//                            the per-row share to a full warp ([1, 32])
//   bins 2..5                one warp per row; rows per tile 4 (scratch pass) or 8 / 4 (single pass, look-back tile)
//   bins 6..8, long rows     one CTA (256 lanes) per row / per chunk of 4096 products ("K-tiling": chunks x [1, 256])
struct WindowChoice {
    bool tiny_quad;
};
static WindowChoice window_choice(const spada_b200* h, const PlanCounters& pc) {
    WindowChoice w{};
    switch (h->opts.accelerator) {
        case SPADA_B200_ACC_IP: w.tiny_quad = false; break;
        case SPADA_B200_ACC_OP: w.tiny_quad = h->opts.lane_num <= 8; break;
        case SPADA_B200_ACC_MULTIROW: w.tiny_quad = h->opts.block_shape[0] >= 4 && h->opts.lane_num <= 8; break;
        default: w.tiny_quad = (uint64_t)pc.tiny_fit * 2 >= pc.bin_rows[1]; break;
    }
    return w;
}
static void window_report(const WindowChoice& w, bool fused, bool tile_mode, const PlanCounters& pc, spada_b200_stats& st) {
    for (int bnum = 0; bnum < NUM_BINS; ++bnum) {
        uint32_t R = 0, lanes = 0;
        if (tile_mode && bnum >= 1 && bnum <= 5) {
            R = (uint32_t)(tile_pass_cut() / bin_capacity(bnum));   // rows of this size that share a work-cut tile
            lanes = 32;
        } else if (bnum == 1) {
            const bool quad = fused && w.tiny_quad;   // the scratch pass runs bin 1 one row per warp
            R = quad ? 4 : 1;
            lanes = quad ? 8 : 32;
        } else if (bnum >= 2 && bnum <= 5) {
            R = 1;
            lanes = 32;
        } else if (bnum >= 6) {
            R = 1;
            lanes = 256;
        }
        st.bin_window_rows[bnum] = pc.bin_rows[bnum] ? R : 0;
        st.bin_window_lanes[bnum] = pc.bin_rows[bnum] ? lanes : 0;
    }
}

// ---- the hot path -------------------------------------------------------------------------
// A product runs in two halves (one ABI call does both; the sharded entry points expose them separately so that the
// ranks of a multi-GPU run can exchange their nnz in between):
//   begin   stage 1 (flop count, bins), then every row that is not computed in place goes through ONE pass into a
//           scratch CSR laid out by product count: sort bins on the main stream, long rows (chunk sorts + merge
//           levels, longrow.cu) on a side stream; the pass records every row's nnz; row_ptr scan.
//   finish  C is sized, the scratch rows are copied to their final place -- into this GPU's C and, for a sharded
//           run, into every peer's C at the shard's global offset (the all-gather fused into the store).
// Single-pass mode (stencils, uniform graphs): rows <= 512 products are expanded, sorted, reduced and placed by a
// look-back scan in one kernel (fused.cu) straight into C; only heavier rows take the scratch route.
struct LaunchRec {
    char name[32];
    cudaEvent_t e0, e1;
    uint32_t grid;
    uint64_t rows, products, nnz;
    int stage;  // 1 flops, 2 first pass, 3 placement, 4 scan
};

static const char* bin_name(int b) {
    static char names[NUM_BINS][12];
    static bool init = false;
    if (!init) {
        snprintf(names[0], sizeof(names[0]), "empty");
        for (int i = 1; i < NUM_BINS; ++i) snprintf(names[i], sizeof(names[i]), "%llu", (unsigned long long)bin_capacity(i));
        init = true;
    }
    return names[b];
}

struct spada_b200_shard {
    spada_b200* h = nullptr;
    spada_b200_result* R = nullptr;
    DevCsr A{}, B{};
    int64_t row_begin = 0, m = 0;
    bool fused = false, tile_mode = false, scratch = false, forked = false, finished = false;
    PlanCounters pc{};
    BinTable tbl{};
    bool identity = false;
    WindowChoice window{};
    uint32_t scratch_lo = 0;      // rows with more than this many products go through the scratch CSR
    uint64_t scratch_products = 0;
    uint32_t n_long = 0;
    uint64_t long_products = 0;
    int max_light_bin = 0;
    uint64_t light_rows = 0;
    int64_t nnz_c = 0;
    // device blocks (pool)
    uint32_t *d_flops = nullptr, *d_long = nullptr, *d_perm = nullptr, *d_nnz = nullptr, *d_masked = nullptr,
             *d_blen = nullptr, *d_aseq = nullptr;
    uint64_t *d_tiles = nullptr, *d_tiles_side = nullptr;
    int64_t* d_prod_ptr = nullptr;
    int32_t* d_tcol = nullptr;
    double* d_tval = nullptr;
    // long-row wave workspace
    // tile pass workspace
    uint32_t *t_w = nullptr, *t_start = nullptr;
    int64_t *t_lp = nullptr, *t_idx = nullptr;
    uint64_t* t_state = nullptr;
    size_t t_bound = 0;
    // long rows: tables of the whole list, the pong buffer of one wave, merge-tile descriptors
    uint32_t *w_p = nullptr, *w_u = nullptr, *w_heads = nullptr, *w_unit_row = nullptr;
    uint64_t* w_tiles = nullptr;
    int64_t *w_prod_off = nullptr, *w_unit_off = nullptr, *w_hoff = nullptr;
    int32_t* w_col = nullptr;
    double* w_val = nullptr;
    bool dense_long = false;
    LongPlan long_plan{};
    std::vector<LaunchRec> recs;
    cudaStream_t rec_stream = nullptr;
    uint32_t kernels = 0;
    cudaEvent_t e_end = nullptr;

    const uint32_t* perm_of(int bin) const { return identity ? nullptr : d_perm + tbl.offset[bin]; }
    void begin_rec(const char* name, int stage, uint32_t grid, uint64_t rows, uint64_t products, cudaStream_t on = nullptr) {
        LaunchRec r{};
        snprintf(r.name, sizeof(r.name), "%s", name);
        r.stage = stage;
        r.grid = grid;
        r.rows = rows;
        r.products = products;
        rec_stream = on ? on : h->stream;
        r.e0 = next_event(h, rec_stream);
        recs.push_back(r);
    }
    void end_rec() { recs.back().e1 = next_event(h, rec_stream); }
    void release_work() {
        if (forked) cudaStreamSynchronize(h->side);
        forked = false;
        uint32_t** u32s[] = {&d_flops, &d_long, &d_perm, &d_nnz, &d_masked, &d_blen, &d_aseq, &w_p, &w_u, &w_heads, &w_unit_row,
                             &t_w, &t_start};
        for (auto pp : u32s) { dfree(h, *pp); *pp = nullptr; }
        uint64_t** u64s[] = {&d_tiles, &d_tiles_side, &w_tiles, &t_state};
        for (auto pp : u64s) { dfree(h, *pp); *pp = nullptr; }
        int64_t** i64s[] = {&d_prod_ptr, &w_prod_off, &w_unit_off, &w_hoff, &t_lp, &t_idx};
        for (auto pp : i64s) { dfree(h, *pp); *pp = nullptr; }
        dfree(h, d_tcol); d_tcol = nullptr;
        dfree(h, d_tval); d_tval = nullptr;
        dfree(h, w_col); w_col = nullptr;
        dfree(h, w_val); w_val = nullptr;
    }
};

namespace {

// waves of long rows: contiguous slices of the long-row list (ascending merge-level count) whose products fit the
// ping-pong buffers.  The host knows rows and products per level bin (stage-1 counters), and inside a bin only the
// bound 4096 * 2^L per row, which is what a bin that has to be cut is planned with.
struct WavePlan {
    uint32_t lo, hi;                 // slice of the long list
    uint64_t products_bound;
    uint64_t unit_bound;
    uint32_t level_lo[24];
    uint32_t level_grid[24];
    uint64_t level_products[24];
    int max_level;
};
std::vector<WavePlan> plan_waves(const PlanCounters& pc, uint64_t budget_products) {
    struct Piece { int level; uint32_t lo, hi; uint64_t pbound, ubound; };
    std::vector<std::vector<Piece>> waves;
    std::vector<Piece> cur;
    uint64_t cur_p = 0;
    uint32_t idx = 0;
    auto flush = [&]() {
        if (!cur.empty()) waves.push_back(cur);
        cur.clear();
        cur_p = 0;
    };
    for (int bnum = BIN_LONG0; bnum < NUM_BINS; ++bnum) {
        const uint32_t rows = pc.bin_rows[bnum];
        if (!rows) continue;
        const int L = bnum - 8;
        const uint64_t cap = (uint64_t)LONG_UNIT << L;          // most products a row of this bin can have
        const uint64_t prods = pc.bin_products[bnum];
        if (cur_p + prods <= budget_products) {                   // the whole bin joins the open wave (exact count)
            cur.push_back({L, idx, idx + rows, prods, prods / LONG_UNIT + rows});
            cur_p += prods;
        } else if (prods <= budget_products) {                    // the whole bin opens a new wave
            flush();
            cur.push_back({L, idx, idx + rows, prods, prods / LONG_UNIT + rows});
            cur_p = prods;
        } else {                                                  // the bin is cut by its per-row bound
            flush();
            const uint64_t per = std::max<uint64_t>(1, budget_products / cap);
            for (uint32_t lo = 0; lo < rows; lo += (uint32_t)std::min<uint64_t>(per, rows)) {
                const uint32_t hi = (uint32_t)std::min<uint64_t>(rows, (uint64_t)lo + per);
                const uint64_t pb = std::min<uint64_t>(prods, (uint64_t)(hi - lo) * cap);
                cur.push_back({L, idx + lo, idx + hi, pb, pb / LONG_UNIT + (hi - lo)});
                cur_p = pb;
                if (hi < rows) flush();
            }
        }
        idx += rows;
    }
    flush();
    std::vector<WavePlan> out;
    for (auto& w : waves) {
        WavePlan P{};
        P.lo = w.front().lo;
        P.hi = w.back().hi;
        P.max_level = w.back().level;
        for (auto& pc_ : w) {
            P.products_bound += pc_.pbound;
            P.unit_bound += pc_.ubound;
        }
        for (int l = 1; l <= P.max_level && l < 24; ++l) {
            uint32_t first = P.hi;
            uint64_t grid = 0, prods = 0;
            for (auto& pc_ : w)
                if (pc_.level >= l) {
                    first = std::min(first, pc_.lo);
                    grid += pc_.ubound;
                    prods += pc_.pbound;
                }
            P.level_lo[l] = first;
            P.level_grid[l] = (uint32_t)grid;
            P.level_products[l] = prods;
        }
        out.push_back(P);
    }
    return out;
}

void shard_free(spada_b200_shard* S) {
    if (!S) return;
    DeviceGuard g(S->h->device);
    S->release_work();
    if (S->R) spada_b200_result_free(S->R);
    delete S;
}

#define TRY(x)                \
    do {                      \
        if ((rc = (x))) {     \
            shard_free(S);    \
            return rc;        \
        }                     \
    } while (0)
#define CUT(expr)                                                                                 \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            shard_free(S);                                                                        \
            return fail(e__ == cudaErrorMemoryAllocation ? SPADA_B200_OOM : SPADA_B200_CUDA_ERROR, \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                         \
    } while (0)

// first half: everything up to the local row_ptr.  force_scratch: sharded runs place every row in the second half.
int shard_begin(spada_b200_t* h, const spada_b200_csr_t* a, const spada_b200_csr_t* b, uint64_t row_begin,
                uint64_t row_end, bool force_scratch, bool host_nnz, spada_b200_shard** out) {
    if (!h || !a || !b || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    if (a->d.cols != b->d.rows)
        return fail(SPADA_B200_DIM_MISMATCH, "A is %lld x %lld but B has %lld rows", (long long)a->d.rows,
                    (long long)a->d.cols, (long long)b->d.rows);
    if (row_end == UINT64_MAX) row_end = (uint64_t)a->d.rows;
    if (row_begin > row_end || row_end > (uint64_t)a->d.rows)
        return fail(SPADA_B200_INVALID_ARG, "row range [%llu, %llu) outside A.rows = %lld", (unsigned long long)row_begin,
                    (unsigned long long)row_end, (long long)a->d.rows);
    DeviceGuard g(h->device);
    cudaStream_t s = h->stream;
    int rc;
    // owned operands get their fiber store the first time they are used as B (wrapped ones: spada_b200_csr_prepare)
    if (b->owned && !b->fib_ready && (rc = build_fibers(h, const_cast<spada_b200_csr*>(b)))) return rc;
    spada_b200_shard* S = new (std::nothrow) spada_b200_shard;
    if (!S) return fail(SPADA_B200_OOM, "host allocation failed");
    S->h = h;
    S->A = a->d;
    S->B = b->d;
    if (b->desc) {
        S->B.desc = b->desc;
        if (b->gcol) {
            S->B.col = b->gcol;
            S->B.val = b->gval;
            S->B.nnz = b->fib_extent;
        }
    }
    const DevCsr& A = S->A;
    const DevCsr& B = S->B;
    const int64_t m = (int64_t)(row_end - row_begin);
    S->m = m;
    S->row_begin = (int64_t)row_begin;

    spada_b200_result* R = new (std::nothrow) spada_b200_result;
    if (!R) {
        delete S;
        return fail(SPADA_B200_OOM, "host allocation failed");
    }
    memset(R, 0, sizeof(*R));
    S->R = R;
    R->h = h;
    R->rows = (uint64_t)m;
    R->cols = (uint64_t)B.cols;
    spada_b200_stats& st = R->stats;
    st.rows = (uint64_t)m;
    st.cols = (uint64_t)B.cols;
    st.nnz_b = (uint64_t)b->d.nnz;
    if (row_begin == 0 && row_end == (uint64_t)A.rows) st.nnz_a = (uint64_t)A.nnz;
    TRY(dalloc(h, &R->ptr, (size_t)m + 1));
    h->ev_used = 0;
    if (m == 0) {
        CUT(cudaMemsetAsync(R->ptr, 0, sizeof(int64_t), s));
        *out = S;
        return 0;
    }
    PlanCounters& pc = S->pc;

    // ---- stage 1: flop count, bins (the window choice) --------------------------------------
    TRY(dalloc(h, &S->d_flops, (size_t)m));
    TRY(dalloc(h, &S->d_long, (size_t)(A.nnz / 256 + 2)));
    TRY(dalloc(h, &S->d_perm, (size_t)m));
    TRY(dalloc(h, &S->d_nnz, (size_t)m));
    TRY(dalloc(h, &S->d_blen, (size_t)std::max<int64_t>(B.rows, 1)));
    TRY(dalloc(h, &S->d_tiles, std::max(scan_tile_state_words(m), fused_tile_state_words(m))));
    S->begin_rec("flop_count", 1, (uint32_t)((m + 255) / 256), (uint64_t)m, 0);
    CUT(cudaMemsetAsync(h->d_ctr, 0, sizeof(PlanCounters), s));
    launch_flops(A, B.ptr, B.rows, S->d_blen, (int64_t)row_begin, m, S->d_flops, S->d_long, h->d_ctr, s);
    CUT(cudaGetLastError());
    launch_publish(h->d_ctr, h->h_ctr, sizeof(PlanCounters), s);
    S->kernels += 3;
    S->end_rec();
    CUT(cudaMemsetAsync(S->d_nnz, 0, (size_t)m * sizeof(uint32_t), s));
    CUT(cudaStreamSynchronize(s));  // host read-back #1: bin sizes
    pc = *h->h_ctr;
    st.products = pc.total_products;
    S->recs[0].products = pc.total_products;
    if (pc.max_flops == 0xffffffffu) {
        shard_free(S);
        return fail(SPADA_B200_TOO_LARGE, "a row of C has 2^32 or more intermediate products");
    }
    BinTable& tbl = S->tbl;
    uint32_t off = 0;
    uint64_t non_empty = 0, dominant = 0;
    for (int bnum = 0; bnum < NUM_BINS; ++bnum) {
        tbl.offset[bnum] = off;
        if (bnum != BIN_EMPTY) off += pc.bin_rows[bnum];
        st.bin_rows[bnum] = pc.bin_rows[bnum];
        st.bin_products[bnum] = pc.bin_products[bnum];
        if (bnum >= 1) non_empty += pc.bin_rows[bnum];
        if (bnum >= 1 && bnum <= 5) {
            dominant = std::max<uint64_t>(dominant, pc.bin_rows[bnum]);
            if (pc.bin_rows[bnum]) {
                S->max_light_bin = bnum;
                S->light_rows += pc.bin_rows[bnum];
            }
        }
        if (bnum >= BIN_LONG0) {
            S->n_long += pc.bin_rows[bnum];
            S->long_products += pc.bin_products[bnum];
        }
        // one sort bin holds every row: no permutation (the long-row path always works from a list)
        if (pc.bin_rows[bnum] == (uint64_t)m && bnum >= 1 && bnum < BIN_LONG0) S->identity = true;
    }
    tbl.offset[NUM_BINS] = off;
    if (!S->identity && off > 0) {
        S->begin_rec("bin_scatter", 1, (uint32_t)((m + 255) / 256), (uint64_t)m, 0);
        launch_bin_scatter(S->d_flops, m, tbl, S->d_perm, h->d_ctr, s);
        CUT(cudaGetLastError());
        S->kernels += 1;
        S->end_rec();
    }

    // Engine mode.  The look-back of the single pass places rows in row order, so a tile finishes with its slowest
    // row: it pays off when the rows look alike (one bin holds >= 80 % of the non-empty rows, e.g. stencils and
    // uniform random graphs: measured 1.3x-1.6x), not on heavy-tailed row lengths (0.9x) -- DESIGN.md section 4.
    bool fused = !force_scratch && !(h->opts.flags & SPADA_B200_FLAG_TWO_PHASE) && S->light_rows > 0 &&
                 (double)pc.total_products * 12.0 <= 0.45 * (double)h->dev_total_mem;
    // rows that look alike: fixed tiles of 4..32 rows (k_fused_light / k_fused_tiny4); mixed lengths: the scratch pass,
    // or with SPADA_B200_TILE_PASS=1 one pass over tiles cut by work (k_tile_pass)
    const bool uniform = dominant * 10 >= non_empty * 8;
    S->tile_mode = fused && !uniform && h->tile_pass;
    if (fused && !uniform && !h->tile_pass && !(h->opts.flags & SPADA_B200_FLAG_SINGLE_PASS)) fused = false;
    S->fused = fused;
    S->window = window_choice(h, pc);
    window_report(S->window, fused, S->tile_mode, pc, st);
    S->scratch_lo = fused ? 512u : 0u;
    for (int bnum = fused ? 6 : 1; bnum < NUM_BINS; ++bnum) S->scratch_products += pc.bin_products[bnum];
    S->scratch = S->scratch_products > 0;

    // long rows: wave plan + workspace
    std::vector<WavePlan> waves;
    uint64_t wave_products = 0, wave_units = 0;
    // few output columns: the long rows keep a dense accumulator in shared memory instead (no merge levels)
    const bool dense_long = S->n_long && dense_rows_fit(B.cols) && !getenv("SPADA_B200_NO_DENSE");
    S->dense_long = dense_long;
    const uint64_t all_units = S->long_products / LONG_UNIT + S->n_long;   // upper bound of the chunks of all long rows
    if (S->n_long && !dense_long) {
        waves = plan_waves(pc, std::max<uint64_t>(h->long_ws_budget / 12, (uint64_t)LONG_UNIT));
        for (auto& w : waves) {
            wave_products = std::max(wave_products, w.products_bound);
            wave_units = std::max(wave_units, w.unit_bound);
        }
    }
    {
        const double need = (double)S->scratch_products * 12.0 + (double)wave_products * 12.0 +
                            (fused ? (double)pc.total_products * 12.0 : 0.0);
        if (need > 0.92 * (double)h->dev_total_mem) {
            shard_free(S);
            return fail(SPADA_B200_OOM,
                        "the first pass needs %.1f GB of scratch rows for %llu intermediate products (device: %.1f GB); "
                        "compute C in row panels with spada_b200_spgemm_stream",
                        need / 1e9, (unsigned long long)pc.total_products, (double)h->dev_total_mem / 1e9);
        }
    }
    if (S->tile_mode) {
        uint64_t slots = 0;
        for (int bnum = 1; bnum <= 5; ++bnum) slots += (uint64_t)pc.bin_rows[bnum] * bin_capacity(bnum);
        S->t_bound = tile_pass_bound(m, slots);
        TRY(dalloc(h, &S->t_w, (size_t)m));
        TRY(dalloc(h, &S->t_lp, (size_t)m + 1));
        TRY(dalloc(h, &S->t_idx, (size_t)m + 1));
        TRY(dalloc(h, &S->t_start, S->t_bound + 1));
        TRY(dalloc(h, &S->t_state, S->t_bound + 1));
    }
    if (S->scratch) {
        TRY(dalloc(h, &S->d_masked, (size_t)m));
        TRY(dalloc(h, &S->d_prod_ptr, (size_t)m + 1));
        TRY(dalloc(h, &S->d_tcol, (size_t)S->scratch_products));
        TRY(dalloc(h, &S->d_tval, (size_t)S->scratch_products));
        S->begin_rec("scratch_ptr", 1, (uint32_t)((m + 4095) / 4096), (uint64_t)m, 0);
        launch_mask_sorted(S->d_flops, m, S->scratch_lo, 0xffffffffu, S->d_masked, s);
        launch_scan_u32_i64(S->d_masked, m, S->d_prod_ptr, S->d_tiles, h->d_ctr, s);
        CUT(cudaGetLastError());
        S->kernels += 2;
        S->end_rec();
    }
    if (S->n_long && !dense_long) {
        TRY(dalloc(h, &S->d_aseq, (size_t)std::max<int64_t>(A.nnz, 1)));
        TRY(dalloc(h, &S->w_p, (size_t)S->n_long));
        TRY(dalloc(h, &S->w_u, (size_t)S->n_long));
        TRY(dalloc(h, &S->w_prod_off, (size_t)S->n_long + 1));
        TRY(dalloc(h, &S->w_unit_off, (size_t)S->n_long + 1));
        TRY(dalloc(h, &S->w_heads, (size_t)all_units));
        TRY(dalloc(h, &S->w_unit_row, (size_t)all_units));
        TRY(dalloc(h, &S->w_hoff, (size_t)all_units + 1));
        TRY(dalloc(h, &S->w_tiles, (size_t)wave_units * 4));   // 32-byte merge-tile descriptors
        TRY(dalloc(h, &S->d_tiles_side, scan_tile_state_words((int64_t)std::max<uint64_t>(all_units, S->n_long))));
        TRY(dalloc(h, &S->w_col, (size_t)wave_products));
        TRY(dalloc(h, &S->w_val, (size_t)wave_products));
    }

    // ---- first pass: every scratch row is expanded, sorted and summed exactly once ---------------------------
    cudaStream_t sh = (h->opts.flags & SPADA_B200_FLAG_SERIAL) ? s : h->side;
    if (sh != s) {
        CUT(cudaEventRecord(h->ev_fork, s));
        CUT(cudaStreamWaitEvent(sh, h->ev_fork, 0));
        S->forked = true;
    }
    if (dense_long) {
        S->begin_rec("long_dense", 2, S->n_long, S->n_long, S->long_products, sh);
        launch_dense_rows(A, B, (int64_t)row_begin, S->perm_of(BIN_LONG0), S->n_long, S->d_prod_ptr, S->d_tcol, S->d_tval,
                          S->d_nnz, sh);
        CUT(cudaGetLastError());
        S->kernels += 1;
        S->end_rec();
    } else if (S->n_long) {
        const uint32_t* long_list = S->perm_of(BIN_LONG0);   // bins >= BIN_LONG0 are adjacent in perm[]
        S->begin_rec("long_prefix", 2, S->n_long, S->n_long, S->long_products, sh);
        launch_long_prefix(A, (int64_t)row_begin, long_list, S->n_long, S->d_blen, S->d_aseq, sh);
        CUT(cudaGetLastError());
        S->kernels += 1;
        S->end_rec();
        LongPlan& LP = S->long_plan;
        LP.rows_list = long_list;
        LP.n_rows = S->n_long;
        LP.unit_bound = all_units;
        LP.p = S->w_p;
        LP.u = S->w_u;
        LP.prod_off = S->w_prod_off;
        LP.unit_off = S->w_unit_off;
        LP.unit_row = S->w_unit_row;
        LP.unit_heads = S->w_heads;
        LP.unit_hoff = S->w_hoff;
        LP.t_ptr = S->d_prod_ptr;
        LP.s_col = S->d_tcol;
        LP.s_val = S->d_tval;
        LP.pong_col = S->w_col;
        LP.pong_val = S->w_val;
        LP.tiles = S->w_tiles;
        LP.tile_state = S->d_tiles_side;
        S->begin_rec("long_setup", 2, (S->n_long + 255) / 256, S->n_long, 0, sh);
        S->kernels += launch_long_setup(LP, S->d_flops, h->d_ctr_side, sh);
        CUT(cudaGetLastError());
        S->end_rec();
        const bool detail = waves.size() <= 2;   // per-stage records; beyond two waves one record for all of them
        int wi = 0;
        for (auto& w : waves) {
            LongWaveRange W{};
            W.lo = w.lo;
            W.hi = w.hi;
            memcpy(W.level_lo, w.level_lo, sizeof(W.level_lo));
            memcpy(W.level_grid, w.level_grid, sizeof(W.level_grid));
            memcpy(W.level_products, w.level_products, sizeof(W.level_products));
            W.products_bound = w.products_bound;
            W.max_level = w.max_level;
            W.unit_bound = w.unit_bound;
            LongStages stages;
            stages.on = [&](const char* what, uint32_t grid, uint64_t products) {
                char name[32];
                if (waves.size() > 1) snprintf(name, sizeof(name), "%s#%d", what, wi);
                else snprintf(name, sizeof(name), "%s", what);
                if (detail) S->begin_rec(name, 2, grid, W.hi - W.lo, products, sh);
            };
            stages.off = [&]() {
                if (detail) S->end_rec();
            };
            if (!detail && wi == 0) S->begin_rec("long_waves", 2, (uint32_t)w.unit_bound, S->n_long, S->long_products, sh);
            S->kernels += launch_long_wave(A, B, (int64_t)row_begin, S->d_aseq, LP, W, S->d_nnz, sh, &stages);
            CUT(cudaGetLastError());
            ++wi;
        }
        if (!detail) S->end_rec();
        S->kernels += launch_long_heads_scan(LP, h->d_ctr_side, sh);
        CUT(cudaGetLastError());
    }
    for (int bnum = fused ? 6 : 1; bnum <= 8; ++bnum) {
        const uint32_t rows = pc.bin_rows[bnum];
        if (!rows) continue;
        char name[32];
        snprintf(name, sizeof(name), "sort_pass<%s>", bin_name(bnum));
        S->begin_rec(name, 2, (uint32_t)esc_grid(bnum, rows), rows, pc.bin_products[bnum], s);
        launch_esc_numeric(bnum, A, B, (int64_t)row_begin, S->perm_of(bnum), rows, S->d_prod_ptr, S->d_tcol, S->d_tval, s,
                           S->d_nnz);
        CUT(cudaGetLastError());
        S->kernels += 1;
        S->end_rec();
    }
    if (S->forked) {
        CUT(cudaEventRecord(h->ev_join, sh));
        CUT(cudaStreamWaitEvent(s, h->ev_join, 0));
        S->forked = false;
    }
    if (!fused) {
        // ---- stage 4: row_ptr ---------------------------------------------------------------
        S->begin_rec("row_ptr_scan", 4, (uint32_t)((m + 4095) / 4096), (uint64_t)m, 0);
        launch_scan_u32_i64(S->d_nnz, m, R->ptr, S->d_tiles, h->d_ctr, s);
        CUT(cudaGetLastError());
        S->kernels += 1;
        S->end_rec();
        if (host_nnz) {
            launch_publish(R->ptr + m, h->h_scalar, sizeof(int64_t), s);
            CUT(cudaStreamSynchronize(s));  // host read-back #2: nnz(C) sizes the output
            S->nnz_c = h->h_scalar[0];
        } else {
            S->nnz_c = -1;   // stays on the device (sharded runs size C by its bound)
        }
    }
    *out = S;
    return 0;
}

// second half.  dst == nullptr: C is allocated here (exact size, or the product bound in single-pass mode) and is
// the only destination.  Otherwise dst lists the C buffers of every GPU (this GPU's first) and the shard's rows go to
// dst->off + local offset in all of them; row_ptr_dst[d] (nullable) get the shard's row pointers shifted by dst->off
// at row offset row_off.
int shard_finish(spada_b200_shard* S, const CopyDst* dst, const RowPtrDst* row_ptr_dst, int64_t row_off,
                 spada_b200_result_t** out, spada_b200_stats* stats_out = nullptr) {
    spada_b200* h = S->h;
    DeviceGuard g(h->device);
    cudaStream_t s = h->stream;
    spada_b200_result* R = S->R;
    spada_b200_stats& st = R->stats;
    const PlanCounters& pc = S->pc;
    const int64_t m = S->m;
    int rc;
    if (m == 0) {
        if (!dst) {
            TRY(dalloc(h, &R->col, 1));
            TRY(dalloc(h, &R->val, 1));
        }
        CUT(cudaStreamSynchronize(s));
        if (out) {
            *out = R;
            S->R = nullptr;
        }
        shard_free(S);
        return 0;
    }
    CopyDst D{};
    if (dst) {
        D = *dst;
    } else {
        const size_t cap = (size_t)(S->fused ? pc.total_products : (uint64_t)S->nnz_c);
        TRY(dalloc(h, &R->col, cap));
        TRY(dalloc(h, &R->val, cap));
        D.col[0] = R->col;
        D.val[0] = R->val;
        D.n = 1;
        D.off = 0;
    }
    if (S->fused) {
        // ---- stages 2+3+4 fused for the warp-per-row bins: straight into C ----------------------------------
        char name[32];
        snprintf(name, sizeof(name), S->tile_mode ? "tile_pass<%s>" : "fused<%s>", bin_name(S->max_light_bin));
        uint64_t light_products = 0;
        for (int bnum = 1; bnum <= 5; ++bnum) light_products += pc.bin_products[bnum];
        S->begin_rec(name, 3, S->tile_mode ? (uint32_t)S->t_bound : (uint32_t)((m + 7) / 8), S->light_rows, light_products);
        if (S->tile_mode) {
            launch_tile_pass(S->A, S->B, S->row_begin, m, S->d_flops, S->d_nnz, R->ptr, R->col, R->val, S->t_w, S->t_lp,
                             S->t_idx, S->t_start, S->t_state, S->t_bound, S->d_tiles, h->d_ctr, s);
            S->kernels += 6;
        } else {
            launch_fused_light(S->max_light_bin, S->window.tiny_quad, S->A, S->B, S->row_begin, m, S->d_flops, S->d_nnz,
                               R->ptr, R->col, R->val, S->d_tiles, h->d_ctr, s);
            S->kernels += 1;
        }
        CUT(cudaGetLastError());
        S->end_rec();
        launch_publish(R->ptr + m, h->h_scalar, sizeof(int64_t), s);
    }
    // ---- placement: scratch rows -> C (and the peers' C) --------------------------------------------------
    if (S->scratch) {
        S->begin_rec(D.n > 1 ? "copy_gather" : "copy_rows", 3, (uint32_t)((m + 7) / 8), (uint64_t)m, S->scratch_products);
        launch_copy_rows(S->d_flops, m, S->scratch_lo, ESC_MAX_PRODUCTS, S->d_prod_ptr, S->d_tcol, S->d_tval, R->ptr, D, s);
        S->kernels += 1;
        if (S->n_long && S->dense_long) {   // long rows finished by the dense path: one CTA per scratch row
            launch_copy_rows_list(S->perm_of(BIN_LONG0), S->n_long, S->d_prod_ptr, S->d_tcol, S->d_tval, R->ptr, D, s);
            S->kernels += 1;
        }
        CUT(cudaGetLastError());
        S->end_rec();
        if (S->n_long && !S->dense_long) {
            // long rows: their scratch rows hold the sorted products; the left-to-right sums go straight into C
            S->begin_rec(D.n > 1 ? "long_reduce_gather" : "long_reduce", 3, (uint32_t)S->long_plan.unit_bound, S->n_long,
                         S->long_products);
            S->kernels += launch_long_reduce(S->long_plan, R->ptr, D, s);
            CUT(cudaGetLastError());
            S->end_rec();
        }
    }
    if (dst && row_ptr_dst) {
        S->begin_rec("row_ptr_gather", 3, (uint32_t)((m + 255) / 256), (uint64_t)m, 0);
        launch_shift_row_ptr(R->ptr, m, D.off, D.shard_nnz, D.shard_idx, *row_ptr_dst, row_off, s);
        CUT(cudaGetLastError());
        S->kernels += 1;
        S->end_rec();
    }
    S->e_end = next_event(h, s);
    S->release_work();
    CUT(cudaStreamSynchronize(s));
    CUT(cudaGetLastError());
    if (S->fused) S->nnz_c = h->h_scalar[0];
    R->nnz = S->nnz_c < 0 ? 0 : (uint64_t)S->nnz_c;
    st.nnz_c = R->nnz;

    // ---- stats ------------------------------------------------------------------------------
    st.n_launches = S->kernels;
    st.n_recorded = 0;
    for (const LaunchRec& r : S->recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (r.stage == 1) st.ms_flops += ms;
        if (r.stage == 2) st.ms_symbolic += ms;
        if (r.stage == 3) st.ms_numeric += ms;
        if (r.stage == 4) st.ms_scan += ms;
        if (st.n_recorded < SPADA_B200_MAX_LAUNCHES) {
            spada_b200_launch& L = st.launches[st.n_recorded++];
            memcpy(L.name, r.name, sizeof(L.name));
            L.ms = ms;
            L.grid = r.grid;
            L.rows = r.rows;
            L.products = r.products;
            L.nnz = 0;
        }
    }
    if (!S->recs.empty()) cudaEventElapsedTime(&st.ms_total, S->recs.front().e0, S->e_end);
    if (stats_out) *stats_out = st;
    if (out) {
        *out = R;
        S->R = nullptr;
    }
    S->finished = true;
    shard_free(S);
    return 0;
}
#undef TRY
#undef CUT

}  // namespace

extern "C" int spada_b200_spgemm_dev(spada_b200_t* h, const spada_b200_csr_t* a, const spada_b200_csr_t* b,
                                     uint64_t row_begin, uint64_t row_end, spada_b200_result_t** out) {
    if (!out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    spada_b200_shard* S = nullptr;
    int rc = shard_begin(h, a, b, row_begin, row_end, false, true, &S);
    if (rc) return rc;
    return shard_finish(S, nullptr, nullptr, 0, out);
}

// ---- sharded runs: A row-sharded over several GPUs, B replicated, C gathered on every GPU ------------------
// (SURVEY.md 8e).  Every GPU owns a full-size set of C buffers (spada_b200_cbuf); the second half of a shard's
// product writes the shard's rows into ALL of them at the shard's global offset -- its own through HBM, the peers'
// through NVLink peer mappings (CUDA IPC between the processes of a one-process-per-GPU run, peer access inside one
// process) -- so the all-gather of C is the store itself, not a pass after it.
struct spada_b200_cbuf {
    spada_b200* h;
    uint64_t rows, cols, cap;
    int64_t* ptr;
    int32_t* col;
    double* val;
    bool imported;
};

extern "C" int spada_b200_cbuf_create(spada_b200_t* h, uint64_t rows, uint64_t cols, uint64_t capacity_nnz,
                                      spada_b200_cbuf_t** out) {
    if (!h || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    DeviceGuard g(h->device);
    spada_b200_cbuf* c = new (std::nothrow) spada_b200_cbuf{h, rows, cols, capacity_nnz, nullptr, nullptr, nullptr, false};
    if (!c) return fail(SPADA_B200_OOM, "host allocation failed");
    int rc;
    if ((rc = dalloc(h, &c->ptr, (size_t)rows + 1)) || (rc = dalloc(h, &c->col, (size_t)std::max<uint64_t>(capacity_nnz, 1))) ||
        (rc = dalloc(h, &c->val, (size_t)std::max<uint64_t>(capacity_nnz, 1)))) {
        dfree(h, c->ptr);
        dfree(h, c->col);
        dfree(h, c->val);
        delete c;
        return rc;
    }
    if (cudaMemsetAsync(c->ptr, 0, ((size_t)rows + 1) * sizeof(int64_t), h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
        spada_b200_cbuf_free(c);
        return fail(SPADA_B200_CUDA_ERROR, "cbuf_create: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = c;
    return 0;
}

extern "C" int spada_b200_cbuf_export(const spada_b200_cbuf_t* c, void* handles) {
    if (!c || !handles) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    if (c->imported) return fail(SPADA_B200_INVALID_ARG, "an imported buffer cannot be exported again");
    DeviceGuard g(c->h->device);
    cudaIpcMemHandle_t* hs = reinterpret_cast<cudaIpcMemHandle_t*>(handles);
    static_assert(sizeof(cudaIpcMemHandle_t) == SPADA_B200_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaIpcGetMemHandle(&hs[0], c->ptr));
    CU(cudaIpcGetMemHandle(&hs[1], c->col));
    CU(cudaIpcGetMemHandle(&hs[2], c->val));
    return 0;
}

extern "C" int spada_b200_cbuf_import(spada_b200_t* h, const void* handles, uint64_t rows, uint64_t cols,
                                      uint64_t capacity_nnz, spada_b200_cbuf_t** out) {
    if (!h || !handles || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    DeviceGuard g(h->device);
    const cudaIpcMemHandle_t* hs = reinterpret_cast<const cudaIpcMemHandle_t*>(handles);
    spada_b200_cbuf* c = new (std::nothrow) spada_b200_cbuf{h, rows, cols, capacity_nnz, nullptr, nullptr, nullptr, true};
    if (!c) return fail(SPADA_B200_OOM, "host allocation failed");
    void* p[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 3; ++i) {
        cudaError_t e = cudaIpcOpenMemHandle(&p[i], hs[i], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int j = 0; j < i; ++j) cudaIpcCloseMemHandle(p[j]);
            delete c;
            cudaGetLastError();
            return fail(SPADA_B200_CUDA_ERROR, "cudaIpcOpenMemHandle: %s (peer buffers need NVLink / P2P access)",
                        cudaGetErrorString(e));
        }
    }
    c->ptr = (int64_t*)p[0];
    c->col = (int32_t*)p[1];
    c->val = (double*)p[2];
    *out = c;
    return 0;
}

extern "C" void spada_b200_cbuf_free(spada_b200_cbuf_t* c) {
    if (!c) return;
    DeviceGuard g(c->h->device);
    if (c->imported) {
        cudaIpcCloseMemHandle(c->ptr);
        cudaIpcCloseMemHandle(c->col);
        cudaIpcCloseMemHandle(c->val);
        cudaGetLastError();
    } else {
        cudaStreamSynchronize(c->h->stream);
        dfree(c->h, c->ptr);
        dfree(c->h, c->col);
        dfree(c->h, c->val);
    }
    delete c;
}

extern "C" int spada_b200_cbuf_device_ptrs(const spada_b200_cbuf_t* c, const int64_t** d_indptr, const int32_t** d_indices,
                                           const double** d_data) {
    if (!c) return fail(SPADA_B200_INVALID_ARG, "buffer is NULL");
    if (d_indptr) *d_indptr = c->ptr;
    if (d_indices) *d_indices = c->col;
    if (d_data) *d_data = c->val;
    return 0;
}

extern "C" int spada_b200_cbuf_nnz(const spada_b200_cbuf_t* c, uint64_t* nnz) {
    if (!c || !nnz) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    DeviceGuard g(c->h->device);
    int64_t v = 0;
    CU(cudaMemcpyAsync(&v, c->ptr + c->rows, sizeof(int64_t), cudaMemcpyDeviceToHost, c->h->stream));
    CU(cudaStreamSynchronize(c->h->stream));
    *nnz = (uint64_t)v;
    return 0;
}

extern "C" int spada_b200_cbuf_copy32(const spada_b200_cbuf_t* c, int64_t* indptr, int32_t* indices, double* data) {
    if (!c || !indptr) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    DeviceGuard g(c->h->device);
    cudaStream_t s = c->h->stream;
    CU(cudaMemcpyAsync(indptr, c->ptr, (c->rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const uint64_t nnz = (uint64_t)indptr[c->rows];
    if (nnz > c->cap) return fail(SPADA_B200_INVALID_ARG, "row_ptr[rows] = %llu exceeds the buffer's capacity", (unsigned long long)nnz);
    if (nnz && indices) CU(cudaMemcpyAsync(indices, c->col, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (nnz && data) CU(cudaMemcpyAsync(data, c->val, nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int spada_b200_shard_begin(spada_b200_t* h, const spada_b200_csr_t* a, const spada_b200_csr_t* b,
                                      uint64_t row_begin, uint64_t row_end, int64_t* d_nnz_local, uint64_t* nnz_local,
                                      spada_b200_shard_t** out) {
    if (!out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    spada_b200_shard* S = nullptr;
    int rc = shard_begin(h, a, b, row_begin, row_end, true, nnz_local != nullptr, &S);
    if (rc) return rc;
    if (nnz_local) *nnz_local = (uint64_t)S->nnz_c;
    if (d_nnz_local) {
        DeviceGuard g(h->device);
        cudaError_t e = cudaMemcpyAsync(d_nnz_local, S->R->ptr + S->m, sizeof(int64_t), cudaMemcpyDeviceToDevice, h->stream);
        if (e != cudaSuccess) {
            shard_free(S);
            return fail(SPADA_B200_CUDA_ERROR, "shard_begin: %s", cudaGetErrorString(e));
        }
    }
    *out = S;
    return 0;
}

extern "C" int spada_b200_shard_finish(spada_b200_shard_t* S, spada_b200_cbuf_t* const* bufs, uint32_t n_bufs,
                                       uint64_t nnz_offset, const int64_t* d_shard_nnz, uint32_t shard_index,
                                       spada_b200_stats* stats) {
    if (!S || !bufs || n_bufs == 0 || n_bufs > 8) {
        if (S) shard_free(S);
        return fail(SPADA_B200_INVALID_ARG, "shard_finish: 1..8 destination buffers");
    }
    CopyDst D{};
    RowPtrDst P{};
    for (uint32_t i = 0; i < n_bufs; ++i) {
        if (!bufs[i] || bufs[i]->rows < (uint64_t)(S->row_begin + S->m)) {
            shard_free(S);
            return fail(SPADA_B200_INVALID_ARG, "shard_finish: destination buffer %u is missing or too small", i);
        }
        D.col[i] = bufs[i]->col;
        D.val[i] = bufs[i]->val;
        P.ptr[i] = bufs[i]->ptr;
    }
    D.n = P.n = (int)n_bufs;
    D.off = (int64_t)nnz_offset;
    D.shard_nnz = d_shard_nnz;
    D.shard_idx = (int)shard_index;
    return shard_finish(S, &D, &P, S->row_begin, nullptr, stats);
}

extern "C" void spada_b200_shard_abort(spada_b200_shard_t* S) { shard_free(S); }

// ---- row-panel streaming (the step after main.rs:113-116, which prints ten rows and drops the rest) ----
// C is produced in panels of consecutive rows.  Every finished panel is copied to pinned host memory on a copy stream
// while the next panel is computed, and handed to the caller's sink: the D2H of C (the bulk of an end-to-end run: 12
// bytes per output entry over PCIe) overlaps the computation, and C never has to fit in HBM -- or be kept at all.
namespace {
struct PanelBuf {
    int64_t* ptr = nullptr;   // pinned
    int32_t* col = nullptr;
    double* val = nullptr;
    cudaEvent_t done = nullptr;
    spada_b200_result* R = nullptr;
    uint64_t row_begin = 0, row_end = 0, nnz_begin = 0;
};
}  // namespace

namespace {
// dst_*: nullable.  With destination arrays the panels are copied straight into them at their global offsets (no
// staging, no sink); otherwise into two pinned staging buffers that take turns and are handed to the sink.
// feed (nullable, direct mode only): panel bounds fixed by the caller and a hook that runs before panel p is computed
// (the host-to-host entry point makes the engine's stream wait for the upload of A's rows of that panel there).
struct PanelFeed {
    std::vector<uint64_t> bounds;
    std::function<int(size_t)> before;
};
int stream_panels(spada_b200_t* h, const spada_b200_csr_t* a, const spada_b200_csr_t* b, uint64_t panel_products,
                  spada_b200_panel_sink sink, void* user, int64_t* dst_ptr, int32_t* dst_col, double* dst_val,
                  uint64_t dst_capacity, spada_b200_stream_stats* stats, const PanelFeed* feed = nullptr) {
    if (a->d.cols != b->d.rows)
        return fail(SPADA_B200_DIM_MISMATCH, "A is %lld x %lld but B has %lld rows", (long long)a->d.rows,
                    (long long)a->d.cols, (long long)b->d.rows);
    DeviceGuard g(h->device);
    const bool direct = dst_ptr != nullptr;
    const uint64_t m = (uint64_t)a->d.rows;
    std::vector<uint64_t> bounds{0};
    uint64_t total = 0, max_panel = 0, max_rows = 0;
    int rc = 0;
    if (feed) {
        bounds = feed->bounds;
        for (size_t p = 0; p + 1 < bounds.size(); ++p) max_rows = std::max(max_rows, bounds[p + 1] - bounds[p]);
    } else {
        std::vector<uint64_t> f((size_t)m);
        if ((rc = spada_b200_flops(h, a, b, &total, f.data()))) return rc;
        // panels: consecutive rows up to panel_products intermediate products each (a row heavier than that is a panel
        // of its own).  Default: what keeps scratch rows + C + staging of a panel within a quarter of the device, at
        // least eight panels so that the copies have something to overlap with.
        if (panel_products == 0) {
            const uint64_t by_mem = (uint64_t)(0.25 * (double)h->dev_total_mem / 36.0);
            panel_products = std::max<uint64_t>(1u << 20, std::min<uint64_t>(by_mem, total / 8 + 1));
        }
        uint64_t acc = 0;
        for (uint64_t r = 0; r < m; ++r) {
            if (acc && acc + f[(size_t)r] > panel_products) {
                max_panel = std::max(max_panel, acc);
                max_rows = std::max(max_rows, r - bounds.back());
                bounds.push_back(r);
                acc = 0;
            }
            acc += f[(size_t)r];
        }
        max_panel = std::max(max_panel, acc);
        max_rows = std::max(max_rows, m - bounds.back());
        bounds.push_back(m);
    }
    const size_t n_panels = bounds.size() - 1;
    cudaStream_t cs = nullptr;
    PanelBuf buf[2];
    float ms_total = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        for (auto& B : buf) {
            if (B.done) cudaEventSynchronize(B.done);
            if (B.R) spada_b200_result_free(B.R);
            if (!direct) {
                if (B.ptr) cudaFreeHost(B.ptr);
                if (B.col) cudaFreeHost(B.col);
                if (B.val) cudaFreeHost(B.val);
            }
            if (B.done) cudaEventDestroy(B.done);
        }
        if (cs) cudaStreamDestroy(cs);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        cudaGetLastError();
    };
    auto body = [&]() -> int {
        CU(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        for (auto& B : buf) {
            if (!direct) {
                CU(cudaMallocHost((void**)&B.ptr, (size_t)(max_rows + 1) * sizeof(int64_t)));
                CU(cudaMallocHost((void**)&B.col, (size_t)std::max<uint64_t>(max_panel, 1) * sizeof(int32_t)));
                CU(cudaMallocHost((void**)&B.val, (size_t)std::max<uint64_t>(max_panel, 1) * sizeof(double)));
            }
            CU(cudaEventCreateWithFlags(&B.done, cudaEventDisableTiming));
        }
        CU(cudaEventRecord(e0, h->stream));
        uint64_t nnz_done = 0;
        auto drain = [&](PanelBuf& B) -> int {   // waits for the panel's copy; staging mode: hands it to the sink
            if (!B.R) return 0;
            CU(cudaEventSynchronize(B.done));
            int src = 0;
            if (!direct) {
                const uint64_t rows = B.row_end - B.row_begin;
                for (uint64_t i = 0; i <= rows; ++i) B.ptr[i] += (int64_t)B.nnz_begin;   // global offsets
                src = sink(user, B.row_begin, B.row_end, B.nnz_begin, B.ptr, B.col, B.val);
            }
            spada_b200_result_free(B.R);
            B.R = nullptr;
            if (src) return fail(SPADA_B200_INVALID_ARG, "the panel sink returned %d for rows [%llu, %llu)", src,
                                 (unsigned long long)B.row_begin, (unsigned long long)B.row_end);
            return 0;
        };
        for (size_t p = 0; p < n_panels; ++p) {
            PanelBuf& B = buf[p & 1];
            int rc2 = drain(B);    // the buffer's previous panel (p - 2): its copy ran beside the computation of p - 1
            if (rc2) return rc2;
            spada_b200_result* R = nullptr;
            if (feed && feed->before && (rc2 = feed->before(p))) return rc2;
            if ((rc2 = spada_b200_spgemm_dev(h, a, b, bounds[p], bounds[p + 1], &R))) return rc2;   // synchronous
            total += feed ? R->stats.products : 0;
            B.R = R;
            B.row_begin = bounds[p];
            B.row_end = bounds[p + 1];
            B.nnz_begin = nnz_done;
            nnz_done += R->nnz;
            // the copy of panel p runs on the copy stream while panel p + 1 is computed
            int64_t* hp = B.ptr;
            int32_t* hc = B.col;
            double* hv = B.val;
            size_t n_ptr = (size_t)R->rows + 1;
            if (direct) {
                if (nnz_done > dst_capacity)
                    return fail(SPADA_B200_INVALID_ARG, "C has more than the %llu entries the destination arrays hold",
                                (unsigned long long)dst_capacity);
                // global offsets on the device, then every array straight to its place (row 0's pointer comes with panel 0)
                RowPtrDst self{};
                self.ptr[0] = R->ptr;
                self.n = 1;
                launch_shift_row_ptr(R->ptr, (int64_t)R->rows, (int64_t)B.nnz_begin, nullptr, 0, self, 0, h->stream);
                CU(cudaGetLastError());
                CU(cudaStreamSynchronize(h->stream));
                hp = dst_ptr + B.row_begin;
                hc = dst_col + B.nnz_begin;
                hv = dst_val + B.nnz_begin;
            }
            CU(cudaMemcpyAsync(hp, R->ptr, n_ptr * sizeof(int64_t), cudaMemcpyDeviceToHost, cs));
            if (R->nnz) {
                CU(cudaMemcpyAsync(hc, R->col, (size_t)R->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
                CU(cudaMemcpyAsync(hv, R->val, (size_t)R->nnz * sizeof(double), cudaMemcpyDeviceToHost, cs));
            }
            CU(cudaEventRecord(B.done, cs));
        }
        for (size_t q = 0; q < 2; ++q) {
            int rc2 = drain(buf[(n_panels + q) & 1]);
            if (rc2) return rc2;
        }
        if (direct && m == 0) dst_ptr[0] = 0;
        CU(cudaEventRecord(e1, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        cudaEventElapsedTime(&ms_total, e0, e1);
        if (stats) {
            stats->panels = n_panels;
            stats->products = total;
            stats->nnz_c = nnz_done;
            stats->max_panel_products = max_panel;
            stats->ms_total = ms_total;
        }
        return 0;
    };
    rc = body();
    cleanup();
    return rc;
}
}  // namespace

extern "C" int spada_b200_spgemm_stream(spada_b200_t* h, const spada_b200_csr_t* a, const spada_b200_csr_t* b,
                                        uint64_t panel_products, spada_b200_panel_sink sink, void* user,
                                        spada_b200_stream_stats* stats) {
    if (!h || !a || !b || !sink) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    return stream_panels(h, a, b, panel_products, sink, user, nullptr, nullptr, nullptr, 0, stats);
}

extern "C" int spada_b200_spgemm_to_host(spada_b200_t* h, const spada_b200_csr_t* a, const spada_b200_csr_t* b,
                                         uint64_t panel_products, int64_t* indptr, int32_t* indices, double* data,
                                         uint64_t capacity_nnz, spada_b200_stream_stats* stats) {
    if (!h || !a || !b || !indptr || (capacity_nnz && (!indices || !data))) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    return stream_panels(h, a, b, panel_products, nullptr, nullptr, indptr, indices, data, capacity_nnz, stats);
}

// Host operands in, whole C in host arrays, everything overlapped that PCIe allows: B goes up first (every row needs all
// of it), then A's column ids and values follow in row panels on an upload stream while the panels already there are
// computed and their results travel down (PCIe is full duplex) -- what the Rust wrapper calls for one product.
extern "C" int spada_b200_spgemm32_host_to_host(spada_b200_t* h, const spada_csr_view32* a, const spada_csr_view32* b,
                                                int64_t* indptr, int32_t* indices, double* data, uint64_t capacity_nnz,
                                                spada_b200_stream_stats* stats) {
    if (!h || !a || !b || !indptr || (capacity_nnz && (!indices || !data))) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    if (a->cols != b->rows)
        return fail(SPADA_B200_DIM_MISMATCH, "A is %llu x %llu but B has %llu rows", (unsigned long long)a->rows,
                    (unsigned long long)a->cols, (unsigned long long)b->rows);
    DeviceGuard g(h->device);
    int rc;
    const bool alias = (const void*)a == (const void*)b || (a->indptr == b->indptr && a->indices == b->indices &&
                                                            a->data == b->data && a->rows == b->rows && a->cols == b->cols);
    spada_b200_csr_t *da = nullptr, *db = nullptr;
    if ((rc = spada_b200_upload32(h, b, &db))) return rc;
    db->fib_ready = true;   // one product: no fiber store
    if (alias || a->nnz == 0 || a->rows == 0) {
        // A is B (or empty): nothing to overlap
        if (!alias && (rc = spada_b200_upload32(h, a, &da))) {
            spada_b200_csr_free(db);
            return rc;
        }
        rc = stream_panels(h, alias ? db : da, db, 0, nullptr, nullptr, indptr, indices, data, capacity_nnz, stats);
        if (da) spada_b200_csr_free(da);
        spada_b200_csr_free(db);
        return rc;
    }
    if ((rc = check_csr_args(a->rows, a->cols, a->nnz, a->indptr, a->indices, a->data)) ||
        (a->nnz >= (1ull << 31) ? (rc = fail(SPADA_B200_TOO_LARGE, "nnz >= 2^31 needs the 64-bit view")) : 0) ||
        ((uint64_t)(int64_t)a->indptr[a->rows] != a->nnz ? (rc = fail(SPADA_B200_INVALID_ARG, "indptr[rows] != nnz")) : 0)) {
        spada_b200_csr_free(db);
        return rc;
    }
    int64_t* ptr;
    int32_t* col;
    double* val;
    if ((rc = make_csr(h, a->rows, a->cols, a->nnz, &da, &ptr, &col, &val))) {
        spada_b200_csr_free(db);
        return rc;
    }
    da->fib_ready = true;
    cudaStream_t up = nullptr;
    std::vector<cudaEvent_t> ev;
    int32_t* tmp = nullptr;
    auto body = [&]() -> int {
        int rc2;
        if ((rc2 = dalloc(h, &tmp, a->rows + 1))) return rc2;
        CU(cudaMemcpyAsync(tmp, a->indptr, (a->rows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        launch_widen_i32(tmp, (int64_t)a->rows + 1, ptr, h->stream);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
        // panels of equal nnz(A): 8 or more, cut on the host from the row pointers it already holds
        const uint64_t n_panels = std::max<uint64_t>(8, std::min<uint64_t>(64, a->nnz / (8u << 20) + 1));
        PanelFeed feed;
        feed.bounds.push_back(0);
        for (uint64_t p = 1; p < n_panels; ++p) {
            const int32_t target = (int32_t)(a->nnz * p / n_panels);
            const int32_t* it = std::lower_bound(a->indptr, a->indptr + a->rows + 1, target);
            const uint64_t r = std::min<uint64_t>((uint64_t)(it - a->indptr), a->rows);
            if (r > feed.bounds.back()) feed.bounds.push_back(r);
        }
        if (feed.bounds.back() < a->rows) feed.bounds.push_back(a->rows);
        CU(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        ev.resize(feed.bounds.size() - 1, nullptr);
        for (size_t p = 0; p + 1 < feed.bounds.size(); ++p) {
            const int64_t lo = a->indptr[feed.bounds[p]], hi = a->indptr[feed.bounds[p + 1]];
            if (hi > lo) {
                CU(cudaMemcpyAsync(col + lo, a->indices + lo, (size_t)(hi - lo) * sizeof(int32_t), cudaMemcpyHostToDevice, up));
                CU(cudaMemcpyAsync(val + lo, a->data + lo, (size_t)(hi - lo) * sizeof(double), cudaMemcpyHostToDevice, up));
            }
            CU(cudaEventCreateWithFlags(&ev[p], cudaEventDisableTiming));
            CU(cudaEventRecord(ev[p], up));
        }
        feed.before = [&](size_t p) -> int {
            CU(cudaStreamWaitEvent(h->stream, ev[p], 0));
            return 0;
        };
        if ((rc2 = stream_panels(h, da, db, 0, nullptr, nullptr, indptr, indices, data, capacity_nnz, stats, &feed))) return rc2;
        // the reference's loaders hand over canonical CSR; a violation is reported once all of A is on the device
        if (h->opts.flags & SPADA_B200_FLAG_VALIDATE) return validate_device_csr(h, da->d);
        return 0;
    };
    rc = body();
    if (up) {
        cudaStreamSynchronize(up);
        cudaStreamDestroy(up);
    }
    for (cudaEvent_t e : ev)
        if (e) cudaEventDestroy(e);
    dfree(h, tmp);
    spada_b200_csr_free(da);
    spada_b200_csr_free(db);
    return rc;
}

// ---- all GPUs of one process (what the spada-sim CLI, a single process, drives: SURVEY.md 8b) -----------------------
// One engine handle per device, peer access between every pair, one host thread per device for the two halves.
struct spada_b200_group {
    std::vector<spada_b200*> h;
    std::vector<spada_b200_cbuf*> bufs;
    uint64_t buf_rows = 0, buf_cap = 0;
};

extern "C" void spada_b200_group_destroy(spada_b200_group_t* G) {
    if (!G) return;
    for (auto* c : G->bufs) spada_b200_cbuf_free(c);
    for (auto* h : G->h) spada_b200_destroy(h);
    delete G;
}

extern "C" int spada_b200_group_create(const spada_b200_opts* opts, uint32_t n_gpus, spada_b200_group_t** out) {
    if (!out || n_gpus == 0 || n_gpus > 8) return fail(SPADA_B200_INVALID_ARG, "group_create: 1..8 GPUs");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SPADA_B200_NO_DEVICE, "no CUDA device; this engine has no CPU fallback");
    }
    if ((int)n_gpus > n) return fail(SPADA_B200_INVALID_ARG, "group_create: %u GPUs asked, %d present", n_gpus, n);
    spada_b200_group* G = new (std::nothrow) spada_b200_group;
    if (!G) return fail(SPADA_B200_OOM, "host allocation failed");
    for (uint32_t d = 0; d < n_gpus; ++d) {
        spada_b200_opts o{};
        if (opts) o = *opts;
        else {
            o.accelerator = SPADA_B200_ACC_SPADA;
            o.flags = SPADA_B200_FLAG_VALIDATE;
        }
        o.device = (int32_t)d;
        o.stream = nullptr;
        spada_b200* h = nullptr;
        int rc = spada_b200_create(&o, &h);
        if (rc) {
            spada_b200_group_destroy(G);
            return rc;
        }
        G->h.push_back(h);
    }
    int prev = 0;
    cudaGetDevice(&prev);
    for (uint32_t i = 0; i < n_gpus; ++i)
        for (uint32_t j = 0; j < n_gpus; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, (int)i, (int)j);
            if (!can) {
                cudaSetDevice(prev);
                spada_b200_group_destroy(G);
                return fail(SPADA_B200_CUDA_ERROR, "device %u cannot map device %u's memory (no NVLink / P2P)", i, j);
            }
            cudaSetDevice((int)i);
            cudaError_t e = cudaDeviceEnablePeerAccess((int)j, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaSetDevice(prev);
                spada_b200_group_destroy(G);
                return fail(SPADA_B200_CUDA_ERROR, "cudaDeviceEnablePeerAccess(%u -> %u): %s", i, j, cudaGetErrorString(e));
            }
            cudaGetLastError();
        }
    cudaSetDevice(prev);
    *out = G;
    return 0;
}

namespace {
// replicate a device operand of device 0 onto device d (NVLink peer copies)
int replicate(spada_b200* h0, const spada_b200_csr* src, spada_b200* hd, spada_b200_csr** out) {
    DeviceGuard g(hd->device);
    int64_t* ptr;
    int32_t* col;
    double* val;
    int rc = make_csr(hd, (uint64_t)src->d.rows, (uint64_t)src->d.cols, (uint64_t)src->d.nnz, out, &ptr, &col, &val);
    if (rc) return rc;
    cudaStream_t s = hd->stream;
    cudaError_t e = cudaMemcpyPeerAsync(ptr, hd->device, src->d.ptr, h0->device, (size_t)(src->d.rows + 1) * sizeof(int64_t), s);
    if (e == cudaSuccess && src->d.nnz)
        e = cudaMemcpyPeerAsync(col, hd->device, src->d.col, h0->device, (size_t)src->d.nnz * sizeof(int32_t), s);
    if (e == cudaSuccess && src->d.nnz)
        e = cudaMemcpyPeerAsync(val, hd->device, src->d.val, h0->device, (size_t)src->d.nnz * sizeof(double), s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        spada_b200_csr_free(*out);
        *out = nullptr;
        return fail(SPADA_B200_CUDA_ERROR, "replicating an operand to device %d: %s", hd->device, cudaGetErrorString(e));
    }
    return 0;
}

template <typename View, typename UploadFn>
int group_spgemm(spada_b200_group* G, const View* a, const View* b, spada_b200_result_t** out, UploadFn upload) {
    if (!G || !a || !b || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    const uint32_t N = (uint32_t)G->h.size();
    if (a->cols != b->rows)
        return fail(SPADA_B200_DIM_MISMATCH, "A is %llu x %llu but B has %llu rows", (unsigned long long)a->rows,
                    (unsigned long long)a->cols, (unsigned long long)b->rows);
    std::vector<spada_b200_csr*> A(N, nullptr), B(N, nullptr);
    std::vector<spada_b200_shard*> S(N, nullptr);
    auto cleanup = [&]() {
        for (uint32_t d = 0; d < N; ++d) {
            if (S[d]) shard_free(S[d]);
            if (B[d] && B[d] != A[d]) spada_b200_csr_free(B[d]);
            if (A[d]) spada_b200_csr_free(A[d]);
        }
    };
    int rc;
    const bool alias = (const void*)a == (const void*)b ||
                       (a->indptr == b->indptr && a->indices == b->indices && a->data == b->data && a->rows == b->rows &&
                        a->cols == b->cols);
    if ((rc = upload(G->h[0], a, &A[0]))) return rc;
    if (alias) B[0] = A[0];
    else if ((rc = upload(G->h[0], b, &B[0]))) { cleanup(); return rc; }
    for (uint32_t d = 1; d < N; ++d) {
        if ((rc = replicate(G->h[0], A[0], G->h[d], &A[d]))) { cleanup(); return rc; }
        if (alias) B[d] = A[d];
        else if ((rc = replicate(G->h[0], B[0], G->h[d], &B[d]))) { cleanup(); return rc; }
    }
    std::vector<uint64_t> bounds(N + 1, 0);
    uint64_t products = 0;
    if ((rc = spada_b200_plan_shards(G->h[0], A[0], B[0], N, bounds.data()))) { cleanup(); return rc; }
    products = G->h[0]->h_ctr->total_products;
    if (G->bufs.empty() || G->buf_rows != a->rows || G->buf_cap < products) {
        for (auto* c : G->bufs) spada_b200_cbuf_free(c);
        G->bufs.assign(N, nullptr);
        for (uint32_t d = 0; d < N; ++d)
            if ((rc = spada_b200_cbuf_create(G->h[d], a->rows, b->cols, products, &G->bufs[d]))) {
                for (auto*& c : G->bufs) { spada_b200_cbuf_free(c); c = nullptr; }
                G->bufs.clear();
                cleanup();
                return rc;
            }
        G->buf_rows = a->rows;
        G->buf_cap = products;
    }
    // first halves, one host thread per device
    std::vector<int> rcs(N, 0);
    std::vector<std::string> errs(N);
    std::vector<uint64_t> nnz(N, 0);
    {
        std::vector<std::thread> th;
        for (uint32_t d = 0; d < N; ++d)
            th.emplace_back([&, d]() {
                rcs[d] = spada_b200_shard_begin(G->h[d], A[d], B[d], bounds[d], bounds[d + 1], nullptr, &nnz[d], &S[d]);
                if (rcs[d]) errs[d] = g_err;
            });
        for (auto& t : th) t.join();
    }
    for (uint32_t d = 0; d < N; ++d)
        if (rcs[d]) {
            cleanup();
            return fail(rcs[d], "device %u: %s", d, errs[d].c_str());
        }
    // second halves: every shard goes into every device's buffers at its global offset
    std::vector<uint64_t> off(N + 1, 0);
    for (uint32_t d = 0; d < N; ++d) off[d + 1] = off[d] + nnz[d];
    std::vector<spada_b200_stats> stats(N);
    {
        std::vector<std::thread> th;
        for (uint32_t d = 0; d < N; ++d)
            th.emplace_back([&, d]() {
                std::vector<spada_b200_cbuf*> order;
                order.push_back(G->bufs[d]);
                for (uint32_t e = 0; e < N; ++e)
                    if (e != d) order.push_back(G->bufs[e]);
                spada_b200_shard* s = S[d];
                S[d] = nullptr;   // finish releases it
                rcs[d] = spada_b200_shard_finish(s, order.data(), N, off[d], nullptr, d, &stats[d]);
                if (rcs[d]) errs[d] = g_err;
            });
        for (auto& t : th) t.join();
    }
    for (uint32_t d = 0; d < N; ++d)
        if (rcs[d]) {
            cleanup();
            return fail(rcs[d], "device %u: %s", d, errs[d].c_str());
        }
    // the gathered C of device 0 becomes the result (its buffers change owner)
    spada_b200_result* R = new (std::nothrow) spada_b200_result;
    if (!R) { cleanup(); return fail(SPADA_B200_OOM, "host allocation failed"); }
    memset(R, 0, sizeof(*R));
    spada_b200_cbuf* c0 = G->bufs[0];
    R->h = G->h[0];
    R->rows = a->rows;
    R->cols = b->cols;
    R->nnz = off[N];
    R->ptr = c0->ptr;
    R->col = c0->col;
    R->val = c0->val;
    R->stats = stats[0];
    R->stats.rows = a->rows;
    R->stats.nnz_a = a->nnz;
    R->stats.products = products;
    R->stats.nnz_c = off[N];
    float ms = 0.f;
    for (uint32_t d = 0; d < N; ++d) ms = std::max(ms, stats[d].ms_total);
    R->stats.ms_total = ms;
    delete c0;
    for (uint32_t d = 1; d < N; ++d) spada_b200_cbuf_free(G->bufs[d]);
    G->bufs.clear();
    cleanup();
    *out = R;
    return 0;
}
}  // namespace

extern "C" int spada_b200_group_spgemm(spada_b200_group_t* G, const spada_csr_view* a, const spada_csr_view* b,
                                       spada_b200_result_t** out) {
    return group_spgemm(G, a, b, out, spada_b200_upload);
}
extern "C" int spada_b200_group_spgemm32(spada_b200_group_t* G, const spada_csr_view32* a, const spada_csr_view32* b,
                                         spada_b200_result_t** out) {
    return group_spgemm(G, a, b, out, spada_b200_upload32);
}

namespace {
template <typename View, typename UploadFn>
int spgemm_host(spada_b200_t* h, const View* a, const View* b, spada_b200_result_t** out, UploadFn upload) {
    if (!h || !a || !b || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = nullptr;
    if (a->cols != b->rows)
        return fail(SPADA_B200_DIM_MISMATCH, "A is %llu x %llu but B has %llu rows", (unsigned long long)a->rows,
                    (unsigned long long)a->cols, (unsigned long long)b->rows);
    DeviceGuard g(h->device);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
    spada_b200_csr_t *da = nullptr, *db = nullptr;
    int rc = upload(h, a, &da);
    if (rc) {
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return rc;
    }
    // the reference clones A into B for square workloads (gemm.rs:42-43); the same host arrays
    // uploaded once are enough
    bool alias = (const void*)a == (const void*)b ||
                 (a->indptr == b->indptr && a->indices == b->indices && a->data == b->data && a->rows == b->rows &&
                  a->cols == b->cols);
    if (alias) db = da;
    else if ((rc = upload(h, b, &db))) {
        spada_b200_csr_free(da);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return rc;
    }
    cudaEventRecord(e1, h->stream);
    // one-shot operands: re-laying B for a single product costs more than the aligned rows win back
    // (rect: +2.5 ms of build against -0.1 ms of kernel time), the kernels gather through row_ptr
    db->fib_ready = true;
    rc = spada_b200_spgemm_dev(h, da, db, 0, UINT64_MAX, out);
    if (rc == 0) {
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&(*out)->stats.ms_h2d, e0, e1);
        (*out)->stats.nnz_a = a->nnz;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (db != da) spada_b200_csr_free(db);
    spada_b200_csr_free(da);
    return rc;
}
}  // namespace

extern "C" int spada_b200_spgemm(spada_b200_t* h, const spada_csr_view* a, const spada_csr_view* b,
                                 spada_b200_result_t** out) {
    return spgemm_host(h, a, b, out, spada_b200_upload);
}
extern "C" int spada_b200_spgemm32(spada_b200_t* h, const spada_csr_view32* a, const spada_csr_view32* b,
                                   spada_b200_result_t** out) {
    return spgemm_host(h, a, b, out, spada_b200_upload32);
}

// ---- results ------------------------------------------------------------------------------
extern "C" int spada_b200_result_shape(const spada_b200_result_t* r, uint64_t* rows, uint64_t* cols, uint64_t* nnz) {
    if (!r) return fail(SPADA_B200_INVALID_ARG, "result is NULL");
    if (rows) *rows = r->rows;
    if (cols) *cols = r->cols;
    if (nnz) *nnz = r->nnz;
    return 0;
}

extern "C" int spada_b200_result_copy32(const spada_b200_result_t* r, int64_t* indptr, int32_t* indices, double* data) {
    if (!r) return fail(SPADA_B200_INVALID_ARG, "result is NULL");
    spada_b200* h = r->h;
    DeviceGuard g(h->device);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto body = [&]() -> int {
        cudaEventRecord(e0, h->stream);
        if (indptr) CU(cudaMemcpyAsync(indptr, r->ptr, (r->rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        if (indices && r->nnz) CU(cudaMemcpyAsync(indices, r->col, r->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        if (data && r->nnz) CU(cudaMemcpyAsync(data, r->val, r->nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        cudaEventRecord(e1, h->stream);
        CU(cudaStreamSynchronize(h->stream));
        cudaEventElapsedTime(&const_cast<spada_b200_result_t*>(r)->stats.ms_d2h, e0, e1);
        return 0;
    };
    const int rc = body();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

extern "C" int spada_b200_result_copy(const spada_b200_result_t* r, uint64_t* indptr, uint64_t* indices, double* data) {
    if (!r) return fail(SPADA_B200_INVALID_ARG, "result is NULL");
    spada_b200* h = r->h;
    DeviceGuard g(h->device);
    if (indptr) CU(cudaMemcpyAsync(indptr, r->ptr, (r->rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (data && r->nnz) CU(cudaMemcpyAsync(data, r->val, r->nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (indices && r->nnz) {
        const uint64_t chunk = 1ull << 25;  // widen i32 -> usize on the device, 256 MB at a time
        uint64_t* tmp;
        int rc;
        if ((rc = dalloc(h, &tmp, (size_t)std::min<uint64_t>(chunk, r->nnz)))) return rc;
        for (uint64_t o = 0; o < r->nnz; o += chunk) {
            uint64_t n = std::min<uint64_t>(chunk, r->nnz - o);
            launch_narrow_result(nullptr, 0, nullptr, r->col + o, (int64_t)n, tmp, h->stream);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(indices + o, tmp, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
        }
        dfree(h, tmp);
    }
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int spada_b200_result_device_ptrs(const spada_b200_result_t* r, const int64_t** d_indptr,
                                             const int32_t** d_indices, const double** d_data) {
    if (!r) return fail(SPADA_B200_INVALID_ARG, "result is NULL");
    if (d_indptr) *d_indptr = r->ptr;
    if (d_indices) *d_indices = r->col;
    if (d_data) *d_data = r->val;
    return 0;
}

extern "C" int spada_b200_result_stats(const spada_b200_result_t* r, spada_b200_stats* out) {
    if (!r || !out) return fail(SPADA_B200_INVALID_ARG, "NULL argument");
    *out = r->stats;
    return 0;
}

extern "C" void spada_b200_result_free(spada_b200_result_t* r) {
    if (!r) return;
    DeviceGuard g(r->h->device);
    dfree(r->h, r->ptr);
    dfree(r->h, r->col);
    dfree(r->h, r->val);
    delete r;
}
