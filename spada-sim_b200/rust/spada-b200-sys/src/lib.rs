//! Raw bindings of `include/spada_b200.h` plus a small safe wrapper (`Engine`) shaped like the
//! call sequence it replaces in spada-sim's `main.rs:74-100`
//! (`Simulator::new` / `execute` / `get_exec_result`).
//!
//! Source only: the repository's build image has no Rust toolchain.  The same symbols are
//! exercised from Python (ctypes) by `tests/test_abi.py` and `tests/test_gpu_parity.py`.
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub const SPADA_B200_OK: c_int = 0;
pub const SPADA_B200_MAX_BINS: usize = 32;
pub const SPADA_B200_MAX_LAUNCHES: usize = 64;
pub const SPADA_B200_FLAG_VALIDATE: u32 = 1;
pub const SPADA_B200_FLAG_TWO_PHASE: u32 = 2;
pub const SPADA_B200_FLAG_SINGLE_PASS: u32 = 4;

/// `Vec<usize>` / `Vec<usize>` / `Vec<f64>` exactly as `CsrMatStorage` holds them (storage.rs:150-160).
#[repr(C)]
pub struct spada_csr_view {
    pub rows: u64,
    pub cols: u64,
    pub nnz: u64,
    pub indptr: *const u64,
    pub indices: *const u64,
    pub data: *const f64,
}

#[repr(C)]
pub struct spada_csr_view32 {
    pub rows: u64,
    pub cols: u64,
    pub nnz: u64,
    pub indptr: *const i32,
    pub indices: *const i32,
    pub data: *const f64,
}

#[repr(C)]
pub struct spada_b200_opts {
    pub device: i32,
    pub accelerator: i32, // 0 Ip, 1 Op, 2 MultiRow, 3 Spada (frontend.rs:33-41)
    pub lane_num: u32,
    pub block_shape: [u32; 2],
    pub flags: u32,
    pub stream: *mut c_void,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct spada_b200_launch {
    pub name: [c_char; 32],
    pub ms: f32,
    pub grid: u32,
    pub rows: u64,
    pub products: u64,
    pub nnz: u64,
}

#[repr(C)]
pub struct spada_b200_stats {
    pub rows: u64,
    pub cols: u64,
    pub nnz_a: u64,
    pub nnz_b: u64,
    pub products: u64,
    pub nnz_c: u64,
    pub bin_rows: [u64; SPADA_B200_MAX_BINS],
    pub bin_products: [u64; SPADA_B200_MAX_BINS],
    pub bin_window_rows: [u32; SPADA_B200_MAX_BINS],
    pub bin_window_lanes: [u32; SPADA_B200_MAX_BINS],
    pub ms_total: f32,
    pub ms_flops: f32,
    pub ms_symbolic: f32,
    pub ms_scan: f32,
    pub ms_numeric: f32,
    pub ms_h2d: f32,
    pub ms_d2h: f32,
    pub n_launches: u32,
    pub n_recorded: u32,
    pub launches: [spada_b200_launch; SPADA_B200_MAX_LAUNCHES],
}

pub enum spada_b200_t {}
pub enum spada_b200_csr_t {}
pub enum spada_b200_result_t {}
pub enum spada_b200_shard_t {}
pub enum spada_b200_cbuf_t {}
pub enum spada_b200_group_t {}
pub const SPADA_B200_IPC_HANDLE_BYTES: usize = 64;

/// Row-panel sink of `spada_b200_spgemm_stream`: the arrays are valid during the call only.
pub type spada_b200_panel_sink = Option<
    unsafe extern "C" fn(
        user: *mut c_void, row_begin: u64, row_end: u64, nnz_begin: u64, indptr: *const i64, indices: *const i32,
        data: *const f64,
    ) -> c_int,
>;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct spada_b200_stream_stats {
    pub panels: u64,
    pub products: u64,
    pub nnz_c: u64,
    pub max_panel_products: u64,
    pub ms_total: f32,
}

extern "C" {
    pub fn spada_b200_abi_version() -> c_int;
    pub fn spada_b200_last_error() -> *const c_char;
    pub fn spada_b200_device_count(count: *mut c_int) -> c_int;
    pub fn spada_b200_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn spada_b200_host_free(ptr: *mut c_void) -> c_int;
    pub fn spada_b200_create(opts: *const spada_b200_opts, out: *mut *mut spada_b200_t) -> c_int;
    pub fn spada_b200_destroy(h: *mut spada_b200_t);
    pub fn spada_b200_set_stream(h: *mut spada_b200_t, cuda_stream: *mut c_void) -> c_int;
    pub fn spada_b200_synchronize(h: *mut spada_b200_t) -> c_int;
    pub fn spada_b200_trim(h: *mut spada_b200_t) -> c_int;
    pub fn spada_b200_upload(h: *mut spada_b200_t, m: *const spada_csr_view, out: *mut *mut spada_b200_csr_t) -> c_int;
    pub fn spada_b200_upload32(h: *mut spada_b200_t, m: *const spada_csr_view32, out: *mut *mut spada_b200_csr_t) -> c_int;
    pub fn spada_b200_csr_wrap_device(
        h: *mut spada_b200_t, rows: u64, cols: u64, nnz: u64, d_indptr: *const i64, d_indices: *const i32,
        d_data: *const f64, out: *mut *mut spada_b200_csr_t,
    ) -> c_int;
    pub fn spada_b200_csr_prepare(h: *mut spada_b200_t, m: *mut spada_b200_csr_t, ms_or_null: *mut f32) -> c_int;
    pub fn spada_b200_csr_set_one_shot(m: *mut spada_b200_csr_t) -> c_int;
    pub fn spada_b200_transpose(h: *mut spada_b200_t, a: *const spada_b200_csr_t, out: *mut *mut spada_b200_csr_t) -> c_int;
    pub fn spada_b200_csr_shape(m: *const spada_b200_csr_t, rows: *mut u64, cols: *mut u64, nnz: *mut u64) -> c_int;
    pub fn spada_b200_csr_device_ptrs(
        m: *const spada_b200_csr_t, d_indptr: *mut *const i64, d_indices: *mut *const i32, d_data: *mut *const f64,
    ) -> c_int;
    pub fn spada_b200_csr_download32(m: *const spada_b200_csr_t, indptr: *mut i64, indices: *mut i32, data: *mut f64) -> c_int;
    pub fn spada_b200_csr_free(m: *mut spada_b200_csr_t);
    pub fn spada_b200_spgemm_dev(
        h: *mut spada_b200_t, a: *const spada_b200_csr_t, b: *const spada_b200_csr_t, row_begin: u64, row_end: u64,
        out: *mut *mut spada_b200_result_t,
    ) -> c_int;
    pub fn spada_b200_spgemm(
        h: *mut spada_b200_t, a: *const spada_csr_view, b: *const spada_csr_view, out: *mut *mut spada_b200_result_t,
    ) -> c_int;
    pub fn spada_b200_spgemm32(
        h: *mut spada_b200_t, a: *const spada_csr_view32, b: *const spada_csr_view32,
        out: *mut *mut spada_b200_result_t,
    ) -> c_int;
    pub fn spada_b200_flops(
        h: *mut spada_b200_t, a: *const spada_b200_csr_t, b: *const spada_b200_csr_t, total_products: *mut u64,
        host_flops_or_null: *mut u64,
    ) -> c_int;
    pub fn spada_b200_plan_shards(
        h: *mut spada_b200_t, a: *const spada_b200_csr_t, b: *const spada_b200_csr_t, n_shards: u32, bounds: *mut u64,
    ) -> c_int;
    // sharded runs: A row-sharded over several GPUs, B replicated, C gathered on every GPU by the placement kernel
    pub fn spada_b200_cbuf_create(
        h: *mut spada_b200_t, rows: u64, cols: u64, capacity_nnz: u64, out: *mut *mut spada_b200_cbuf_t,
    ) -> c_int;
    pub fn spada_b200_cbuf_export(c: *const spada_b200_cbuf_t, handles: *mut c_void) -> c_int;
    pub fn spada_b200_cbuf_import(
        h: *mut spada_b200_t, handles: *const c_void, rows: u64, cols: u64, capacity_nnz: u64,
        out: *mut *mut spada_b200_cbuf_t,
    ) -> c_int;
    pub fn spada_b200_cbuf_free(c: *mut spada_b200_cbuf_t);
    pub fn spada_b200_cbuf_device_ptrs(
        c: *const spada_b200_cbuf_t, d_indptr: *mut *const i64, d_indices: *mut *const i32, d_data: *mut *const f64,
    ) -> c_int;
    pub fn spada_b200_cbuf_nnz(c: *const spada_b200_cbuf_t, nnz: *mut u64) -> c_int;
    pub fn spada_b200_cbuf_copy32(c: *const spada_b200_cbuf_t, indptr: *mut i64, indices: *mut i32, data: *mut f64) -> c_int;
    pub fn spada_b200_shard_begin(
        h: *mut spada_b200_t, a: *const spada_b200_csr_t, b: *const spada_b200_csr_t, row_begin: u64, row_end: u64,
        d_nnz_local: *mut i64, nnz_local: *mut u64, out: *mut *mut spada_b200_shard_t,
    ) -> c_int;
    pub fn spada_b200_shard_finish(
        s: *mut spada_b200_shard_t, bufs: *const *mut spada_b200_cbuf_t, n_bufs: u32, nnz_offset: u64,
        d_shard_nnz: *const i64, shard_index: u32, stats_or_null: *mut spada_b200_stats,
    ) -> c_int;
    pub fn spada_b200_shard_abort(s: *mut spada_b200_shard_t);
    // all GPUs of this process behind one call (what main.rs drives when n_gpus > 1)
    pub fn spada_b200_group_create(opts: *const spada_b200_opts, n_gpus: u32, out: *mut *mut spada_b200_group_t) -> c_int;
    pub fn spada_b200_group_spgemm(
        g: *mut spada_b200_group_t, a: *const spada_csr_view, b: *const spada_csr_view, out: *mut *mut spada_b200_result_t,
    ) -> c_int;
    pub fn spada_b200_group_spgemm32(
        g: *mut spada_b200_group_t, a: *const spada_csr_view32, b: *const spada_csr_view32,
        out: *mut *mut spada_b200_result_t,
    ) -> c_int;
    pub fn spada_b200_group_destroy(g: *mut spada_b200_group_t);
    // row panels: C larger than HBM, D2H of a panel beside the next panel's kernels
    pub fn spada_b200_spgemm_stream(
        h: *mut spada_b200_t, a: *const spada_b200_csr_t, b: *const spada_b200_csr_t, panel_products: u64,
        sink: spada_b200_panel_sink, user: *mut c_void, stats_or_null: *mut spada_b200_stream_stats,
    ) -> c_int;
    pub fn spada_b200_spgemm_to_host(
        h: *mut spada_b200_t, a: *const spada_b200_csr_t, b: *const spada_b200_csr_t, panel_products: u64,
        indptr: *mut i64, indices: *mut i32, data: *mut f64, capacity_nnz: u64,
        stats_or_null: *mut spada_b200_stream_stats,
    ) -> c_int;
    pub fn spada_b200_spgemm32_host_to_host(
        h: *mut spada_b200_t, a: *const spada_csr_view32, b: *const spada_csr_view32, indptr: *mut i64, indices: *mut i32,
        data: *mut f64, capacity_nnz: u64, stats_or_null: *mut spada_b200_stream_stats,
    ) -> c_int;
    pub fn spada_b200_result_shape(r: *const spada_b200_result_t, rows: *mut u64, cols: *mut u64, nnz: *mut u64) -> c_int;
    pub fn spada_b200_result_copy(r: *const spada_b200_result_t, indptr: *mut u64, indices: *mut u64, data: *mut f64) -> c_int;
    pub fn spada_b200_result_copy32(r: *const spada_b200_result_t, indptr: *mut i64, indices: *mut i32, data: *mut f64) -> c_int;
    pub fn spada_b200_result_device_ptrs(
        r: *const spada_b200_result_t, d_indptr: *mut *const i64, d_indices: *mut *const i32, d_data: *mut *const f64,
    ) -> c_int;
    pub fn spada_b200_result_stats(r: *const spada_b200_result_t, out: *mut spada_b200_stats) -> c_int;
    pub fn spada_b200_result_free(r: *mut spada_b200_result_t);
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(spada_b200_last_error()).to_string_lossy().into_owned() }
}

/// The reference panics on every failure (`unwrap()` / `panic!`, main.rs:32-39); so does this wrapper.
fn check(rc: c_int) {
    if rc != SPADA_B200_OK {
        panic!("spada_b200 error {}: {}", rc, last_error());
    }
}

/// One output row: what `CsrRow::new_from_data(rowptr, data, indptr)` takes (storage.rs:52-59).
pub struct RowOut {
    pub rowptr: usize,
    pub data: Vec<f64>,
    pub indptr: Vec<usize>, // column ids (the reference's naming)
}

/// Drop-in for the `Simulator::new` / `execute` / `get_exec_result` triple of main.rs:74-100.
pub struct Engine {
    h: *mut spada_b200_t,
    result: *mut spada_b200_result_t,
}

impl Engine {
    /// `accelerator`: 0 Ip, 1 Op, 2 MultiRow, 3 Spada; `block_shape`/`lane_num` from `OmegaConfig`.
    pub fn new(accelerator: i32, lane_num: usize, block_shape: [usize; 2]) -> Engine {
        let opts = spada_b200_opts {
            device: -1,
            accelerator,
            lane_num: lane_num as u32,
            block_shape: [block_shape[0].min(u32::MAX as usize) as u32, block_shape[1].min(u32::MAX as usize) as u32],
            flags: SPADA_B200_FLAG_VALIDATE,
            stream: std::ptr::null_mut(),
        };
        let mut h = std::ptr::null_mut();
        check(unsafe { spada_b200_create(&opts, &mut h) });
        Engine { h, result: std::ptr::null_mut() }
    }

    /// `execute`: C = A x B over the borrowed `CsrMatStorage` buffers (zero-copy views; usize == u64).
    /// `a_shape`/`b_shape` are (rows, cols).
    pub fn execute(
        &mut self, a_shape: (usize, usize), a_indptr: &[usize], a_indices: &[usize], a_data: &[f64],
        b_shape: (usize, usize), b_indptr: &[usize], b_indices: &[usize], b_data: &[f64],
    ) {
        let va = spada_csr_view {
            rows: a_shape.0 as u64, cols: a_shape.1 as u64, nnz: a_data.len() as u64,
            indptr: a_indptr.as_ptr() as *const u64, indices: a_indices.as_ptr() as *const u64, data: a_data.as_ptr(),
        };
        let vb = spada_csr_view {
            rows: b_shape.0 as u64, cols: b_shape.1 as u64, nnz: b_data.len() as u64,
            indptr: b_indptr.as_ptr() as *const u64, indices: b_indices.as_ptr() as *const u64, data: b_data.as_ptr(),
        };
        if !self.result.is_null() {
            unsafe { spada_b200_result_free(self.result) };
            self.result = std::ptr::null_mut();
        }
        check(unsafe { spada_b200_spgemm(self.h, &va, &vb, &mut self.result) });
    }

    /// `get_exec_result`: one row per A row in raw row order, empty rows kept (simulator.rs:1034-1062).
    pub fn get_exec_result(&self) -> Vec<RowOut> {
        let (mut rows, mut cols, mut nnz) = (0u64, 0u64, 0u64);
        check(unsafe { spada_b200_result_shape(self.result, &mut rows, &mut cols, &mut nnz) });
        let mut indptr = vec![0u64; rows as usize + 1];
        let mut indices = vec![0u64; nnz as usize];
        let mut data = vec![0f64; nnz as usize];
        check(unsafe { spada_b200_result_copy(self.result, indptr.as_mut_ptr(), indices.as_mut_ptr(), data.as_mut_ptr()) });
        (0..rows as usize)
            .map(|r| {
                let (s, e) = (indptr[r] as usize, indptr[r + 1] as usize);
                RowOut { rowptr: r, data: data[s..e].to_vec(), indptr: indices[s..e].iter().map(|&c| c as usize).collect() }
            })
            .collect()
    }

    pub fn stats(&self) -> Box<spada_b200_stats> {
        let mut st: Box<spada_b200_stats> = unsafe { Box::new(std::mem::zeroed()) };
        check(unsafe { spada_b200_result_stats(self.result, &mut *st) });
        st
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe {
            if !self.result.is_null() {
                spada_b200_result_free(self.result);
            }
            spada_b200_destroy(self.h);
        }
    }
}
