//! Raw bindings of include/cgvec.h (hand-written; the header is the source of truth).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct cgvec_index {
    _private: [u8; 0],
}

pub const CGVEC_F32: c_int = 0;
pub const CGVEC_F16: c_int = 1;
pub const CGVEC_COSINE: c_int = 0;
pub const CGVEC_DOT: c_int = 1;
pub const CGVEC_L2: c_int = 2;
pub const CGVEC_FORMULA_SIMD: c_int = 0;
pub const CGVEC_FORMULA_SEQ: c_int = 2;
pub const CGVEC_OK: c_int = 0;
pub const CGVEC_ERR_NOT_FOUND: c_int = -6;
pub const CGVEC_ERR_DISABLED: c_int = -9;

#[repr(C)]
pub struct cgvec_search_opts {
    pub struct_size: u32,
    pub metric: c_int,
    pub formula: c_int,
    pub path: c_int,
    pub stream: *mut c_void,
    pub device_io: c_int,
}

#[repr(C)]
pub struct cgvec_server {
    _private: [u8; 0],
}

extern "C" {
    // resident batch-1 sessions (include/cgvec.h, "resident batch-1 sessions"): the scan kernel stays on the GPU between calls
    pub fn cgvec_serve_open(idx: *mut cgvec_index, k: u32, metric: c_int, out: *mut *mut cgvec_server) -> c_int;
    pub fn cgvec_serve_search(
        s: *mut cgvec_server,
        query: *const f32,
        out_rows: *mut u64,
        out_ids: *mut [u8; 16],
        out_scores: *mut f32,
        out_count: *mut u32,
    ) -> c_int;
    pub fn cgvec_serve_pause(s: *mut cgvec_server) -> c_int;
    pub fn cgvec_serve_close(s: *mut cgvec_server) -> c_int;
}

extern "C" {
    pub fn cgvec_create(dim: u32, storage: c_int, device_ids: *const c_int, n_devices: c_int, out: *mut *mut cgvec_index) -> c_int;
    pub fn cgvec_create_from_env(dim: u32, storage: c_int, enable_gpu: c_int, out: *mut *mut cgvec_index) -> c_int;
    pub fn cgvec_destroy(idx: *mut cgvec_index) -> c_int;
    pub fn cgvec_reserve(idx: *mut cgvec_index, n_rows: u64) -> c_int;
    pub fn cgvec_add(idx: *mut cgvec_index, ids: *const [u8; 16], rows_f32: *const f32, n: u64) -> c_int;
    pub fn cgvec_len(idx: *const cgvec_index) -> u64;
    pub fn cgvec_search(
        idx: *const cgvec_index, queries: *const f32, nq: u32, k: u32, metric: c_int,
        out_rows: *mut u64, out_ids: *mut [u8; 16], out_scores: *mut f32, out_counts: *mut u32,
    ) -> c_int;
    pub fn cgvec_get(idx: *const cgvec_index, id: *const u8, out_row: *mut f32) -> c_int;
    pub fn cgvec_row_of_id(idx: *const cgvec_index, id: *const u8, out_local_row: *mut u64) -> c_int;
    pub fn cgvec_rescore(
        idx: *const cgvec_index, query: *const f32, local_rows: *const u64, n: u32, metric: c_int, formula: c_int, out_scores: *mut f32,
    ) -> c_int;
    pub fn cgvec_search_ex(
        idx: *const cgvec_index, queries: *const f32, nq: u32, k: u32, opts: *const cgvec_search_opts,
        out_rows: *mut u64, out_ids: *mut [u8; 16], out_scores: *mut f32, out_counts: *mut u32,
    ) -> c_int;
    pub fn cgvec_normalize_rows(idx: *mut cgvec_index) -> c_int;
    pub fn cgvec_quantize_i8(idx: *mut cgvec_index) -> c_int;
    pub fn cgvec_search_i8(
        idx: *const cgvec_index, query: *const f32, limit: u32, out_rows: *mut u64, out_scores: *mut f32, out_count: *mut u32,
    ) -> c_int;
    pub fn cgvec_save_flat(idx: *const cgvec_index, path: *const c_char) -> c_int;
    pub fn cgvec_load_flat(idx: *mut cgvec_index, path: *const c_char, out_rows_loaded: *mut u64) -> c_int;
    pub fn cgvec_distances_first(idx: *const cgvec_index, query: *const f32, limit: u64, out: *mut f32, out_n: *mut u64) -> c_int;
    pub fn cgvec_last_error() -> *const c_char;
}
