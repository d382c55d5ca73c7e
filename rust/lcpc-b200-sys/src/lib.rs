//! Raw bindings to include/lcpc_b200.h (ABI version 1).  One declaration per exported
//! function used by the safe wrapper; see the header for the contract of each.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct lcpc_ctx { _p: [u8; 0] }
#[repr(C)] pub struct lcpc_plan { _p: [u8; 0] }
#[repr(C)] pub struct lcpc_stream { _p: [u8; 0] }
#[repr(C)] pub struct lcpc_commit { _p: [u8; 0] }
#[repr(C)] pub struct lcpc_transcript { _p: [u8; 0] }

/// sprs::CsMat<F> in CSC storage (indptr / indices / data), host pointers.
#[repr(C)]
pub struct lcpc_csc {
    pub rows: u64,
    pub cols: u64,
    pub indptr: *const u64,
    pub indices: *const u64,
    pub data: *const u64,
}

pub const LCPC_FT63: i32 = 0;
pub const LCPC_FT127: i32 = 1;
pub const LCPC_FT191: i32 = 2;
pub const LCPC_FT255: i32 = 3;
/// proof-of-storage `Ft253_192` (big-endian `to_repr`)
pub const LCPC_FT253_192: i32 = 4;

pub const LCPC_OK: i32 = 0;
pub const LCPC_ERR_TOO_BIG: i32 = -1;
pub const LCPC_ERR_ENCODE: i32 = -2;
pub const LCPC_ERR_COMMIT: i32 = -3;
pub const LCPC_ERR_COLUMN_NUMBER: i32 = -4;
pub const LCPC_ERR_OUTER_TENSOR: i32 = -5;

extern "C" {
    pub fn lcpc_abi_version() -> u32;
    pub fn lcpc_last_error() -> *const c_char;
    pub fn lcpc_ctx_create(device: i32, out: *mut *mut lcpc_ctx) -> i32;
    pub fn lcpc_ctx_destroy(ctx: *mut lcpc_ctx);
    pub fn lcpc_plan_ligero(ctx: *mut lcpc_ctx, field: i32, n_per_row: usize, n_cols: usize,
                            root_of_unity_mont: *const u64, out: *mut *mut lcpc_plan) -> i32;
    pub fn lcpc_plan_brakedown(ctx: *mut lcpc_ctx, field: i32, n_per_row: usize, n_cols: usize, n_levels: usize,
                               precodes: *const lcpc_csc, postcodes: *const lcpc_csc, out: *mut *mut lcpc_plan) -> i32;
    pub fn lcpc_plan_destroy(plan: *mut lcpc_plan);
    pub fn lcpc_encode_rows(plan: *mut lcpc_plan, rows: *mut u64, n_rows: usize) -> i32;
    pub fn lcpc_commit_host(plan: *mut lcpc_plan, coeffs: *const u64, n_coeffs: usize, coeffs_out: *mut u64,
                            comm_out: *mut u64, hashes_out: *mut u8, keep: *mut *mut lcpc_commit) -> i32;
    pub fn lcpc_commit_bytes_host(plan: *mut lcpc_plan, file_bytes: *const u8, n_bytes: usize, coeffs_out: *mut u64,
                                  comm_out: *mut u64, hashes_out: *mut u8, keep: *mut *mut lcpc_commit) -> i32;
    pub fn lcpc_commit_root(c: *mut lcpc_commit, root_out: *mut u8) -> i32;
    pub fn lcpc_commit_free(c: *mut lcpc_commit);
    pub fn lcpc_fold_host(c: *mut lcpc_commit, which: i32, tensors: *const u64, n_tensors: usize, out: *mut u64) -> i32;
    pub fn lcpc_open_columns_host(c: *mut lcpc_commit, cols: *const u64, n: usize, cols_out: *mut u64,
                                  paths_out: *mut u8) -> i32;
    pub fn lcpc_leaves_host(c: *mut lcpc_commit, cols: *const u64, n: usize, leaves_out: *mut u8) -> i32;
    // streaming commit (EncodedFileWriter + ColumnDigestAccumulator) and edits on a resident commitment
    pub fn lcpc_stream_begin(plan: *mut lcpc_plan, max_rows: usize, block_rows: usize, sink: *mut u8,
                             sink_row_capacity: usize, out: *mut *mut lcpc_stream) -> i32;
    pub fn lcpc_stream_push_elems_host(s: *mut lcpc_stream, elems: *const u64, n_elems: usize) -> i32;
    pub fn lcpc_stream_push_bytes_host(s: *mut lcpc_stream, bytes: *const u8, n_bytes: usize) -> i32;
    pub fn lcpc_stream_finish(s: *mut lcpc_stream, hashes_out: *mut u8, n_rows_out: *mut usize) -> i32;
    pub fn lcpc_stream_free(s: *mut lcpc_stream);
    pub fn lcpc_commit_update_rows_host(c: *mut lcpc_commit, row0: usize, n_rows: usize, coeff_rows: *const u64,
                                        comm_rows_out: *mut u64, hashes_out: *mut u8) -> i32;
    pub fn lcpc_commit_append_rows_host(c: *mut lcpc_commit, row0: usize, n_rows: usize, coeff_rows: *const u64,
                                        comm_rows_out: *mut u64, hashes_out: *mut u8) -> i32;
}

#[allow(dead_code)]
fn _unused(_: *mut c_void) {}
