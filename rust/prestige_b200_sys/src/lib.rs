//! Raw bindings of include/prestige_b200.h.  One-to-one with the C declarations; no logic.
//! UNCOMPILED in this repository (no Rust toolchain in the image) -- kept in sync by hand with the header.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct pst_ctx {
    _private: [u8; 0],
}

pub type pst_status = c_int;
pub const PST_OK: pst_status = 0;
pub const PST_EINVAL: pst_status = 1;
pub const PST_ENOMEM: pst_status = 2;
pub const PST_ECUDA: pst_status = 3;
pub const PST_ENCCL: pst_status = 4;
pub const PST_EOVERFLOW: pst_status = 5;
pub const PST_ESTATE: pst_status = 6;

pub const PST_F32: c_int = 0;
pub const PST_F64: c_int = 1;
pub const PST_U32: c_int = 2;
pub const PST_I32: c_int = 3;
pub const PST_REAL: c_int = 15;
pub const PST_KEY_LINEAR: c_int = 0;
pub const PST_KEY_MORTON: c_int = 1;
pub const PST_PHYS_WCSPH: u32 = 1;
pub const PST_PHYS_DEM: u32 = 2;
pub const PST_ARRAY_PERSISTENT: u32 = 1;
pub const PST_ARRAY_OUTPUT: u32 = 2;
pub const PST_COMM_ID_BYTES: usize = 128;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pst_config {
    pub struct_size: u32,
    pub device: i32,
    pub dim: i32,
    pub real: i32,
    pub key: i32,
    pub max_contacts: i32,
    pub physics: u32,
    pub reserved: u32,
    pub capacity: u64,
    pub ghost_capacity: u64,
    pub lo: [c_double; 3],
    pub hi: [c_double; 3],
    pub cell_size: c_double,
}

unsafe extern "C" {
    pub fn pst_version() -> *const c_char;
    pub fn pst_create(cfg: *const pst_config, out: *mut *mut pst_ctx) -> pst_status;
    pub fn pst_destroy(ctx: *mut pst_ctx);
    pub fn pst_last_error(ctx: *const pst_ctx) -> *const c_char;
    pub fn pst_stream(ctx: *mut pst_ctx) -> *mut c_void;
    pub fn pst_sync(ctx: *mut pst_ctx) -> pst_status;
    pub fn pst_set_param(ctx: *mut pst_ctx, name: *const c_char, value: c_double) -> pst_status;
    pub fn pst_get_param(ctx: *mut pst_ctx, name: *const c_char, value: *mut c_double) -> pst_status;
    pub fn pst_set_count(ctx: *mut pst_ctx, n: u64) -> pst_status;
    pub fn pst_get_count(ctx: *mut pst_ctx, n_owned: *mut u64, n_ghost: *mut u64) -> pst_status;
    pub fn pst_array_create(ctx: *mut pst_ctx, name: *const c_char, dtype: c_int, flags: u32) -> pst_status;
    pub fn pst_array(ctx: *mut pst_ctx, name: *const c_char, dev_ptr: *mut *mut c_void, n: *mut usize, dtype: *mut c_int, rows: *mut c_int) -> pst_status;
    pub fn pst_upload(ctx: *mut pst_ctx, name: *const c_char, host: *const c_void, n: usize) -> pst_status;
    pub fn pst_download(ctx: *mut pst_ctx, name: *const c_char, host: *mut c_void, n: usize) -> pst_status;
    pub fn pst_upload_async(ctx: *mut pst_ctx, name: *const c_char, host: *const c_void, n: usize) -> pst_status;
    pub fn pst_download_async(ctx: *mut pst_ctx, name: *const c_char, host: *mut c_void, n: usize) -> pst_status;
    pub fn pst_wait_transfers(ctx: *mut pst_ctx) -> pst_status;
    pub fn pst_host_alloc(bytes: usize) -> *mut c_void;
    pub fn pst_host_free(p: *mut c_void);
    pub fn pst_build_neighbours(ctx: *mut pst_ctx) -> pst_status;
    pub fn pst_apply(ctx: *mut pst_ctx, eq_names: *const *const c_char, n_eq: c_int) -> pst_status;
    pub fn pst_dump_pairs(ctx: *mut pst_ctx, mode: c_int, i: *mut u32, j: *mut u32, cap: usize, n_pairs: *mut usize) -> pst_status;
    pub fn pst_step(ctx: *mut pst_ctx, dt: c_double, n_steps: c_int) -> pst_status;
    pub fn pst_integrate(ctx: *mut pst_ctx, dt: c_double) -> pst_status;
    pub fn pst_get_stat(ctx: *mut pst_ctx, name: *const c_char, value: *mut c_double) -> pst_status;
    pub fn pst_kernel_name(ctx: *mut pst_ctx, stage: *const c_char, buf: *mut c_char, cap: usize) -> pst_status;
    pub fn pst_set_option(ctx: *mut pst_ctx, name: *const c_char, value: c_int) -> pst_status;
    pub fn pst_comm_unique_id(id_bytes: *mut c_void) -> pst_status;
    pub fn pst_comm_init(ctx: *mut pst_ctx, id_bytes: *const c_void, rank: c_int, n_ranks: c_int) -> pst_status;
    pub fn pst_halo_exchange(ctx: *mut pst_ctx) -> pst_status;
    pub fn pst_bodies_create(ctx: *mut pst_ctx, n_bodies: u32) -> pst_status;
    pub fn pst_bodies_setup(ctx: *mut pst_ctx) -> pst_status;
    pub fn pst_bodies_restore(ctx: *mut pst_ctx) -> pst_status;
    pub fn pst_bodies_state(ctx: *mut pst_ctx, name: *const c_char, host: *mut c_double, n: usize, write: c_int) -> pst_status;
}
