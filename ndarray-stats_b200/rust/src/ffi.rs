//! UNVERIFIED (never compiled). `extern "C"` declarations of include/ndstats_b200.h, one per symbol used.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct nds_ctx {
    _private: [u8; 0],
}

pub const NDS_OK: c_int = 0;
pub const NDS_EMPTY_INPUT: c_int = 1;
pub const NDS_INVALID_QUANTILE: c_int = 2;
pub const NDS_NAN_IN_INPUT: c_int = 3;
pub const NDS_CAST_OVERFLOW: c_int = 8;
pub const NDS_UNDEFINED_ORDER: c_int = 10;

extern "C" {
    pub fn nds_ctx_create(device: c_int, out: *mut *mut nds_ctx) -> c_int;
    pub fn nds_ctx_destroy(ctx: *mut nds_ctx);
    pub fn nds_ctx_last_error(ctx: *mut nds_ctx) -> *const c_char;
    /// QuantileExt::quantiles_axis_mut / quantile_axis_mut / quantile_axis_skipnan_mut, Quantile1dExt::*
    pub fn nds_quantiles_axis(
        ctx: *mut nds_ctx, dtype: c_int, data: *const c_void, ndim: c_int, shape: *const usize,
        strides_elems: *const isize, axis: c_int, qs: *const f64, nq: usize, interp: c_int, skipnan: c_int,
        out: *mut c_void, bad_q_index: *mut usize,
    ) -> c_int;
    /// QuantileExt::{argmin, argmax, min, max}(+_skipnan)
    pub fn nds_minmax(
        ctx: *mut nds_ctx, dtype: c_int, data: *const c_void, ndim: c_int, shape: *const usize,
        strides_elems: *const isize, want_max: c_int, skipnan: c_int, out_value: *mut c_void, out_index: *mut usize,
    ) -> c_int;
    /// Sort1dExt::get_many_from_sorted_mut / get_from_sorted_mut
    pub fn nds_get_many_from_sorted(
        ctx: *mut nds_ctx, dtype: c_int, data: *const c_void, n: usize, stride_elems: isize,
        indexes: *const usize, k: usize, out_values: *mut c_void,
    ) -> c_int;
    /// Sort1dExt::partition_mut (in place)
    pub fn nds_partition(
        ctx: *mut nds_ctx, dtype: c_int, data: *mut c_void, n: usize, stride_elems: isize,
        pivot_index: usize, partition_index: *mut usize,
    ) -> c_int;
}
