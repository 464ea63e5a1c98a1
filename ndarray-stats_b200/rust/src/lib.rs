//! UNVERIFIED (never compiled: there is no Rust toolchain in the build image).
//!
//! Drop-in for the bulk quantile path of ndarray-stats: the same sealed extension traits
//! (`src/quantile/mod.rs:12-277, 550-610` of the reference), implemented for the primitive element
//! types the device path covers by calling libndstats_b200.so.  Anything else keeps using the
//! reference crate; there is no CPU fallback inside this crate.
mod ffi;

use ndarray::{Array, ArrayBase, Axis, Data, DataMut, Dimension, Ix1, RemoveAxis};
use ndarray_stats::errors::QuantileError;
use noisy_float::types::{n64, N32, N64};
use std::ptr;

/// `interpolate::{Lower, Higher, Nearest, Midpoint, Linear}` (`src/quantile/interpolate.rs:51-62`) as the
/// enum the C ABI takes.
pub trait DeviceInterpolate {
    const CODE: i32;
}
macro_rules! interp { ($($t:ident = $c:expr),*) => { $(impl DeviceInterpolate for ndarray_stats::interpolate::$t { const CODE: i32 = $c; })* } }
interp!(Lower = 0, Higher = 1, Nearest = 2, Midpoint = 3, Linear = 4);

mod private {
    pub trait Sealed {}
}
/// Element types with a device representation.  N32/N64 have the layout of f32/f64
/// (the reference pointer-casts between them itself, `src/maybe_nan/mod.rs:82-107`).
pub trait DeviceElem: private::Sealed + Copy {
    const DTYPE: i32;
}
macro_rules! elem { ($($t:ty = $c:expr),*) => { $(impl private::Sealed for $t {} impl DeviceElem for $t { const DTYPE: i32 = $c; })* } }
elem!(i8 = 0, i16 = 1, i32 = 2, i64 = 3, u8 = 4, u16 = 5, u32 = 6, u64 = 7, N32 = 8, N64 = 9, f32 = 8, f64 = 9);

thread_local! {
    /// One context per host thread (a ctx is not thread-safe; the reference is stateless).
    static CTX: *mut ffi::nds_ctx = {
        let mut ctx = ptr::null_mut();
        let st = unsafe { ffi::nds_ctx_create(0, &mut ctx) };
        assert_eq!(st, ffi::NDS_OK, "ndstats_b200: no CUDA device (this crate has no CPU path)");
        ctx
    };
}

fn call<A: DeviceElem, S: Data<Elem = A>, D: Dimension>(
    a: &ArrayBase<S, D>, axis: Axis, qs: &[f64], interp: i32, skipnan: bool, out: &mut [A],
) -> Result<(), QuantileError> {
    // ndarray's `len_of` panics on an out-of-bounds axis before anything else (as the reference does)
    let _ = a.len_of(axis);
    let shape: Vec<usize> = a.shape().to_vec();
    let strides: Vec<isize> = a.strides().to_vec();
    let mut bad = 0usize;
    let st = CTX.with(|&ctx| unsafe {
        ffi::nds_quantiles_axis(
            ctx, A::DTYPE, a.as_ptr() as *const _, shape.len() as i32, shape.as_ptr(), strides.as_ptr(),
            axis.index() as i32, qs.as_ptr(), qs.len(), interp, skipnan as i32, out.as_mut_ptr() as *mut _, &mut bad,
        )
    });
    match st {
        ffi::NDS_OK => Ok(()),
        ffi::NDS_EMPTY_INPUT => Err(QuantileError::EmptyInput),
        ffi::NDS_INVALID_QUANTILE => Err(QuantileError::InvalidQuantile(n64(qs[bad]))),
        ffi::NDS_NAN_IN_INPUT => panic!("NaN in a NaN-free float array (noisy_float would have panicked)"),
        ffi::NDS_CAST_OVERFLOW => panic!("called `Option::unwrap()` on a `None` value"), // interpolate.rs:142
        other => panic!("ndstats_b200 status {}", other),
    }
}

/// Same names and signatures as `ndarray_stats::QuantileExt` for the quantile methods
/// (`src/quantile/mod.rs:186-277`); `&mut self` is kept for source compatibility, the data is not reordered.
pub trait QuantileExtB200<A, S, D>: private::Sealed
where
    S: Data<Elem = A>,
    D: Dimension,
{
    fn quantiles_axis_mut<S2, I>(&mut self, axis: Axis, qs: &ArrayBase<S2, Ix1>, interpolate: &I) -> Result<Array<A, D>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, S: DataMut, S2: Data<Elem = N64>, I: DeviceInterpolate;
    fn quantile_axis_mut<I>(&mut self, axis: Axis, q: N64, interpolate: &I) -> Result<Array<A, D::Smaller>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, S: DataMut, I: DeviceInterpolate;
    fn quantile_axis_skipnan_mut<I>(&mut self, axis: Axis, q: N64, interpolate: &I) -> Result<Array<A, D::Smaller>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, S: DataMut, I: DeviceInterpolate;
}

impl<A, S: Data<Elem = A>, D: Dimension> private::Sealed for ArrayBase<S, D> {}

impl<A, S, D> QuantileExtB200<A, S, D> for ArrayBase<S, D>
where
    S: Data<Elem = A>,
    D: Dimension,
{
    fn quantiles_axis_mut<S2, I>(&mut self, axis: Axis, qs: &ArrayBase<S2, Ix1>, _i: &I) -> Result<Array<A, D>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, S: DataMut, S2: Data<Elem = N64>, I: DeviceInterpolate,
    {
        let qv: Vec<f64> = qs.iter().map(|q| q.raw()).collect();
        let mut shape = self.raw_dim();
        shape[axis.index()] = qv.len();                       // mod.rs:451-452
        let mut out = Vec::<A>::with_capacity(shape.size());
        unsafe { out.set_len(shape.size()) };
        call(self, axis, &qv, I::CODE, false, &mut out)?;
        Ok(Array::from_shape_vec(shape, out).unwrap())        // C order, mod.rs:469
    }
    fn quantile_axis_mut<I>(&mut self, axis: Axis, q: N64, i: &I) -> Result<Array<A, D::Smaller>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, S: DataMut, I: DeviceInterpolate,
    {
        self.quantiles_axis_mut(axis, &ndarray::aview1(&[q]), i).map(|a| a.index_axis_move(axis, 0)) // mod.rs:507
    }
    fn quantile_axis_skipnan_mut<I>(&mut self, axis: Axis, q: N64, _i: &I) -> Result<Array<A, D::Smaller>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, S: DataMut, I: DeviceInterpolate,
    {
        let mut shape = self.raw_dim();
        shape[axis.index()] = 1;
        let mut out = Vec::<A>::with_capacity(shape.size());
        unsafe { out.set_len(shape.size()) };
        call(self, axis, &[q.raw()], I::CODE, true, &mut out)?;
        Ok(Array::from_shape_vec(shape, out).unwrap().index_axis_move(axis, 0))
    }
}
