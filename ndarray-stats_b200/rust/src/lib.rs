//! UNVERIFIED (never compiled: there is no Rust toolchain in the build image).
//!
//! Drop-in for the bulk quantile path of ndarray-stats: the same sealed extension traits
//! (`src/quantile/mod.rs:12-277, 550-610` of the reference), implemented for the primitive element
//! types the device path covers by calling libndstats_b200.so.  Anything else keeps using the
//! reference crate; there is no CPU fallback inside this crate.
mod ffi;

use indexmap::IndexMap;
use ndarray::{Array, Array1, ArrayRef, Axis, Dimension, Ix1, RemoveAxis};
use ndarray_stats::errors::{MinMaxError, QuantileError};
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
    pub trait SealedArray {}
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

fn call<A: DeviceElem, D: Dimension>(
    a: &ArrayRef<A, D>, axis: Axis, qs: &[f64], interp: i32, skipnan: bool, out: &mut [A],
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

/// Same names and signatures as the quantile methods of `ndarray_stats::QuantileExt<A, D>`, which v0.7 implements for
/// `ArrayRef<A, D>` and whose `qs` are `&ArrayRef<N64, Ix1>` (`src/quantile/mod.rs:12-15, 186-277, 279`).  `&mut self`
/// is kept for source compatibility; the data is not reordered.
pub trait QuantileExtB200<A, D>: private::SealedArray
where
    D: Dimension,
{
    fn quantiles_axis_mut<I>(&mut self, axis: Axis, qs: &ArrayRef<N64, Ix1>, interpolate: &I) -> Result<Array<A, D>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, I: DeviceInterpolate;
    fn quantile_axis_mut<I>(&mut self, axis: Axis, q: N64, interpolate: &I) -> Result<Array<A, D::Smaller>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, I: DeviceInterpolate;
    fn quantile_axis_skipnan_mut<I>(&mut self, axis: Axis, q: N64, interpolate: &I) -> Result<Array<A, D::Smaller>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, I: DeviceInterpolate;
    /// `src/quantile/mod.rs:283-409`: the eight min / max methods are one device reduction each.
    fn argmin(&self) -> Result<D::Pattern, MinMaxError> where A: DeviceElem;
    fn argmax(&self) -> Result<D::Pattern, MinMaxError> where A: DeviceElem;
    fn min(&self) -> Result<A, MinMaxError> where A: DeviceElem;
    fn max(&self) -> Result<A, MinMaxError> where A: DeviceElem;
}

impl<A, D: Dimension> private::SealedArray for ArrayRef<A, D> {}

fn minmax<A: DeviceElem, D: Dimension>(a: &ArrayRef<A, D>, want_max: bool, skipnan: bool) -> Result<(A, Vec<usize>), MinMaxError> {
    let shape: Vec<usize> = a.shape().to_vec();
    let strides: Vec<isize> = a.strides().to_vec();
    let mut value = std::mem::MaybeUninit::<A>::uninit();
    let mut index = vec![0usize; shape.len().max(1)];
    let st = CTX.with(|&ctx| unsafe {
        ffi::nds_minmax(ctx, A::DTYPE, a.as_ptr() as *const _, shape.len() as i32, shape.as_ptr(), strides.as_ptr(),
                        want_max as i32, skipnan as i32, value.as_mut_ptr() as *mut _, index.as_mut_ptr())
    });
    match st {
        ffi::NDS_OK => Ok((unsafe { value.assume_init() }, index)),
        ffi::NDS_EMPTY_INPUT => Err(MinMaxError::EmptyInput),
        ffi::NDS_UNDEFINED_ORDER => Err(MinMaxError::UndefinedOrder),
        other => panic!("ndstats_b200 status {}", other),
    }
}

impl<A, D> QuantileExtB200<A, D> for ArrayRef<A, D>
where
    D: Dimension,
{
    fn quantiles_axis_mut<I>(&mut self, axis: Axis, qs: &ArrayRef<N64, Ix1>, _i: &I) -> Result<Array<A, D>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, I: DeviceInterpolate,
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
        D: RemoveAxis, A: DeviceElem, I: DeviceInterpolate,
    {
        self.quantiles_axis_mut(axis, &ndarray::aview1(&[q]), i).map(|a| a.index_axis_move(axis, 0)) // mod.rs:507
    }
    fn quantile_axis_skipnan_mut<I>(&mut self, axis: Axis, q: N64, _i: &I) -> Result<Array<A, D::Smaller>, QuantileError>
    where
        D: RemoveAxis, A: DeviceElem, I: DeviceInterpolate,
    {
        let mut shape = self.raw_dim();
        shape[axis.index()] = 1;
        let mut out = Vec::<A>::with_capacity(shape.size());
        unsafe { out.set_len(shape.size()) };
        call(self, axis, &[q.raw()], I::CODE, true, &mut out)?;
        Ok(Array::from_shape_vec(shape, out).unwrap().index_axis_move(axis, 0))
    }
    fn argmin(&self) -> Result<D::Pattern, MinMaxError> where A: DeviceElem {
        let (_, idx) = minmax(self, false, false)?;
        let mut d = D::zeros(idx.len());                      // the index as the dimension's pattern (mod.rs:283-300)
        d.slice_mut().copy_from_slice(&idx);
        Ok(d.into_pattern())
    }
    fn argmax(&self) -> Result<D::Pattern, MinMaxError> where A: DeviceElem {
        let (_, idx) = minmax(self, true, false)?;
        let mut d = D::zeros(idx.len());
        d.slice_mut().copy_from_slice(&idx);
        Ok(d.into_pattern())
    }
    fn min(&self) -> Result<A, MinMaxError> where A: DeviceElem { minmax(self, false, false).map(|(v, _)| v) }
    fn max(&self) -> Result<A, MinMaxError> where A: DeviceElem { minmax(self, true, false).map(|(v, _)| v) }
}

/// `Quantile1dExt<A>` for `ArrayRef<A, Ix1>` (`src/quantile/mod.rs:550-633`).
pub trait Quantile1dExtB200<A>: private::SealedArray {
    fn quantile_mut<I>(&mut self, q: N64, interpolate: &I) -> Result<A, QuantileError> where A: DeviceElem, I: DeviceInterpolate;
    fn quantiles_mut<I>(&mut self, qs: &ArrayRef<N64, Ix1>, interpolate: &I) -> Result<Array1<A>, QuantileError>
    where A: DeviceElem, I: DeviceInterpolate;
}
impl<A> Quantile1dExtB200<A> for ArrayRef<A, Ix1> {
    fn quantile_mut<I>(&mut self, q: N64, i: &I) -> Result<A, QuantileError> where A: DeviceElem, I: DeviceInterpolate {
        Ok(QuantileExtB200::quantile_axis_mut(self, Axis(0), q, i)?.into_scalar())              // mod.rs:613-621
    }
    fn quantiles_mut<I>(&mut self, qs: &ArrayRef<N64, Ix1>, i: &I) -> Result<Array1<A>, QuantileError>
    where A: DeviceElem, I: DeviceInterpolate {
        QuantileExtB200::quantiles_axis_mut(self, Axis(0), qs, i)                                // mod.rs:623-633
    }
}

/// `Sort1dExt::{get_from_sorted_mut, get_many_from_sorted_mut}` (`src/sort.rs:99-131`): the values at the given sorted
/// positions; the array is left as it is.  `partition_mut` (`src/sort.rs:133-168`) works in place and leaves the
/// reference's arrangement.
pub trait Sort1dExtB200<A>: private::SealedArray {
    fn partition_mut(&mut self, pivot_index: usize) -> usize where A: DeviceElem;
    fn get_from_sorted_mut(&mut self, i: usize) -> A where A: DeviceElem;
    fn get_many_from_sorted_mut(&mut self, indexes: &ArrayRef<usize, Ix1>) -> IndexMap<usize, A> where A: DeviceElem;
}
impl<A> Sort1dExtB200<A> for ArrayRef<A, Ix1> {
    fn partition_mut(&mut self, pivot_index: usize) -> usize where A: DeviceElem {
        let mut new_index = 0usize;
        let st = CTX.with(|&ctx| unsafe {
            ffi::nds_partition(ctx, A::DTYPE, self.as_mut_ptr() as *mut _, self.len(), self.strides()[0],
                               pivot_index, &mut new_index)
        });
        assert_eq!(st, ffi::NDS_OK, "pivot_index out of bounds (the reference panics, sort.rs:77-78)");
        new_index
    }
    fn get_from_sorted_mut(&mut self, i: usize) -> A where A: DeviceElem {
        self.get_many_from_sorted_mut(&ndarray::aview1(&[i]))[&i]
    }
    fn get_many_from_sorted_mut(&mut self, indexes: &ArrayRef<usize, Ix1>) -> IndexMap<usize, A> where A: DeviceElem {
        let mut idx: Vec<usize> = indexes.iter().copied().collect();
        idx.sort_unstable();
        idx.dedup();                                           // sort.rs:113-116
        let mut out = Vec::<A>::with_capacity(idx.len());
        unsafe { out.set_len(idx.len()) };
        let st = CTX.with(|&ctx| unsafe {
            ffi::nds_get_many_from_sorted(ctx, A::DTYPE, self.as_ptr() as *const _, self.len(), self.strides()[0],
                                          idx.as_ptr(), idx.len(), out.as_mut_ptr() as *mut _)
        });
        assert_eq!(st, ffi::NDS_OK, "index out of bounds (the reference panics, sort.rs:27,38)");
        idx.into_iter().zip(out).collect()
    }
}
