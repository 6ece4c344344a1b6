//! Host-side helpers kept from the reference crate's `utils` module (same signatures): `MatExt`, `SetMValue`, checked
//! `imread`, the `EccMatchParameters -> TermCriteria` conversion and the `KeyPointMatchParameters` defaults.
use crate::{EccMatchParameters, StackerError};
use opencv::{imgcodecs, prelude::*};

/// Extension trait for more ergonomic Mat conversions (reference: src/utils.rs:9-23; part of the public `utils`
/// surface, so it stays — the GPU path itself never materialises the converted frame).
pub trait MatExt {
    /// Convert matrix to specified type with scaling: `dst = self * alpha + beta` as `rtype`.
    fn convert(&self, rtype: i32, alpha: f64, beta: f64) -> Result<Mat, StackerError>;
}

impl MatExt for Mat {
    fn convert(&self, rtype: i32, alpha: f64, beta: f64) -> Result<Mat, StackerError> {
        let mut dst = Mat::default();
        self.convert_to(&mut dst, rtype, alpha, beta)?;
        Ok(dst)
    }
}

/// Trait for setting a value in a 2d `Mat<T>` (reference: src/utils.rs:39-71).
pub trait SetMValue {
    fn set_2d<T: opencv::prelude::DataType>(&mut self, row: i32, col: i32, value: T) -> Result<(), StackerError>;
}

impl SetMValue for Mat {
    #[inline]
    /// ```
    /// # use libstacker::{prelude::*, opencv::prelude::*, opencv::prelude::MatTraitConst};
    /// # use crate::libstacker::utils::SetMValue;
    /// let mut m = unsafe { opencv::core::Mat::new_rows_cols(1, 3, opencv::core::CV_64FC1).unwrap() };
    /// m.set_2d::<f64>(0, 0, -1.0).unwrap();
    /// m.set_2d::<f64>(0, 1, -2.0).unwrap();
    /// m.set_2d::<f64>(0, 2, -3.0).unwrap();
    /// assert_eq!(-1.0, *m.at_2d::<f64>(0, 0).unwrap());
    /// assert_eq!(-2.0, *m.at_2d::<f64>(0, 1).unwrap());
    /// assert_eq!(-3.0, *m.at_2d::<f64>(0, 2).unwrap());
    /// ```
    fn set_2d<T: opencv::prelude::DataType>(&mut self, row: i32, col: i32, value: T) -> Result<(), StackerError> {
        let v = self.at_2d_mut::<T>(row, col)?;
        *v = value;
        Ok(())
    }
}

/// `imread` with a checked path (reference: src/utils.rs:111-117).
#[inline(always)]
pub fn imread<P: AsRef<std::path::Path>>(path: P, imread_flags: i32) -> Result<Mat, StackerError> {
    let path_str = path
        .as_ref()
        .to_str()
        .ok_or_else(|| StackerError::InvalidPathEncoding(path.as_ref().to_path_buf()))?;
    Ok(imgcodecs::imread(path_str, imread_flags)?)
}

/// Decode as-is (IMREAD_UNCHANGED) and check what the device path needs: 8-bit, 3 or 4 channels, continuous.
/// The grey conversion and the `* 1/255` of the reference's `read_grey_and_f32` happen on the GPU.
pub(crate) fn read_frame(path: &std::path::Path) -> Result<Mat, StackerError> {
    let img = imread(path, imgcodecs::IMREAD_UNCHANGED)?;
    if img.depth() != opencv::core::CV_8U || !(img.channels() == 3 || img.channels() == 4) {
        // the reference fails here too: cvtColor(BGR2GRAY) / findTransformECC reject anything else
        return Err(StackerError::InvalidParams(format!(
            "unsupported image type (depth {}, {} channels) in {:?}", img.depth(), img.channels(), path
        )));
    }
    if img.is_continuous() { Ok(img) } else { Ok(img.try_clone()?) }
}

/// Aspect-preserving INTER_AREA downscale for the keypoint front end (reference: src/utils.rs:186-214): the
/// SMALLER dimension becomes `scale_down`, both new sizes are truncated.  (The ECC path does this on the GPU.)
pub(crate) fn scale_image(img: &Mat, scale_down: f32) -> Result<Mat, StackerError> {
    let size = img.size()?;
    let f = if size.width < size.height { scale_down as f64 / size.width as f64 } else { scale_down as f64 / size.height as f64 };
    let mut resized = Mat::default();
    opencv::imgproc::resize(
        img, &mut resized,
        opencv::core::Size::new((size.width as f64 * f) as i32, (size.height as f64 * f) as i32),
        0.0, 0.0, opencv::imgproc::INTER_AREA,
    )?;
    Ok(resized)
}

impl From<EccMatchParameters> for Result<opencv::core::TermCriteria, StackerError> {
    /// ```
    /// # use libstacker::{prelude::*, opencv::core::TermCriteria_Type};
    /// let t: Result<opencv::core::TermCriteria, StackerError> = EccMatchParameters {
    ///     motion_type: MotionType::Euclidean, max_count: None, epsilon: Some(0.1), gauss_filt_size: 3 }.into();
    /// let t = t.unwrap();
    /// assert_eq!(t.epsilon, 0.1);
    /// assert_eq!(t.typ, TermCriteria_Type::EPS as i32);
    /// ```
    fn from(r: EccMatchParameters) -> Result<opencv::core::TermCriteria, StackerError> {
        let mut rv = opencv::core::TermCriteria::default()?;
        if let Some(max_count) = r.max_count {
            rv.typ |= opencv::core::TermCriteria_Type::COUNT as i32;
            rv.max_count = max_count;
        }
        if let Some(epsilon) = r.epsilon {
            rv.typ |= opencv::core::TermCriteria_Type::EPS as i32;
            rv.epsilon = epsilon;
        }
        Ok(rv)
    }
}

impl Default for crate::KeyPointMatchParameters {
    fn default() -> Self {
        Self {
            method: opencv::calib3d::RANSAC,
            ransac_reproj_threshold: 3.0,
            match_keep_ratio: 0.75,
            match_ratio: 0.8,
            border_mode: opencv::core::BORDER_CONSTANT,
            border_value: opencv::core::Scalar::default(),
        }
    }
}
