//! libstacker — same public API as eadf/libstacker.rs, with the ECC align-and-stack path (and the Tenengrad
//! metric and the final warp/accumulate of `keypoint_match`) running on an NVIDIA B200 through the C ABI in
//! `include/stacker_cuda.h`.  File decode and the ORB / BFMatcher / findHomography stages stay on the host
//! in OpenCV, exactly where the reference has them.
//!
//! This crate is provided as the binding a maintainer adds; it is not compiled in this repository (no Rust
//! toolchain in the build image).  The C++ (`../host`) and Python (`../api.py`) mirrors drive the same ABI
//! and are what the test-suite runs.
pub mod ffi;
pub mod utils;

pub use opencv;
use opencv::core::{self, Mat, Point2f, Vector};
use opencv::{calib3d, features2d, imgproc, prelude::*};
use rayon::prelude::*;
use std::path::PathBuf;
use thiserror::Error;

#[derive(Error, Debug)]
pub enum StackerError {
    #[error(transparent)]
    OpenCvError(#[from] opencv::Error),
    #[error("Not enough files")]
    NotEnoughFiles,
    #[error("Not implemented")]
    NotImplemented,
    #[error(transparent)]
    IoError(#[from] std::io::Error),
    // kept for source compatibility with callers that match on it (reference src/lib.rs:37-38); this crate holds no
    // MatExpr behind a lock, so it is never produced here
    #[error(transparent)]
    PoisonError(#[from] std::sync::PoisonError<core::MatExprResult<core::MatExpr>>),
    #[error("Invalid path encoding {0}")]
    InvalidPathEncoding(PathBuf),
    #[error("Invalid parameter(s) {0}")]
    InvalidParams(String),
    #[error("Internal error {0}")]
    ProcessingError(String),
}

/// Status code of the CUDA library -> the variant the reference would have produced at that point.
fn check(rc: i32) -> Result<(), StackerError> {
    use ffi::*;
    match rc {
        STK_OK => Ok(()),
        // findTransformECC's StsNoConv / CV_Assert surface as opencv::Error in the reference (src/lib.rs:777)
        STK_ERR_ECC_NOCONV | STK_ERR_ECC_NAN | STK_ERR_CRITERIA => Err(StackerError::OpenCvError(opencv::Error::new(
            core::StsNoConv, last_error()))),
        STK_ERR_BAD_ARG => Err(StackerError::InvalidParams(last_error())),
        STK_ERR_NOT_ENOUGH => Err(StackerError::NotEnoughFiles),
        STK_ERR_UNSUPPORTED => Err(StackerError::NotImplemented),
        _ => Err(StackerError::ProcessingError(last_error())),
    }
}

#[derive(Debug, Clone, Copy)]
pub struct KeyPointMatchParameters {
    pub method: i32,
    pub ransac_reproj_threshold: f64,
    pub match_keep_ratio: f32,
    pub match_ratio: f32,
    pub border_mode: i32,
    pub border_value: core::Scalar,
}

#[derive(Debug, Copy, Clone, PartialEq, Eq)]
pub enum MotionType {
    Homography = opencv::video::MOTION_HOMOGRAPHY as isize,
    Affine = opencv::video::MOTION_AFFINE as isize,
    Euclidean = opencv::video::MOTION_EUCLIDEAN as isize,
    Translation = opencv::video::MOTION_TRANSLATION as isize,
}

#[derive(Debug, Copy, Clone)]
pub struct EccMatchParameters {
    pub motion_type: MotionType,
    pub max_count: Option<i32>,
    pub epsilon: Option<f64>,
    pub gauss_filt_size: i32,
}

/// CUDA devices a stack is sharded over: all of the box, or the first `STACKER_GPUS` of them.
fn devices() -> Result<Vec<i32>, StackerError> {
    let mut n = 0;
    check(unsafe { ffi::stk_device_count(&mut n) })?;
    let want = std::env::var("STACKER_GPUS").ok().and_then(|v| v.parse::<i32>().ok()).unwrap_or(n);
    Ok((0..n.min(want).max(1)).collect())
}

fn new_ctx_on(first: &Mat, ecc: Option<(EccMatchParameters, core::TermCriteria)>, ecc_size: Option<(i32, i32)>,
              device: i32, seed_reference: bool) -> Result<ffi::Ctx, StackerError> {
    let mut cfg = ffi::stk_ecc_config {
        width: first.cols(), height: first.rows(), channels: first.channels(),
        device, lanes: 0, seed_reference: seed_reference as i32, ..Default::default()
    };
    if let Some((w, h)) = ecc_size {
        cfg.ecc_width = w;
        cfg.ecc_height = h;
    }
    if let Some((p, c)) = ecc {
        cfg.align = 1;
        cfg.motion_type = p.motion_type as i32;
        cfg.criteria_type = c.typ;
        cfg.max_count = c.max_count;
        cfg.epsilon = c.epsilon;
        cfg.gauss_filt_size = p.gauss_filt_size;
    }
    let mut raw = std::ptr::null_mut();
    check(unsafe { ffi::stk_ecc_create(&cfg, &mut raw) })?;
    Ok(ffi::Ctx(raw))
}

fn finish(ctx: &ffi::Ctx, first: &Mat, divisor: usize) -> Result<Mat, StackerError> {
    let typ = core::CV_MAKETYPE(core::CV_32F, first.channels());
    let mut out = unsafe { Mat::new_rows_cols(first.rows(), first.cols(), typ)? };
    let pitch = out.mat_step().get(0);
    check(unsafe { ffi::stk_ecc_finish(ctx.0, divisor as i32, out.data_mut() as *mut f32, pitch) })?;
    Ok(out)
}

/// Aligns every image to the first with ECC and averages them.  Same contract as the reference
/// (`src/lib.rs:702-717`): `Ok(Mat)` is CV_32FC3 in [0, 1]; `scale_down_width` selects the scaled variant.
pub fn ecc_match<I, P>(files: I, params: EccMatchParameters, scale_down_width: Option<f32>) -> Result<Mat, StackerError>
where
    I: IntoIterator<Item = P>,
    P: AsRef<std::path::Path>,
{
    let files: Vec<PathBuf> = files.into_iter().map(|p| p.as_ref().to_path_buf()).collect();
    if files.is_empty() {
        return Err(StackerError::NotEnoughFiles);
    }
    let criteria = Result::<core::TermCriteria, StackerError>::from(params)?;
    let first = utils::read_frame(&files[0])?;
    // ecc_match_scaling_down (reference src/lib.rs:849-1028): the library validates the width
    // (:876-888), applies utils::scale_image's size rule, resizes the greys with INTER_AREA on the GPU,
    // runs ECC there and takes the matrix back to full resolution before the full-size warp.
    let ecc_size = match scale_down_width {
        Some(sd) => {
            let (mut sw, mut sh) = (0, 0);
            check(unsafe { ffi::stk_scaled_size(first.cols(), first.rows(), sd, &mut sw, &mut sh) })?;
            Some((sw, sh))
        }
        None => None,
    };
    // one context per GPU of the box: every device gets frame 0 (only the first seeds its accumulator with it,
    // reference src/lib.rs:752-754), the other frames are dealt round-robin
    let devs = devices()?;
    let mut ctxs = Vec::with_capacity(devs.len());
    for (k, d) in devs.iter().enumerate() {
        let c = new_ctx_on(&first, Some((params, criteria)), ecc_size, *d, k == 0)?;
        check(unsafe { ffi::stk_ecc_set_reference(c.0, first.data(), first.mat_step().get(0)) })?;
        ctxs.push(c);
    }
    if ctxs.len() > 1 {
        let raw: Vec<*mut ffi::stk_ecc_ctx> = ctxs.iter().map(|c| c.0).collect();
        check(unsafe { ffi::stk_ecc_peer_connect_local(raw.as_ptr(), raw.len() as i32) })?;
    }
    let ctxs = &ctxs;
    // Decode on the Rayon pool, one task per frame (reference: src/lib.rs:746-749).  A task decodes its frame and copies
    // the rows into a pinned ring buffer of the context the frame is dealt to (no library lock is held while it copies;
    // `imgcodecs::imdecode_to` on a Mat wrapped around `buf` would save this copy too).  The buffers are handed over in
    // FILE ORDER by this thread, window by window, so the f32 summation order — and therefore the result — does not
    // depend on thread timing (the reference's own order is whatever Rayon's split gives).  A window never exceeds one
    // ring (2 buffers per lane, 4 lanes), so every task of a window gets its buffer.
    struct Filled(*mut u8);
    unsafe impl Send for Filled {}
    let window = 8usize;
    let order: Vec<usize> = (1..files.len()).collect();
    for chunk in order.chunks(window * ctxs.len()) {
        let filled: Vec<Filled> = chunk.par_iter().with_min_len(1).map(|&i| -> Result<Filled, StackerError> {
            let img = utils::read_frame(&files[i])?;
            if img.size()? != first.size()? || img.channels() != first.channels() {
                return Err(StackerError::InvalidParams(format!("{:?}: size differs from the first frame", files[i])));
            }
            let (mut buf, mut pitch) = (std::ptr::null_mut::<u8>(), 0usize);
            let ctx = &ctxs[i % ctxs.len()];
            check(unsafe { ffi::stk_ecc_acquire_frame_buffer(ctx.0, &mut buf, &mut pitch) })?;
            let step = img.mat_step().get(0);
            for y in 0..img.rows() as usize {
                unsafe { std::ptr::copy_nonoverlapping(img.data().add(y * step), buf.add(y * pitch), pitch) };
            }
            Ok(Filled(buf))
        }).collect::<Result<Vec<_>, _>>()?;
        for (&i, f) in chunk.iter().zip(filled.iter()) {
            check(unsafe { ffi::stk_ecc_submit_acquired(ctxs[i % ctxs.len()].0, f.0, i as i64) })?;
        }
    }
    if ctxs.len() == 1 {
        return finish(&ctxs[0], &first, files.len());
    }
    finish_on_devices(ctxs, &first, files.len())
}

/// Rayon's try_reduce + `/ n` (reference src/lib.rs:819-839, :319-346) as ONE exchange step over NVLink peer memory:
/// every device reduces and scales its slice of the stack and copies it into the result Mat over its own PCIe link.
/// All exchanges are queued before the first copy-out (a copy into pageable memory blocks this thread).
fn finish_on_devices(ctxs: &[ffi::Ctx], first: &Mat, divisor: usize) -> Result<Mat, StackerError> {
    let typ = core::CV_MAKETYPE(core::CV_32F, first.channels());
    let mut out = unsafe { Mat::new_rows_cols(first.rows(), first.cols(), typ)? };
    for c in ctxs.iter() {
        check(unsafe { ffi::stk_ecc_peer_reduce_scatter(c.0, divisor as i32, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) })?;
    }
    for c in ctxs.iter() {
        check(unsafe { ffi::stk_ecc_peer_slice_to_host(c.0, out.data_mut() as *mut f32) })?;
    }
    for c in ctxs.iter() {
        check(unsafe { ffi::stk_ecc_sync(c.0) })?;      // also surfaces a frame's ECC failure (reference :777)
    }
    Ok(out)
}

/// Feature-based alignment: ORB / BFMatcher / findHomography on the host exactly as the reference
/// (`src/lib.rs:146-287`), final `warp_perspective` + accumulate + divide on the GPU (`:289-350`).
pub fn keypoint_match<I, P>(files: I, params: KeyPointMatchParameters, scale_down_width: Option<f32>) -> Result<(i32, Mat), StackerError>
where
    I: IntoIterator<Item = P>,
    P: AsRef<std::path::Path>,
{
    let files: Vec<PathBuf> = files.into_iter().map(|p| p.as_ref().to_path_buf()).collect();
    if files.is_empty() {
        return Err(StackerError::NotEnoughFiles);
    }
    let first = utils::read_frame(&files[0])?;
    // keypoint_match_scale_down (reference src/lib.rs:355-600): only the upper bound is validated (:378-383)
    if let Some(sd) = scale_down_width {
        if sd >= first.cols() as f32 {
            return Err(StackerError::InvalidParams(format!(
                "scale_down_to was larger (or equal) to the full image width: full_size:{}, scale_down_to:{}", first.cols(), sd)));
        }
    }
    // grey plane the features are detected on: full size, or utils::scale_image'd (INTER_AREA) like the reference
    let grey = |img: &Mat| -> Result<Mat, StackerError> {
        let mut g = Mat::default();
        imgproc::cvt_color(img, &mut g, imgproc::COLOR_BGR2GRAY, 0, core::AlgorithmHint::ALGO_HINT_DEFAULT)?;
        match scale_down_width {
            Some(sd) => utils::scale_image(&g, sd),
            None => Ok(g),
        }
    };
    let orb = |g: &Mat| -> Result<(Vector<core::KeyPoint>, Mat), StackerError> {
        let mut orb = features2d::ORB::create_def()?;
        let (mut kp, mut des) = (Vector::new(), Mat::default());
        orb.detect_and_compute(g, &Mat::default(), &mut kp, &mut des, false)?;
        Ok((kp, des))
    };
    let (kp0, des0) = orb(&grey(&first)?)?;
    // one warp-only context per GPU of the box (BASELINE configs[4]): every device gets frame 0 (only the first seeds
    // its accumulator with it), accepted frames are dealt by index, the partial stacks meet in ONE exchange + divide
    let devs = devices()?;
    let mut ctxs = Vec::with_capacity(devs.len());
    for (k, d) in devs.iter().enumerate() {
        let c = new_ctx_on(&first, None, None, *d, k == 0)?;
        check(unsafe { ffi::stk_ecc_set_reference(c.0, first.data(), first.mat_step().get(0)) })?;
        ctxs.push(c);
    }
    if ctxs.len() > 1 {
        let raw: Vec<*mut ffi::stk_ecc_ctx> = ctxs.iter().map(|c| c.0).collect();
        check(unsafe { ffi::stk_ecc_peer_connect_local(raw.as_ptr(), raw.len() as i32) })?;
    }
    let ctxs = &ctxs;
    let border = [params.border_value[0], params.border_value[1], params.border_value[2], params.border_value[3]];
    let dropped: i32 = (1..files.len()).into_par_iter().with_min_len(1).map(|i| -> Result<i32, StackerError> {
        let img = utils::read_frame(&files[i])?;
        let g = grey(&img)?;
        let small = g.size()?;
        let (kp, des) = orb(&g)?;
        let mut matcher = features2d::BFMatcher::create(core::NORM_HAMMING, false)?;
        matcher.add(&des)?;
        let mut knn = Vector::<Vector<core::DMatch>>::new();
        matcher.knn_match(&des0, &mut knn, 2, &Mat::default(), false)?;
        let mut good: Vec<core::DMatch> = knn.iter()
            .filter_map(|m| (m.len() == 2 && m.get(0).unwrap().distance < params.match_ratio * m.get(1).unwrap().distance)
                .then(|| m.get(0).unwrap()))
            .collect();
        good.sort_by(|a, b| a.distance.partial_cmp(&b.distance).unwrap_or(std::cmp::Ordering::Equal));
        good.truncate((good.len() as f32 * params.match_keep_ratio).round() as usize);
        if good.len() < 5 {
            return Ok(1);
        }
        let mut src = Vector::<Point2f>::with_capacity(good.len());
        let mut dst = Vector::<Point2f>::with_capacity(good.len());
        for m in &good {
            src.push(kp0.get(m.query_idx as usize)?.pt());
            dst.push(kp.get(m.train_idx as usize)?.pt());
        }
        let h = match calib3d::find_homography(&dst, &src, &mut Mat::default(), params.method, params.ransac_reproj_threshold) {
            Ok(h) => h,
            Err(_) => return Ok(1),
        };
        if h.empty() || h.rows() != 3 || h.cols() != 3 || core::determinant(&h)?.abs() < 1e-6 {
            return Ok(1);
        }
        let mut hv: Vec<f64> = h.data_typed::<f64>()?.to_vec();
        if scale_down_width.is_some() {
            // adjust_homography_for_scale_f64 (reference src/utils.rs:218-248)
            let (sx, sy) = (img.cols() as f64 / small.width as f64, img.rows() as f64 / small.height as f64);
            hv[2] *= sx;
            hv[5] *= sy;
            hv[6] /= sx;
            hv[7] /= sy;
        }
        check(unsafe {
            ffi::stk_ecc_submit_warp(ctxs[i % ctxs.len()].0, img.data(), img.mat_step().get(0), hv.as_ptr(), params.border_mode, border.as_ptr(), i as i64)
        })?;
        Ok(0)
    }).try_reduce(|| 0, |a, b| Ok(a + b))?;
    if files.len() as i32 - dropped <= 0 {
        return Err(StackerError::InvalidParams("All images discarded: try modifying KeyPointMatchParameters::match_distance_threshold".into()));
    }
    let kept = files.len() - dropped as usize;
    if ctxs.len() == 1 {
        return Ok((dropped, finish(&ctxs[0], &first, kept)?));
    }
    Ok((dropped, finish_on_devices(ctxs, &first, kept)?))
}

/// Tenengrad sharpness (Krotkov86) of an 8-bit single-channel image; bit-identical to the reference's
/// CV_64F Sobel pipeline (`src/lib.rs:1101-1147`).
pub fn sharpness_tenengrad(src_grey_mat: &Mat, k_size: i32) -> Result<f64, StackerError> {
    if ![1, 3, 5, 7].contains(&k_size) {
        return Err(StackerError::InvalidParams("Kernel size must be 1, 3, 5, or 7".into()));
    }
    if src_grey_mat.depth() != core::CV_8U || src_grey_mat.channels() != 1 {
        return Err(StackerError::NotImplemented);
    }
    let mut out = 0f64;
    check(unsafe {
        ffi::stk_tenengrad(src_grey_mat.data(), src_grey_mat.mat_step().get(0), src_grey_mat.cols(), src_grey_mat.rows(), 1, k_size, -1, &mut out)
    })?;
    Ok(out)
}

/// LAPM, LAPV, TENG(3) and GLVN of an 8-bit single-channel image in one pass over the plane
/// (`examples/main.rs:43-46` computes exactly these four per file).
pub fn sharpness_all(src_grey_mat: &Mat) -> Result<[f64; 4], StackerError> {
    if src_grey_mat.depth() != core::CV_8U || src_grey_mat.channels() != 1 {
        return Err(StackerError::NotImplemented);
    }
    let mut out = [0f64; 4];
    check(unsafe {
        ffi::stk_sharpness_all(src_grey_mat.data(), src_grey_mat.mat_step().get(0), src_grey_mat.cols(), src_grey_mat.rows(), 1, -1, out.as_mut_ptr())
    })?;
    Ok(out)
}

/// 'LAPM' (Nayar89) — same contract as the reference (`src/lib.rs:1032-1068`).
pub fn sharpness_modified_laplacian(src_mat: &Mat) -> Result<f64, StackerError> {
    Ok(sharpness_all(src_mat)?[0])
}

/// 'LAPV' (Pech2000) — same contract as the reference (`src/lib.rs:1070-1090`).
pub fn sharpness_variance_of_laplacian(src_mat: &Mat) -> Result<f64, StackerError> {
    Ok(sharpness_all(src_mat)?[1])
}

/// 'GLVN' (Santos97) — same contract as the reference (`src/lib.rs:1151-1166`).
pub fn sharpness_normalized_gray_level_variance(src_mat: &Mat) -> Result<f64, StackerError> {
    Ok(sharpness_all(src_mat)?[3])
}

pub mod prelude {
    pub use super::{EccMatchParameters, KeyPointMatchParameters, MotionType, StackerError, ecc_match, keypoint_match};
}
