//! `h263-rs-yuv` (`yuv/src/lib.rs`, `yuv/src/bt601.rs`) with the conversion on the GPU.
pub mod bt601 {
    /// Convert planar YUV 4:2:0 data into interleaved RGBA 8888 data (`bt601.rs:105-196`): BT.601 limited range to
    /// full-range RGB, nearest chroma sample, alpha 255.  Same signature, same preconditions (`bt601.rs:100-104`):
    /// `y.len()` a multiple of `y_width`, chroma planes of `ceil(w/2) x ceil(h/2)`; empty in, empty out.
    pub fn yuv420_to_rgba(y: &[u8], chroma_b: &[u8], chroma_r: &[u8], y_width: usize) -> Vec<u8> {
        if y.is_empty() {
            return Vec::new(); // bt601.rs:106-112
        }
        let h = y.len() / y_width;
        let (cw, ch) = ((y_width + 1) / 2, (h + 1) / 2);
        // the reference indexes out of bounds (and aborts) on short planes; check up front instead
        assert!(y.len() % y_width == 0 && chroma_b.len() >= cw * ch && chroma_r.len() >= cw * ch, "plane sizes");
        let mut out = vec![0u8; y.len() * 4];
        let rc = unsafe {
            h263cu_sys::h263cu_yuv420_to_rgba(y.as_ptr(), chroma_b.as_ptr(), chroma_r.as_ptr(), y.len(), y_width, out.as_mut_ptr())
        };
        assert_eq!(rc, 0, "h263cu_yuv420_to_rgba failed: {}", rc);
        out
    }

    #[cfg(test)]
    mod tests {
        use super::yuv420_to_rgba;
        // vectors of bt601.rs:199-225 (test_yuv_to_rgb), through the 1 x 1 picture form
        #[test]
        fn test_yuv_to_rgb() {
            let px = |y: u8, cb: u8, cr: u8| yuv420_to_rgba(&[y], &[cb], &[cr], 1);
            for (y, v) in [(17u8, 1u8), (16, 0), (15, 0), (0, 0), (234, 254), (235, 255), (236, 255), (255, 255), (125, 127), (126, 128)] {
                assert_eq!(px(y, 128, 128), vec![v, v, v, 255], "y = {}", y);
            }
        }
    }
}
