//! `h263-rs-deblock` (`deblock/src/lib.rs`, `deblock/src/deblock.rs`) with the filter on the GPU.
pub mod deblock {
    /// `deblock.rs:5-8`
    pub const QUANT_TO_STRENGTH: [u8; 32] = [
        0, 1, 1, 2, 2, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9, 10, 10, 10, 11, 11, 11, 12, 12, 12,
    ];

    /// Applies the deblocking filter to the horizontal and vertical block edges of one plane
    /// (`deblock.rs:305-315`): a post-process on the decoder's output, never fed back as a reference
    /// (`deblock/src/lib.rs:1-2`).  `strength` is 1..=12 (`deblock.rs:30`).
    pub fn deblock(data: &[u8], width: usize, strength: u8) -> Vec<u8> {
        debug_assert!((1..=12).contains(&strength));
        if data.is_empty() {
            return Vec::new();
        }
        assert!(width > 0 && data.len() % width == 0, "plane size");
        let mut out = vec![0u8; data.len()];
        let rc = unsafe { h263cu_sys::h263cu_deblock(data.as_ptr(), data.len(), width, strength, out.as_mut_ptr()) };
        assert_eq!(rc, 0, "h263cu_deblock failed: {}", rc);
        out
    }

    #[cfg(test)]
    mod tests {
        use super::*;
        #[test]
        fn table_matches_the_library() {
            assert_eq!(QUANT_TO_STRENGTH, unsafe { h263cu_sys::h263cu_quant_to_strength });
        }
        #[test]
        fn flat_planes_are_left_alone() {
            // deblock.rs:324-349: a flat quadruple passes through process() unchanged
            let plane = vec![77u8; 32 * 24];
            assert_eq!(deblock(&plane, 32, 5), plane);
        }
    }
}
