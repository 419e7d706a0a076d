//! Decodes a packet file with the real h263-rs crates and prints one line of FNV-1a hashes per
//! picture (Y, Cb, Cr, RGBA, deblocked RGBA) in the format of `dump_stream.py`, so that the two
//! outputs can be diffed.  File format: repeated { u32 little-endian length, packet bytes },
//! one packet per picture (Sorenson Spark ends a picture at end of input, state.rs:193,411).
use h263_rs::parser::H263Reader;
use h263_rs::{DecoderOption, H263State};
use h263_rs_deblock::deblock::{deblock, QUANT_TO_STRENGTH};
use h263_rs_yuv::bt601::yuv420_to_rgba;
use std::io::Read;

fn fnv(data: &[u8]) -> u64 {
    let mut h: u64 = 0xcbf29ce484222325;
    for b in data {
        h ^= *b as u64;
        h = h.wrapping_mul(0x100000001b3);
    }
    h
}

fn main() {
    let path = std::env::args().nth(1).expect("usage: crosscheck <packets.h263pk>");
    let mut file = std::fs::File::open(path).expect("open");
    let mut blob = Vec::new();
    file.read_to_end(&mut blob).expect("read");
    let mut state = H263State::new(DecoderOption::SORENSON_SPARK_BITSTREAM);
    let mut pos = 0usize;
    let mut index = 0usize;
    while pos + 4 <= blob.len() {
        let len = u32::from_le_bytes([blob[pos], blob[pos + 1], blob[pos + 2], blob[pos + 3]]) as usize;
        pos += 4;
        let packet = &blob[pos..pos + len];
        pos += len;
        let mut reader = H263Reader::from_source(packet);
        match state.decode_next_picture(&mut reader) {
            Err(e) => println!("{} error {:?}", index, e),
            Ok(()) => {
                let pic = state.get_last_picture().expect("picture");
                let (y, cb, cr) = pic.as_yuv();
                let w = pic.luma_samples_per_row();
                let cw = pic.chroma_samples_per_row();
                let rgba = yuv420_to_rgba(y, cb, cr, w);
                let s = QUANT_TO_STRENGTH[pic.as_header().quantizer as usize];
                let (dy, dcb, dcr) = (deblock(y, w, s), deblock(cb, cw, s), deblock(cr, cw, s));
                let drgba = yuv420_to_rgba(&dy, &dcb, &dcr, w);
                println!("{} {:016x} {:016x} {:016x} {:016x} {:016x}", index, fnv(y), fnv(cb), fnv(cr), fnv(&rgba), fnv(&drgba));
            }
        }
        index += 1;
    }
}
