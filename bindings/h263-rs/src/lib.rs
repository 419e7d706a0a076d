//! `h263-rs` with its reconstruction path on a B200: the public API of ruffle-rs/h263-rs
//! (`h263/src/lib.rs:16-18`: `H263State`, `DecoderOption`, `Error`, `Result`, `PictureTypeCode`, `parser::H263Reader`)
//! as a thin facade over `libh263cu.so`.  Call sites do not change:
//!
//! ```ignore
//! let mut state = H263State::new(DecoderOption::SORENSON_SPARK_BITSTREAM);
//! let mut reader = H263Reader::from_source(&packet[..]);      // one reader per picture packet (FLV video tag)
//! state.decode_next_picture(&mut reader)?;
//! let (y, cb, cr) = state.get_last_picture().unwrap().as_yuv();
//! ```
//!
//! What differs from the reference, and why:
//! * The bitstream parse runs in the library's host front end, so `H263Reader` here only OWNS the source: it hands
//!   the packet's bytes to the parser.  The reference's `H263Reader` cannot yield its source
//!   (`h263/src/parser/reader.rs:15-43` keeps it private and offers bit reads only), so the facade brings its own
//!   reader type with the same name and constructor instead of asking for a patch of the reference tree.
//!   A Sorenson picture ends at the end of its source (`state.rs:193,411`), which is why one reader per packet is
//!   the only usable form in the reference as well.
//! * `decode_next_picture` stays a transaction (`state.rs:120-137`): on `Err` neither the reader nor the decoder has
//!   moved; the bytes read from the source so far stay buffered in the reader, so a retry (after more data has been
//!   streamed into the source) sees the same bits.
//! * Without a usable GPU every decode fails with `Error::NoDevice`: there is no CPU fallback.
use std::io::Read;

use h263cu_sys as sys;

bitflags::bitflags! {
    /// Options which influence the decoding of a bitstream (`decoder/types.rs:3-18`).
    #[derive(Copy, Clone)]
    pub struct DecoderOption : u8 {
        const SORENSON_SPARK_BITSTREAM = 0b1;
        const USE_SCALABILITY_MODE = 0b10;
    }
}

/// `h263::Error` (`error.rs:6-57`), plus the library's own failures.
#[derive(Debug)]
pub enum Error {
    InternalDecoderError,
    MiddleOfBitstream,
    InvalidMacroblockHeader,
    InvalidMacroblockCodedBits,
    InvalidIntraDc,
    InvalidShortCoefficient,
    InvalidLongCoefficient,
    InvalidMvd,
    InvalidPType,
    InvalidPlusPType,
    InvalidGobHeader,
    InvalidBitstream,
    PictureFormatMissing,
    PictureFormatInvalid,
    UncodedIFrameBlocks,
    UnhandledIoError(std::io::Error),
    UnimplementedDecoding,
    /// Input on which the reference itself panics (`panic = "abort"`): reported instead of aborting.
    ReferenceWouldAbort,
    /// No usable CUDA device, a CUDA failure, or a capacity limit of the device context.
    NoDevice,
    Cuda,
    Capacity,
    OutOfMemory,
}

impl Error {
    /// Maps a library return code (`include/h263cu.h`) to the reference's error.
    pub fn from_code(code: i32) -> Self {
        match code {
            -1 => Error::InternalDecoderError,
            -2 => Error::MiddleOfBitstream,
            -3 => Error::InvalidMacroblockHeader,
            -4 => Error::InvalidMacroblockCodedBits,
            -5 => Error::InvalidIntraDc,
            -6 => Error::InvalidShortCoefficient,
            -7 => Error::InvalidLongCoefficient,
            -8 => Error::InvalidMvd,
            -9 => Error::InvalidPType,
            -10 => Error::InvalidPlusPType,
            -11 => Error::InvalidGobHeader,
            -12 => Error::InvalidBitstream,
            -13 => Error::PictureFormatMissing,
            -14 => Error::PictureFormatInvalid,
            -15 => Error::UncodedIFrameBlocks,
            -16 => Error::UnhandledIoError(std::io::Error::from(std::io::ErrorKind::UnexpectedEof)),
            -17 => Error::UnimplementedDecoding,
            -104 => Error::ReferenceWouldAbort,
            -102 => Error::NoDevice,
            -103 => Error::Capacity,
            -106 => Error::OutOfMemory,
            _ => Error::Cuda,
        }
    }
    /// EOF errors end the current picture (`error.rs:66-76`).
    pub fn is_eof_error(&self) -> bool {
        matches!(self, Error::UnhandledIoError(e) if e.kind() == std::io::ErrorKind::UnexpectedEof)
    }
    /// `error.rs:78-86`
    pub fn is_macroblock_error(&self) -> bool {
        matches!(self, Error::InvalidMacroblockHeader | Error::InvalidMacroblockCodedBits)
    }
    /// `error.rs:88-93`
    pub fn is_gob_error(&self) -> bool {
        matches!(self, Error::InvalidGobHeader)
    }
}

impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{:?}", self)
    }
}
impl std::error::Error for Error {}
impl From<std::io::Error> for Error {
    fn from(e: std::io::Error) -> Self {
        Error::UnhandledIoError(e)
    }
}
pub type Result<T> = std::result::Result<T, Error>;

/// `types.rs:251-288`
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum PictureTypeCode {
    IFrame,
    PFrame,
    PbFrame,
    ImprovedPbFrame,
    BFrame,
    EiFrame,
    EpFrame,
    Reserved(u8),
    DisposablePFrame,
}

/// The header fields the reference keeps with a decoded picture (`types.rs:13-100`), as far as the decode path
/// produces them: Sorenson and baseline pictures carry no PLUSPTYPE extensions (`picture.rs:611-660`).
#[derive(Clone, Debug)]
pub struct Picture {
    pub version: Option<u8>,
    pub temporal_reference: u16,
    /// (width, height) of the source format (`SourceFormat::into_width_and_height`, `types.rs:168-182`)
    pub format: Option<(u16, u16)>,
    pub picture_type: PictureTypeCode,
    pub quantizer: u8,
    /// `PictureOption::USE_DEBLOCKER`: Sorenson's advisory deblocking flag (`types.rs:213-216`)
    pub use_deblocker: bool,
}

pub mod parser {
    //! `h263::parser`: the reader.  The bit-level functions of the reference's parser module are not part of the
    //! facade: the parse happens inside the library (`h263cu_parse_picture`).
    use std::io::Read;

    /// Same name and constructor as `h263::parser::H263Reader` (`reader.rs:15-43`).  Owns the source; the bytes it
    /// has read stay buffered until a picture has been decoded from them (`commit`, `reader.rs:376-389`).
    pub struct H263Reader<R: Read> {
        source: R,
        buffer: Vec<u8>,
        exhausted: bool,
    }

    impl<R: Read> H263Reader<R> {
        pub fn from_source(source: R) -> Self {
            Self { source, buffer: Vec::new(), exhausted: false }
        }
        /// Everything the source holds from the current position on (read to its end once).
        pub(crate) fn packet(&mut self) -> std::io::Result<&[u8]> {
            if !self.exhausted {
                self.source.read_to_end(&mut self.buffer)?;
                self.exhausted = true;
            }
            Ok(&self.buffer)
        }
        /// A picture has been decoded from the buffered bytes: they are consumed (a Sorenson picture takes its whole
        /// packet, `state.rs:193,411`).  More data streamed into the source afterwards forms the next picture.
        pub(crate) fn commit(&mut self) {
            self.buffer.clear();
            self.exhausted = false;
        }
    }
}
pub use parser::H263Reader;

/// `h263::DecodedPicture` (`decoder/picture.rs:8-143`): header + tight row-major planes.
pub struct DecodedPicture {
    header: Picture,
    width: usize,
    luma: Vec<u8>,
    chroma_b: Vec<u8>,
    chroma_r: Vec<u8>,
}

impl DecodedPicture {
    pub fn as_header(&self) -> &Picture {
        &self.header
    }
    /// (width, height) in luma samples (`picture.rs:66-68` returns the `SourceFormat`; the facade returns what
    /// `SourceFormat::into_width_and_height` would give).
    pub fn format(&self) -> (u16, u16) {
        (self.width as u16, (self.luma.len() / self.width.max(1)) as u16)
    }
    pub fn as_luma(&self) -> &[u8] {
        &self.luma
    }
    pub fn as_chroma_b(&self) -> &[u8] {
        &self.chroma_b
    }
    pub fn as_chroma_r(&self) -> &[u8] {
        &self.chroma_r
    }
    pub fn luma_samples_per_row(&self) -> usize {
        self.width
    }
    pub fn chroma_samples_per_row(&self) -> usize {
        (self.width + 1) / 2 // picture.rs:95-97
    }
    pub fn as_yuv(&self) -> (&[u8], &[u8], &[u8]) {
        (&self.luma, &self.chroma_b, &self.chroma_r)
    }
}

/// `h263::H263State` (`decoder/state.rs:16-489`) for one stream: host parse + device reconstruction.
pub struct H263State {
    decoder_options: DecoderOption,
    device: i32,
    parser: *mut sys::h263cu_parser,
    ctx: *mut sys::h263cu_ctx,
    ctx_w: u32,
    ctx_h: u32,
    last: Option<DecodedPicture>,
}

// One state per stream, used from one thread at a time (`&mut self`), like the reference's.
unsafe impl Send for H263State {}

impl H263State {
    pub fn new(decoder_options: DecoderOption) -> Self {
        Self::with_device(decoder_options, 0)
    }
    /// EXTENSION beyond the reference: a state that decodes Sorenson disposable P pictures (shown, never predicted
    /// from) instead of failing them with `UnimplementedDecoding` like the reference (`macroblock.rs:461-465`).
    pub fn with_disposable_pictures(decoder_options: DecoderOption) -> Self {
        Self::with_bits(decoder_options, decoder_options.bits() as u32 | sys::H263CU_OPT_DECODE_DISPOSABLE, 0)
    }
    /// Like `new`, on a chosen CUDA device.
    pub fn with_device(decoder_options: DecoderOption, device: i32) -> Self {
        Self::with_bits(decoder_options, decoder_options.bits() as u32, device)
    }
    fn with_bits(decoder_options: DecoderOption, bits: u32, device: i32) -> Self {
        let parser = unsafe { sys::h263cu_parser_create(bits) };
        assert!(!parser.is_null(), "out of memory");
        Self { decoder_options, device, parser, ctx: std::ptr::null_mut(), ctx_w: 0, ctx_h: 0, last: None }
    }
    pub fn is_sorenson(&self) -> bool {
        self.decoder_options.contains(DecoderOption::SORENSON_SPARK_BITSTREAM)
    }
    pub fn get_last_picture(&self) -> Option<&DecodedPicture> {
        self.last.as_ref()
    }
    /// The reference picture is the last non-disposable picture; disposable pictures do not decode in the reference
    /// (`macroblock.rs:461-465`), so it is the last picture (`state.rs:72-78`).
    pub fn get_reference_picture(&self) -> Option<&DecodedPicture> {
        self.last.as_ref()
    }
    /// `state.rs:81-98` drops every picture but the last and the reference one; the device context never holds more.
    pub fn cleanup_buffers(&mut self) {}

    /// Header only, no decoder state changes (`state.rs:102-111`).  `Ok(None)` when the source does not start with a
    /// picture (`decode_picture`'s `Ok(None)`, `picture.rs:611-626`).
    pub fn parse_picture<R: Read>(&self, reader: &mut H263Reader<R>, _previous_picture: Option<&Picture>) -> Result<Option<Picture>> {
        let packet = reader.packet()?;
        let mut pic = sys::h263cu_pic::default();
        let rc = unsafe { sys::h263cu_peek_picture(self.decoder_options.bits() as u32, packet.as_ptr(), packet.len(), &mut pic) };
        match rc {
            0 => Ok(Some(header_of(&pic))),
            sys::H263CU_ERR_MIDDLE_OF_BITSTREAM => Ok(None),
            e => Err(Error::from_code(e)),
        }
    }

    /// Decode the next picture in the bitstream (`state.rs:138-489`).
    pub fn decode_next_picture<R: Read>(&mut self, reader: &mut H263Reader<R>) -> Result<()> {
        let options = unsafe { sys::h263cu_parser_options(self.parser) };
        let packet = reader.packet()?;
        let mut pic = sys::h263cu_pic::default();
        let rc = unsafe { sys::h263cu_peek_picture(options, packet.as_ptr(), packet.len(), &mut pic) };
        if rc != 0 {
            return Err(Error::from_code(rc));
        }
        if pic.width == 0 || pic.height == 0 {
            return Err(Error::PictureFormatInvalid);
        }
        // a larger picture needs a larger context: it replaces the old one only once the packet has decoded
        let (w, h) = (pic.width as u32, pic.height as u32);
        let mut fresh: *mut sys::h263cu_ctx = std::ptr::null_mut();
        let ctx = if self.ctx.is_null() || w > self.ctx_w || h > self.ctx_h {
            let mut err = 0;
            fresh = unsafe { sys::h263cu_create(self.device, 1, w.max(16), h.max(16), 0, &mut err) };
            if fresh.is_null() {
                return Err(Error::from_code(err));
            }
            fresh
        } else {
            self.ctx
        };
        let drop_fresh = |c: *mut sys::h263cu_ctx| {
            if !c.is_null() {
                unsafe { sys::h263cu_destroy(c) }
            }
        };
        // one call: parse on the host, upload, reconstruct; the parser advances only when the device stage accepted
        let parsers = [self.parser];
        let packets = [packet.as_ptr()];
        let lens = [packet.len()];
        let ids = [0u32];
        let mut perr = [0i32];
        let mut n_decoded = 0u32;
        let rc = unsafe {
            sys::h263cu_decode_step(ctx, parsers.as_ptr(), packets.as_ptr(), lens.as_ptr(), ids.as_ptr(), 1, 1, 0, std::ptr::null_mut(), 0, perr.as_mut_ptr(), &mut n_decoded)
        };
        let rc = if rc != 0 { rc } else { perr[0] };
        if rc != 0 {
            drop_fresh(fresh);
            return Err(Error::from_code(rc));
        }
        let rc = unsafe { sys::h263cu_sync(ctx) };
        if rc != 0 {
            drop_fresh(fresh);
            return Err(Error::from_code(rc));
        }
        if !fresh.is_null() {
            if !self.ctx.is_null() {
                unsafe { sys::h263cu_destroy(self.ctx) };
            }
            self.ctx = fresh;
            self.ctx_w = w.max(16);
            self.ctx_h = h.max(16);
        }
        // DecodedPicture: tight planes, chroma = ceil(w/2) x ceil(h/2) (picture.rs:39-58)
        let (cw, ch) = ((w as usize + 1) / 2, (h as usize + 1) / 2);
        let mut luma = vec![0u8; w as usize * h as usize];
        let mut chroma_b = vec![0u8; cw * ch];
        let mut chroma_r = vec![0u8; cw * ch];
        let rc = unsafe { sys::h263cu_read_yuv(self.ctx, 0, luma.as_mut_ptr(), chroma_b.as_mut_ptr(), chroma_r.as_mut_ptr()) };
        if rc != 0 {
            return Err(Error::from_code(rc));
        }
        // the full header (quantiser, type, TR) as the parser saw it
        let mut info = [0u32; 5];
        unsafe { sys::h263cu_stream_info(self.ctx, 0, &mut info[0], &mut info[1], &mut info[2], &mut info[3], &mut info[4]) };
        let mut header = header_of(&pic);
        header.quantizer = info[3] as u8;
        self.last = Some(DecodedPicture { header, width: w as usize, luma, chroma_b, chroma_r });
        reader.commit();
        Ok(())
    }
}

impl Drop for H263State {
    fn drop(&mut self) {
        unsafe {
            if !self.ctx.is_null() {
                sys::h263cu_destroy(self.ctx);
            }
            sys::h263cu_parser_destroy(self.parser);
        }
    }
}

fn header_of(pic: &sys::h263cu_pic) -> Picture {
    Picture {
        version: if pic.version == 0xFF { None } else { Some(pic.version) },
        temporal_reference: pic.temporal_reference,
        format: if pic.width != 0 { Some((pic.width, pic.height)) } else { None },
        picture_type: match pic.pic_type {
            sys::H263CU_PIC_I => PictureTypeCode::IFrame,
            sys::H263CU_PIC_P => PictureTypeCode::PFrame,
            sys::H263CU_PIC_DISPOSABLE_P => PictureTypeCode::DisposablePFrame,
            _ => PictureTypeCode::PbFrame,
        },
        quantizer: pic.pquant,
        use_deblocker: pic.flags & sys::H263CU_PICFLAG_DEBLOCK != 0,
    }
}

#[cfg(test)]
mod tests {
    use super::*;

    #[test]
    fn a_source_without_a_picture_is_reported_like_the_reference() {
        let state = H263State::new(DecoderOption::SORENSON_SPARK_BITSTREAM);
        let garbage = [0xFFu8; 8];
        let mut reader = H263Reader::from_source(&garbage[..]);
        assert!(matches!(state.parse_picture(&mut reader, None), Ok(None)));
        let mut state = state;
        let e = state.decode_next_picture(&mut reader).unwrap_err();
        assert!(matches!(e, Error::MiddleOfBitstream));
        assert!(state.get_last_picture().is_none());
    }

    #[test]
    fn eof_is_an_eof_error() {
        let mut state = H263State::new(DecoderOption::SORENSON_SPARK_BITSTREAM);
        let short = [0u8; 2];
        let mut reader = H263Reader::from_source(&short[..]);
        assert!(state.decode_next_picture(&mut reader).unwrap_err().is_eof_error());
    }
}
