//! FFI binding of the B200 renderer (include/tuun_b200.h) and the lowering of a reference
//! `waveform::Waveform<M, S>` tree (src/lib/waveform.rs:23-100) to the ABI's op list.
//!
//! Not compiled in the repository that ships it (no Rust toolchain there); mirrors the header 1:1.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

use tuun::waveform::{Operator, Waveform};

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct TbNode {
    pub kind: u32,
    pub op: u32,
    pub a: i32,
    pub b: i32,
    pub c: i32,
    pub value: f32,
    pub param_slot: i32,
    pub list_off: u32,
    pub ff_count: u32,
    pub fb_count: u32,
    pub mark_id: u32,
    pub reserved: u32,
    pub fixed_off: u64,
    pub fixed_len: u64,
}

#[repr(C)]
pub struct TbProgram {
    _private: [u8; 0],
}

/// `tb_program_info` (include/tuun_b200.h): launch geometry and launch counters of a program.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct TbProgramInfo {
    pub n_nodes: u32,
    pub n_code_words: u32,
    pub n_slots: u32,
    pub state_words: u32,
    pub tile: u32,
    pub threads: u32,
    pub smem_bytes: u32,
    pub n_params: u32,
    pub kernel_launches: u64,
    pub lane_launches: u64,
    pub lane_smem_bytes: u32,
    pub lane_min_voices: u32,
    pub lane_capacity: u32,
    pub lane_fm_capacity: u32,
    pub split_passes: u32,
    pub split_segments: u32,
    pub split_seg_samples: u64,
    pub split_rounds: u64,
    pub sequence_parts: u32,
    pub split_fm_rounds: u32,
    pub sequence_renders: u64,
    pub lane_fm_ws_capacity: u32,
    pub reserved0: u32,
    pub fm_ws_launches: u64,
}

pub const TB_OUT_DEVICE: u32 = 1;
pub const TB_PARAMS_DEVICE: u32 = 2;
pub const TB_NO_VOICE_OUT: u32 = 4;

extern "C" {
    pub fn tb_program_create(
        nodes: *const TbNode, n_nodes: u32, lists: *const i32, n_lists: u32, fixed_pool: *const f32,
        fixed_len: u64, sample_rate: u32, device: c_int, out_program: *mut *mut TbProgram,
    ) -> c_int;
    pub fn tb_program_destroy(p: *mut TbProgram);
    pub fn tb_render(
        p: *mut TbProgram, params: *const f32, n_params: u32, n_voices: u32, n_samples: u64,
        out: *mut f32, out_stride: u64, out_len: *mut u64, flags: u32,
    ) -> c_int;
    pub fn tb_render_mix(
        p: *mut TbProgram, params: *const f32, n_params: u32, n_voices: u32, n_samples: u64,
        out: *mut f32, out_stride: u64, out_len: *mut u64, mix: *mut f32, flags: u32,
    ) -> c_int;
    pub fn tb_length(
        p: *mut TbProgram, params: *const f32, n_params: u32, n_voices: u32, max: u64, len: *mut u64,
        flags: u32,
    ) -> c_int;
    pub fn tb_reset(p: *mut TbProgram) -> c_int;
    pub fn tb_substitute(p: *mut TbProgram, mark_id: u32, value: f32, n_replaced: *mut u32) -> c_int;
    pub fn tb_segments_begin(
        p: *mut TbProgram, params: *const f32, n_params: u32, n_voices: u32, n_segments: u32, seg_samples: u64,
        flags: u32, n_passes: *mut u32,
    ) -> c_int;
    pub fn tb_segments_pass(
        p: *mut TbProgram, pass: u32, seg_lo: u32, seg_hi: u32, out: *mut f32, out_stride: u64, flags: u32,
    ) -> c_int;
    pub fn tb_segments_states(p: *mut TbProgram, states: *mut *mut core::ffi::c_void, bytes_per_segment: *mut u64) -> c_int;
    pub fn tb_segments_fix(p: *mut TbProgram, pass: u32) -> c_int;
    pub fn tb_segments_end(p: *mut TbProgram) -> c_int;
    pub fn tb_stream(p: *mut TbProgram) -> *mut c_void;
    pub fn tb_set_stream(p: *mut TbProgram, cuda_stream: *mut c_void) -> c_int;
    pub fn tb_seed_noise(p: *mut TbProgram, seed: u64, first_voice: u64) -> c_int;
    pub fn tb_program_get_info(p: *const TbProgram, info: *mut TbProgramInfo) -> c_int;
    pub fn tb_lane_kernel_times(p: *mut TbProgram, ms: *mut f32, cap: u32, n: *mut u32) -> c_int;
    pub fn tb_lower_check(
        nodes: *const TbNode, n_nodes: u32, lists: *const i32, n_lists: u32, fixed_len: u64,
        info: *mut TbProgramInfo,
    ) -> c_int;
    pub fn tb_last_error() -> *const c_char;
    pub fn tb_abi_version() -> u32;
}

/// The flat program: children before parents, root last.
#[derive(Default)]
pub struct OpList {
    pub nodes: Vec<TbNode>,
    pub lists: Vec<i32>,
    pub fixed_pool: Vec<f32>,
}

fn node(kind: u32) -> TbNode {
    TbNode { kind, a: -1, b: -1, c: -1, param_slot: -1, ..Default::default() }
}

impl OpList {
    fn push(&mut self, n: TbNode) -> i32 {
        self.nodes.push(n);
        (self.nodes.len() - 1) as i32
    }

    /// Post-order flattening; `mark` maps a MarkId to the u32 carried by TB_MARKED nodes.
    pub fn lower<M, S>(&mut self, w: &Waveform<M, S>, mark: &dyn Fn(&M) -> u32) -> i32 {
        use Waveform::*;
        match w {
            Const(v) => self.push(TbNode { value: *v, ..node(0) }),
            Time(_) => self.push(node(1)),
            Noise => self.push(node(2)),
            Fixed(samples, _) => {
                let off = self.fixed_pool.len() as u64;
                self.fixed_pool.extend_from_slice(samples);
                self.push(TbNode { fixed_off: off, fixed_len: samples.len() as u64, ..node(3) })
            }
            Fin { length, waveform } => {
                let a = self.lower(length, mark);
                let b = self.lower(waveform, mark);
                self.push(TbNode { a, b, ..node(4) })
            }
            Append(x, y, _) => {
                let a = self.lower(x, mark);
                let b = self.lower(y, mark);
                self.push(TbNode { a, b, ..node(5) })
            }
            Sine { frequency, phase, .. } => {
                let a = self.lower(frequency, mark);
                let b = self.lower(phase, mark);
                self.push(TbNode { a, b, ..node(6) })
            }
            Filter { waveform, feed_forward, feedback, .. } => {
                let a = self.lower(waveform, mark);
                let mut idx: Vec<i32> = feed_forward.iter().map(|c| self.lower(c, mark)).collect();
                idx.extend(feedback.iter().map(|c| self.lower(c, mark)));
                let off = self.lists.len() as u32;
                self.lists.extend(idx);
                self.push(TbNode {
                    a, list_off: off, ff_count: feed_forward.len() as u32, fb_count: feedback.len() as u32,
                    ..node(7)
                })
            }
            BinaryPointOp(op, x, y) => {
                let a = self.lower(x, mark);
                let b = self.lower(y, mark);
                let op = match op {
                    Operator::Add => 0, Operator::Subtract => 1, Operator::Multiply => 2,
                    Operator::Divide => 3, Operator::Merge => 4, Operator::Power => 5,
                };
                self.push(TbNode { op, a, b, ..node(8) })
            }
            Reset { trigger, waveform, .. } => {
                let a = self.lower(trigger, mark);
                let b = self.lower(waveform, mark);
                self.push(TbNode { a, b, ..node(9) })
            }
            Alt { trigger, positive_waveform, negative_waveform } => {
                let a = self.lower(trigger, mark);
                let b = self.lower(positive_waveform, mark);
                let c = self.lower(negative_waveform, mark);
                self.push(TbNode { a, b, c, ..node(10) })
            }
            Marked { id, waveform } => {
                let a = self.lower(waveform, mark);
                self.push(TbNode { a, mark_id: mark(id), ..node(11) })
            }
            Captured { waveform, .. } => {
                let a = self.lower(waveform, mark);
                self.push(TbNode { a, ..node(12) })
            }
        }
    }
}

#[derive(Debug)]
pub struct Error(pub c_int, pub String);

fn check(rc: c_int) -> Result<(), Error> {
    if rc == 0 {
        Ok(())
    } else {
        let msg = unsafe { CStr::from_ptr(tb_last_error()) }.to_string_lossy().into_owned();
        Err(Error(rc, msg))
    }
}

/// Drop-in for `generator::initialize_state` + `Generator::generate` on one waveform
/// (src/lib/generator.rs:39, :86): same call shape, state behind the handle.
pub struct B200Waveform {
    handle: *mut TbProgram,
}

impl B200Waveform {
    pub fn initialize_state<M, S>(w: &Waveform<M, S>, sample_rate: u32, mark: &dyn Fn(&M) -> u32) -> Result<Self, Error> {
        let mut ops = OpList::default();
        ops.lower(w, mark);
        let mut handle = std::ptr::null_mut();
        check(unsafe {
            tb_program_create(
                ops.nodes.as_ptr(), ops.nodes.len() as u32, ops.lists.as_ptr(), ops.lists.len() as u32,
                ops.fixed_pool.as_ptr(), ops.fixed_pool.len() as u64, sample_rate, -1, &mut handle,
            )
        })?;
        Ok(B200Waveform { handle })
    }

    /// `Generator::generate(&mut self, waveform, out) -> usize`
    pub fn generate(&mut self, out: &mut [f32]) -> Result<usize, Error> {
        if out.is_empty() {
            return Ok(0);
        }
        let mut len: u64 = 0;
        check(unsafe {
            tb_render(self.handle, std::ptr::null(), 0, 1, out.len() as u64, out.as_mut_ptr(), out.len() as u64, &mut len, 0)
        })?;
        Ok(len as usize)
    }

    /// `Generator::length(&mut self, waveform, max) -> usize`
    pub fn length(&mut self, max: usize) -> Result<usize, Error> {
        let mut len: u64 = 0;
        check(unsafe { tb_length(self.handle, std::ptr::null(), 0, 1, max as u64, &mut len, 0) })?;
        Ok(len as usize)
    }

    /// `waveform::set_state(w, State::Initial)`
    pub fn reset(&mut self) -> Result<(), Error> {
        check(unsafe { tb_reset(self.handle) })
    }

    /// `waveform::substitute(&mut w, &mark_id, &Const(value))` (waveform.rs:396): the number of `Marked`
    /// nodes whose contents were replaced.  The stream continues from the carried state.
    pub fn substitute_const(&mut self, mark_id: u32, value: f32) -> Result<u32, Error> {
        let mut n: u32 = 0;
        check(unsafe { tb_substitute(self.handle, mark_id, value, &mut n) })?;
        Ok(n)
    }
}

impl Drop for B200Waveform {
    fn drop(&mut self) {
        unsafe { tb_program_destroy(self.handle) }
    }
}
