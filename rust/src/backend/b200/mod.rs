//! Safe wrapper over `ffi.rs`: RAII handles in the style of `MlxArray` (`src/backend/mlx/array.rs:15-29`),
//! status codes turned into `anyhow::Error` (the reference's module-level error type, `src/inference.rs:30,89`).
//! No `Tensor` crosses this boundary: the hot span of `transcribe()` (steps 2-8, `src/inference.rs:94-200`) is one call.
pub mod ffi;

use anyhow::{anyhow, Result};
use std::ffi::{CStr, CString};
use std::ptr;

fn check(status: i32) -> Result<()> {
    if status == ffi::ASRB_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(ffi::asrb_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow!("asr_b200 error {status}: {msg}"))
}

struct Ctx(*mut ffi::asrb_ctx);
impl Drop for Ctx { fn drop(&mut self) { unsafe { ffi::asrb_ctx_free(self.0); } } }
struct Model(*mut ffi::asrb_model);
impl Drop for Model { fn drop(&mut self) { unsafe { ffi::asrb_model_free(self.0); } } }
struct Session(*mut ffi::asrb_session);
impl Drop for Session { fn drop(&mut self) { unsafe { ffi::asrb_session_free(self.0); } } }

/// Device context + immutable weights + one session (KV cache, scratch, stream) for batch-1 `transcribe()` calls.
/// The session is created lazily and re-created only when a clip is longer than its capacity (the reference handles
/// arbitrary-length audio; a session sized for the worst case up front would pin ~1 GB of scratch per minute of audio).
/// Field order = drop order: session before model before context.
pub struct B200Engine {
    session: std::cell::RefCell<Option<(Session, usize)>>,   // (handle, capacity in samples)
    model: Model,
    _ctx: Ctx,
    max_new_tokens: usize,
}

// `transcribe(&self)` is `&self` in the reference and single-threaded (`src/inference.rs:89`); a session is not
// re-entrant, so the engine is Send but deliberately not Sync.
unsafe impl Send for B200Engine {}

impl B200Engine {
    /// Replaces the loaders of `AsrInference::load` (`src/inference.rs:39-74`): config.json + safetensors (single or
    /// sharded) are read by the library, bf16 stays bf16.
    pub fn load(model_dir: &str, device: i32, max_new_tokens: usize) -> Result<Self> {
        let mut ctx = ptr::null_mut();
        check(unsafe { ffi::asrb_init(device, &mut ctx) })?;
        let ctx = Ctx(ctx);
        let dir = CString::new(model_dir)?;
        let mut model = ptr::null_mut();
        check(unsafe { ffi::asrb_model_load(ctx.0, dir.as_ptr(), &mut model) })?;
        let model = Model(model);
        Ok(Self { session: std::cell::RefCell::new(None), model, _ctx: ctx, max_new_tokens })
    }

    /// Session with room for `n_samples`: capacity grows in 30 s steps, the old session is dropped first.
    fn session_for(&self, n_samples: usize) -> Result<*mut ffi::asrb_session> {
        let mut slot = self.session.borrow_mut();
        if let Some((s, cap)) = slot.as_ref() {
            if *cap >= n_samples { return Ok(s.0); }
        }
        *slot = None;
        let cap = ((n_samples + 479_999) / 480_000).max(1) * 480_000;
        let mut session = ptr::null_mut();
        check(unsafe { ffi::asrb_session_create(self.model.0, 1, cap as i64, 16, self.max_new_tokens as i32, &mut session) })?;
        *slot = Some((Session(session), cap));
        Ok(session)
    }

    pub fn dims(&self) -> Result<ffi::AsrbDims> {
        let mut d = ffi::AsrbDims::default();
        check(unsafe { ffi::asrb_model_dims(self.model.0, &mut d) })?;
        Ok(d)
    }

    /// Steps 2-8 of `transcribe()`: 16 kHz mono f32 samples (+ optional forced-language prompt ids,
    /// `src/inference.rs:246-250`) -> generated token ids, EOS excluded.
    pub fn transcribe_ids(&self, samples: &[f32], lang_ids: Option<&[i64]>) -> Result<Vec<i64>> {
        let mut ids = vec![0i32; self.max_new_tokens];
        let mut n = 0i32;
        let sp = [samples.as_ptr()];
        let sl = [samples.len() as i64];
        let lp = [lang_ids.map_or(ptr::null(), |v| v.as_ptr())];
        let ll = [lang_ids.map_or(0, |v| v.len() as i32)];
        let session = self.session_for(samples.len())?;
        check(unsafe {
            ffi::asrb_transcribe_ids(session, sp.as_ptr(), sl.as_ptr(), 1, lp.as_ptr(), ll.as_ptr(),
                                     self.max_new_tokens as i32, ids.as_mut_ptr(), &mut n)
        })?;
        Ok(ids[..n as usize].iter().map(|&t| t as i64).collect())
    }
}
