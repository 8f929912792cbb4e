//! Drop-in batch entry points for looping over `elastic_elgamal`'s per-item `verify` calls.
//!
//! ```ignore
//! // before (examples/voting.rs:189-204):
//! for ballot in &ballots { let cts = ballot.verify(&params)?; for (t, c) in totals.iter_mut().zip(cts) { *t += *c; } }
//! // after:
//! let engine = Engine::new(0)?; engine.set_receiver(params.receiver())?;
//! let outcome = engine.verify_choices(&params, &ballots)?;   // outcome.verdicts[i], outcome.tally
//! ```
//!
//! Source only in the build image (no rustc); see INTEGRATION.md for how it maps onto the reference's types.

use elastic_elgamal::{
    app::{ChoiceParams, ChoiceVerificationError, EncryptedChoice, SingleChoice},
    group::Ristretto,
    Ciphertext, PublicKey, RingProof, VerificationError,
};
use elastic_elgamal_b200_sys as sys;
use std::{ffi::CStr, ptr};

/// API / CUDA failures (never a per-item outcome).
#[derive(Debug)]
pub struct EngineError {
    pub status: sys::eg_status,
    pub message: String,
}

/// One context per GPU; `!Sync` (use from one thread at a time), independent across instances.
pub struct Engine {
    ctx: *mut sys::eg_ctx,
}

unsafe impl Send for Engine {}

pub struct ChoiceOutcome {
    /// `Ok(())` or the error `EncryptedChoice::verify` would have returned, per ballot.
    pub verdicts: Vec<Result<(), ChoiceVerificationError>>,
    /// Sum of the verified ballots' ciphertexts, one per option.
    pub tally: Vec<Ciphertext<Ristretto>>,
}

impl Engine {
    pub fn new(device: i32) -> Result<Self, EngineError> {
        let mut ctx = ptr::null_mut();
        let status = unsafe { sys::eg_ctx_create(device, &mut ctx) };
        if status != sys::EG_SUCCESS {
            return Err(EngineError { status, message: "eg_ctx_create failed (no CUDA device?)".into() });
        }
        Ok(Self { ctx })
    }

    fn check(&self, status: sys::eg_status) -> Result<(), EngineError> {
        if status == sys::EG_SUCCESS {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(sys::eg_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(EngineError { status, message })
    }

    /// Mirrors `PublicKey::from_bytes` validation and builds the key's fixed-base table.
    pub fn set_receiver(&self, key: &PublicKey<Ristretto>) -> Result<(), EngineError> {
        self.check(unsafe { sys::eg_ctx_set_receiver(self.ctx, key.as_bytes().as_ptr()) })
    }

    /// `ballots.iter().map(|b| b.verify(params))` + the tally fold.
    pub fn verify_choices(
        &self,
        params: &ChoiceParams<Ristretto, SingleChoice>,
        ballots: &[EncryptedChoice<Ristretto, SingleChoice>],
    ) -> Result<ChoiceOutcome, EngineError> {
        let m = params.options_count();
        let n = ballots.len();
        let (mut choices, mut rings, mut sums) = (Vec::with_capacity(n * m * 64), Vec::with_capacity(n * (1 + 2 * m) * 32), Vec::with_capacity(n * 64));
        let mut wrong_len = vec![false; n];
        for (i, ballot) in ballots.iter().enumerate() {
            if ballot.len() != m {
                // OptionsLenMismatch is decided on the host (choice.rs:149-158); the slot is filled with a dummy
                wrong_len[i] = true;
                choices.resize(choices.len() + m * 64, 0);
                rings.resize(rings.len() + (1 + 2 * m) * 32, 0);
                sums.resize(sums.len() + 64, 0);
                continue;
            }
            for ct in ballot.choices_unchecked() {
                choices.extend_from_slice(&ct.to_bytes());
            }
            rings.extend_from_slice(&ballot.range_proof().to_bytes());
            sums.extend_from_slice(&ballot.sum_proof().to_bytes());
        }
        let mut verdicts = vec![0_u8; n];
        let mut tally = vec![0_u8; m * 64];
        self.check(unsafe {
            sys::eg_verify_choice_batch(self.ctx, n, m as u32, 1, choices.as_ptr(), rings.as_ptr(), sums.as_ptr(),
                                        verdicts.as_mut_ptr(), tally.as_mut_ptr())
        })?;
        let verdicts = verdicts
            .iter()
            .zip(&wrong_len)
            .zip(ballots)
            .map(|((&v, &wrong), ballot)| match (wrong, v) {
                (true, _) => Err(ChoiceVerificationError::OptionsLenMismatch { expected: m, actual: ballot.len() }),
                (_, sys::EG_V_OK) => Ok(()),
                (_, sys::EG_V_CHOICE_SUM) => Err(ChoiceVerificationError::Sum(VerificationError::ChallengeMismatch)),
                _ => Err(ChoiceVerificationError::Range(VerificationError::ChallengeMismatch)),
            })
            .collect();
        let tally = tally
            .chunks(64)
            .map(|c| {
                use elastic_elgamal::group::ElementOps;
                Ciphertext::from_elements(
                    Ristretto::deserialize_element(&c[..32]).expect("engine emits canonical encodings"),
                    Ristretto::deserialize_element(&c[32..]).expect("engine emits canonical encodings"),
                )
            })
            .collect();
        Ok(ChoiceOutcome { verdicts, tally })
    }

    /// `cts.iter().zip(proofs).map(|(ct, p)| key.verify_bool(*ct, p))`.
    pub fn verify_bools(&self, cts: &[Ciphertext<Ristretto>], proofs: &[RingProof<Ristretto>]) -> Result<Vec<Result<(), VerificationError>>, EngineError> {
        assert_eq!(cts.len(), proofs.len());
        let n = cts.len();
        let (mut c, mut p) = (Vec::with_capacity(n * 64), Vec::with_capacity(n * 96));
        let mut wrong_len = vec![None; n];
        for (i, (ct, proof)) in cts.iter().zip(proofs).enumerate() {
            c.extend_from_slice(&ct.to_bytes());
            let bytes = proof.to_bytes();
            if bytes.len() != 96 {
                wrong_len[i] = Some(bytes.len() / 32 - 1);      // check_lengths("items in all rings", ..), ring.rs:310-315
                p.resize(p.len() + 96, 0);
            } else {
                p.extend_from_slice(&bytes);
            }
        }
        let mut verdicts = vec![0_u8; n];
        self.check(unsafe { sys::eg_verify_bool_batch(self.ctx, n, c.as_ptr(), p.as_ptr(), verdicts.as_mut_ptr()) })?;
        Ok(verdicts
            .iter()
            .zip(wrong_len)
            .map(|(&v, wrong)| match (wrong, v) {
                (Some(actual), _) => Err(VerificationError::LenMismatch { collection: "items in all rings", expected: actual, actual: 2 }),
                (None, sys::EG_V_OK) => Ok(()),
                _ => Err(VerificationError::ChallengeMismatch),
            })
            .collect())
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { sys::eg_ctx_destroy(self.ctx) }
    }
}
