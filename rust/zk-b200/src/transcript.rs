//! transcript/src/lib.rs:5-35 — the library's host Keccak-256 sponge (original Keccak padding, digest re-absorbed
//! after every challenge, `from_be_bytes_mod_order`), so that Rust callers and the prover inside the library share
//! one implementation.
use ark_ff::PrimeField;
use zk_b200_sys as sys;

pub struct Transcript {
    handle: *mut sys::zk_transcript,
}

impl Transcript {
    /// :10
    pub fn new() -> Self {
        Self { handle: unsafe { sys::zk_transcript_new() } }
    }

    /// :16
    pub fn append(&mut self, new_data: &[u8]) {
        unsafe { sys::zk_transcript_append(self.handle, new_data.as_ptr(), new_data.len()) }
    }

    /// :27
    pub fn sample_field_element<F: PrimeField>(&mut self) -> F {
        let field = sys::field_id_of::<F>().expect(crate::UNSUPPORTED_FIELD);
        let mut out = [F::zero()];
        let st = unsafe { sys::zk_transcript_sample_field_element(self.handle, field, sys::as_limbs_mut(&mut out)) };
        assert_eq!(st, sys::ZK_OK);
        out[0]
    }

    /// :32
    pub fn sample_n_field_elements<F: PrimeField>(&mut self, n: usize) -> Vec<F> {
        (0..n).map(|_| self.sample_field_element()).collect()
    }
}

impl Default for Transcript {
    fn default() -> Self {
        Self::new()
    }
}

impl Drop for Transcript {
    fn drop(&mut self) {
        unsafe { sys::zk_transcript_free(self.handle) }
    }
}
