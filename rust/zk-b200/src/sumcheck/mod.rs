//! sumcheck/src/lib.rs:1-29
pub mod prover;
pub mod verifier;

use ark_ff::PrimeField;

/// The round polynomials (evaluations at 0..=MAX_VAR_DEGREE) and the prover's claimed sum (:8-11).
#[derive(Debug)]
pub struct SumcheckProof<F: PrimeField> {
    pub(crate) sum: F,
    pub(crate) round_polys: Vec<Vec<F>>,
}

/// What is left to check when the verifier skips the final oracle query: sum == initial_poly(challenges) (:17-20).
pub struct SubClaim<F: PrimeField> {
    pub(crate) sum: F,
    pub(crate) challenges: Vec<F>,
}

impl<F: PrimeField> SubClaim<F> {
    // Accessors are an addition: the reference keeps both fields crate-private and offers none (lib.rs:17-20), so a
    // caller outside the `sumcheck` crate cannot finish the check a sub-claim stands for.
    pub fn sum(&self) -> F {
        self.sum
    }
    pub fn challenges(&self) -> &[F] {
        &self.challenges
    }
}

impl<F: PrimeField> SumcheckProof<F> {
    /// Proof dump of SURVEY.md Appendix A.5 (the reference defines no serialiser): BE32(sum) || BE32 of every round
    /// evaluation; `zk_sumcheck_proof_dump` on the C side.
    pub fn to_bytes(&self) -> Vec<u8> {
        use zk_b200_sys as sys;
        let field = sys::field_id_of::<F>().expect(crate::UNSUPPORTED_FIELD);
        let flat: Vec<F> = self.round_polys.iter().flatten().copied().collect();
        let degree = self.round_polys.first().map_or(0, |r| r.len().saturating_sub(1)) as u32;
        let sum = [self.sum];
        let mut len = 0usize;
        let args = |out: *mut u8, cap: usize, len: &mut usize| unsafe {
            sys::zk_sumcheck_proof_dump(
                field, sys::as_limbs(&sum), sys::as_limbs(&flat), self.round_polys.len() as u32, degree, core::ptr::null(),
                core::ptr::null(), 0, out, cap, len, core::ptr::null_mut(),
            )
        };
        assert_eq!(args(core::ptr::null_mut(), 0, &mut len), sys::ZK_OK);
        let mut out = vec![0u8; len];
        assert_eq!(args(out.as_mut_ptr(), len, &mut len), sys::ZK_OK);
        out
    }
}
