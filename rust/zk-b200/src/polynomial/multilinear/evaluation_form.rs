//! polynomial/src/multilinear/evaluation_form.rs:7-103 over the B200 library.
use crate::device::DeviceTable;
use ark_ff::PrimeField;
use zk_b200_sys as sys;

/// Dense evaluation-form multilinear polynomial (:7-10).  The evaluations stay available on the host
/// (`evaluation_slice` returns `&[F]` like the reference); every operation uploads them, runs on the GPU and, where
/// the reference returns a polynomial, downloads the result.  Callers that chain operations on large tables should
/// hold `zk_table` handles through `zk_b200_sys` instead.
#[derive(Clone, Debug, PartialEq)]
pub struct MultiLinearPolynomial<F: PrimeField> {
    n_vars: usize,
    evaluations: Vec<F>,
}

impl<F: PrimeField> MultiLinearPolynomial<F> {
    /// :15-27
    pub fn new(n_vars: usize, evaluations: Vec<F>) -> Result<Self, &'static str> {
        if n_vars >= usize::BITS as usize || evaluations.len() != (1usize << n_vars) {
            return Err("evaluation vec len should equal 2^n_vars");
        }
        Ok(Self { n_vars, evaluations })
    }

    /// :30
    pub fn n_vars(&self) -> usize {
        self.n_vars
    }

    /// :40-80 — binds the `assignments.len()` consecutive variables starting at `initial_var`.
    pub fn partial_evaluate(&self, initial_var: usize, assignments: &[F]) -> Result<Self, &'static str> {
        let table = DeviceTable::upload(&self.evaluations, self.n_vars)?;
        let mut out: *mut sys::zk_table = core::ptr::null_mut();
        sys::check(unsafe {
            sys::zk_mle_partial_evaluate(
                sys::ctx(), table.0, initial_var as u32, sys::as_limbs(assignments), assignments.len() as u32, &mut out,
            )
        })?;
        let out = DeviceTable(out);
        Ok(Self { n_vars: out.n_vars(), evaluations: out.download()? })
    }

    /// :83-89
    pub fn evaluate(&self, assignments: &[F]) -> Result<F, &'static str> {
        if assignments.len() != self.n_vars {
            return Err("evaluate must assign to all variables");
        }
        let table = DeviceTable::upload(&self.evaluations, self.n_vars)?;
        let mut out = [F::zero()];
        sys::check(unsafe {
            sys::zk_mle_evaluate(sys::ctx(), table.0, sys::as_limbs(assignments), assignments.len() as u32, sys::as_limbs_mut(&mut out))
        })?;
        Ok(out[0])
    }

    /// :92
    pub fn evaluation_slice(&self) -> &[F] {
        &self.evaluations
    }

    /// :97-103 — 32-byte big-endian canonical integers, in index order.
    pub fn to_bytes(&self) -> Vec<u8> {
        let mut out = vec![0u8; 32 * self.evaluations.len()];
        let field = sys::field_id_of::<F>().expect(crate::UNSUPPORTED_FIELD);
        // host-side conversion (the device kernel serves the prover's absorb pipeline, zk_sumcheck_prove)
        let st = unsafe { sys::zk_field_to_bytes_be(field, sys::as_limbs(&self.evaluations), self.evaluations.len(), out.as_mut_ptr()) };
        assert_eq!(st, sys::ZK_OK);
        out
    }
}
