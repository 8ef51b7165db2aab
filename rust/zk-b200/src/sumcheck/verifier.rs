//! sumcheck/src/verifier.rs:9-79 over the B200 library.
use crate::device::DeviceTable;
use crate::polynomial::product_poly::ProductPoly;
use crate::sumcheck::{SubClaim, SumcheckProof};
use ark_ff::PrimeField;
use std::marker::PhantomData;
use zk_b200_sys as sys;

pub struct SumcheckVerifier<F: PrimeField> {
    _marker: PhantomData<F>,
}

impl<F: PrimeField> SumcheckVerifier<F> {
    /// :15-33 — Ok(true) / Ok(false) / the reference's three Err strings.
    pub fn verify(poly: ProductPoly<F>, proof: SumcheckProof<F>) -> Result<bool, &'static str> {
        sys::field_id_of::<F>().ok_or(crate::UNSUPPORTED_FIELD)?;
        if proof.round_polys.len() != poly.n_vars() {
            return Err("invalid proof: require 1 round poly for each variable in poly");
        }
        let tables = poly
            .polynomials()
            .iter()
            .map(|p| DeviceTable::upload(p.evaluation_slice(), p.n_vars()))
            .collect::<Result<Vec<_>, _>>()?;
        let handles: Vec<*const sys::zk_table> = tables.iter().map(|t| t.0 as *const _).collect();
        let flat: Vec<F> = proof.round_polys.iter().flatten().copied().collect();
        let degree = proof.round_polys.first().map_or(0, |r| r.len().saturating_sub(1)) as u32;
        let sum = [proof.sum];
        let st = unsafe {
            sys::zk_sumcheck_verify(
                sys::ctx(), handles.as_ptr(), handles.len() as u32, sys::as_limbs(&sum), sys::as_limbs(&flat),
                proof.round_polys.len() as u32, degree,
            )
        };
        match st {
            sys::ZK_OK => Ok(true),
            sys::ZK_VERIFY_FALSE => Ok(false),
            _ => Err(sys::status_to_err(st)),
        }
    }

    /// The same for a sum of products proved with `SumcheckProver::prove_sum_of_products` (`zk_sumcheck_verify_sop`).
    pub fn verify_sum_of_products(
        poly: crate::sum_of_products::SumOfProductsPoly<F>,
        proof: SumcheckProof<F>,
    ) -> Result<bool, &'static str> {
        if proof.round_polys.len() != poly.n_vars() {
            return Err("invalid proof: require 1 round poly for each variable in poly");
        }
        poly.verify(&proof.sum, &proof.round_polys)
    }

    /// :38-41 — host only.
    pub fn verify_partial(proof: SumcheckProof<F>) -> Result<SubClaim<F>, &'static str> {
        let field = sys::field_id_of::<F>().ok_or(crate::UNSUPPORTED_FIELD)?;
        let flat: Vec<F> = proof.round_polys.iter().flatten().copied().collect();
        let degree = proof.round_polys.first().map_or(0, |r| r.len().saturating_sub(1)) as u32;
        let sum = [proof.sum];
        let mut sub = [F::zero()];
        let mut challenges = vec![F::zero(); proof.round_polys.len()];
        sys::check(unsafe {
            sys::zk_sumcheck_verify_partial(
                field, sys::as_limbs(&sum), sys::as_limbs(&flat), proof.round_polys.len() as u32, degree,
                sys::as_limbs_mut(&mut sub), sys::as_limbs_mut(&mut challenges),
            )
        })?;
        Ok(SubClaim { sum: sub[0], challenges })
    }
}
