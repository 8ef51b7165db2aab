//! sumcheck/src/prover.rs:9-74 over the B200 library (`zk_sumcheck_prove_host`: uploads, round loop with the host
//! Keccak transcript between kernel launches, proof back on the host).
use crate::polynomial::product_poly::ProductPoly;
use crate::sum_of_products::SumOfProductsPoly;
use crate::sumcheck::SumcheckProof;
use ark_ff::PrimeField;
use std::marker::PhantomData;
use zk_b200_sys as sys;

pub struct SumcheckProver<const MAX_VAR_DEGREE: u8, F: PrimeField> {
    _marker: PhantomData<F>,
}

impl<const MAX_VAR_DEGREE: u8, F: PrimeField> SumcheckProver<MAX_VAR_DEGREE, F> {
    /// :15-20 — absorbs `poly.to_bytes()` into the transcript first.  Takes `poly` by value like the reference.
    pub fn prove(poly: ProductPoly<F>, sum: F) -> Result<SumcheckProof<F>, &'static str> {
        Ok(Self::run(poly, sum, true)?.0)
    }

    /// :24-30 — the transcript starts empty (the verifier has no access to the initial polynomial).
    pub fn prove_partial(poly: ProductPoly<F>, sum: F) -> Result<(SumcheckProof<F>, Vec<F>), &'static str> {
        Self::run(poly, sum, false)
    }

    fn run(poly: ProductPoly<F>, sum: F, absorb: bool) -> Result<(SumcheckProof<F>, Vec<F>), &'static str> {
        let field = sys::field_id_of::<F>().ok_or(crate::UNSUPPORTED_FIELD)?;
        let tables: Vec<&[F]> = poly.polynomials().iter().map(|p| p.evaluation_slice()).collect();
        let (round_polys, challenges) = sys::prove::<F>(field, &tables, MAX_VAR_DEGREE as u32, &sum, absorb)?;
        Ok((SumcheckProof { sum, round_polys }, challenges))
    }

    /// `prove` over a sum of products: absorbs the tables' `to_bytes()` first.
    pub fn prove_sum_of_products(poly: SumOfProductsPoly<F>, sum: F) -> Result<SumcheckProof<F>, &'static str> {
        let (round_polys, _) = poly.prove(MAX_VAR_DEGREE as u32, &sum, true)?;
        Ok(SumcheckProof { sum, round_polys })
    }

    /// The same loop over a sum of products (`zk_sumcheck_prove_sop`).
    pub fn prove_partial_sum_of_products(
        poly: SumOfProductsPoly<F>,
        sum: F,
    ) -> Result<(SumcheckProof<F>, Vec<F>), &'static str> {
        let (round_polys, challenges) = poly.prove(MAX_VAR_DEGREE as u32, &sum, false)?;
        Ok((SumcheckProof { sum, round_polys }, challenges))
    }
}
