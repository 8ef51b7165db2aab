//! P(x) = sum_t prod_{k in terms[t]} polynomials[k](x) — SURVEY.md 8f-4, beyond the reference's `ProductPoly`
//! (polynomial/src/product_poly.rs:4-10): the GKR layer polynomial add.(Wb + Wc) + mul.Wb.Wc is
//! `polynomials = [add, mul, Wb, Wc]`, `terms = [[0, 2], [0, 3], [1, 2, 3]]`.
use crate::device::DeviceTable;
use crate::polynomial::multilinear::evaluation_form::MultiLinearPolynomial;
use ark_ff::PrimeField;
use zk_b200_sys as sys;

#[derive(Clone, Debug, PartialEq)]
pub struct SumOfProductsPoly<F: PrimeField> {
    n_vars: usize,
    polynomials: Vec<MultiLinearPolynomial<F>>,
    term_len: Vec<u8>,
    term_factors: Vec<u8>,
}

impl<F: PrimeField> SumOfProductsPoly<F> {
    pub fn new(polynomials: Vec<MultiLinearPolynomial<F>>, terms: &[Vec<u8>]) -> Result<Self, &'static str> {
        if polynomials.is_empty() || terms.is_empty() || terms.iter().any(|t| t.is_empty()) {
            return Err("cannot create product polynomial from empty polynomials");
        }
        let n_vars = polynomials[0].n_vars();
        if polynomials.iter().any(|p| p.n_vars() != n_vars) {
            return Err("cannot create product polynomial from polynomial that don't share the same number of variables");
        }
        if terms.iter().flatten().any(|&k| k as usize >= polynomials.len()) {
            return Err("invalid argument");
        }
        Ok(Self {
            n_vars,
            polynomials,
            term_len: terms.iter().map(|t| t.len() as u8).collect(),
            term_factors: terms.iter().flatten().copied().collect(),
        })
    }

    pub fn n_vars(&self) -> usize {
        self.n_vars
    }

    fn upload_all(&self) -> Result<Vec<DeviceTable>, &'static str> {
        self.polynomials.iter().map(|p| DeviceTable::upload(p.evaluation_slice(), p.n_vars())).collect()
    }

    /// Sum of P over the boolean hypercube: the honest claim.
    pub fn sum(&self) -> Result<F, &'static str> {
        let tables = self.upload_all()?;
        let handles: Vec<*const sys::zk_table> = tables.iter().map(|t| t.0 as *const _).collect();
        let mut out = [F::zero()];
        sys::check(unsafe {
            sys::zk_sop_sum(
                sys::ctx(), handles.as_ptr(), handles.len() as u32, self.term_len.as_ptr(), self.term_factors.as_ptr(),
                self.term_len.len() as u32, sys::as_limbs_mut(&mut out),
            )
        })?;
        Ok(out[0])
    }

    /// P(assignments): the verifier's final check against `SubClaim.sum`.
    pub fn evaluate(&self, assignments: &[F]) -> Result<F, &'static str> {
        if assignments.len() != self.n_vars {
            return Err("evaluate must assign to all variables");
        }
        let tables = self.upload_all()?;
        let handles: Vec<*const sys::zk_table> = tables.iter().map(|t| t.0 as *const _).collect();
        let mut out = [F::zero()];
        sys::check(unsafe {
            sys::zk_sop_evaluate(
                sys::ctx(), handles.as_ptr(), handles.len() as u32, self.term_len.as_ptr(), self.term_factors.as_ptr(),
                self.term_len.len() as u32, sys::as_limbs(assignments), assignments.len() as u32, sys::as_limbs_mut(&mut out),
            )
        })?;
        Ok(out[0])
    }

    /// The tables' `to_bytes()` in order (what `prove` absorbs first).
    pub fn to_bytes(&self) -> Vec<u8> {
        self.polynomials.iter().flat_map(|p| p.to_bytes()).collect()
    }

    /// `SumcheckVerifier::verify` for P (sumcheck/src/verifier.rs:15-33): Ok(true) / Ok(false) / the reference's errors.
    pub(crate) fn verify(&self, sum: &F, round_polys: &[Vec<F>]) -> Result<bool, &'static str> {
        let tables = self.upload_all()?;
        let handles: Vec<*const sys::zk_table> = tables.iter().map(|t| t.0 as *const _).collect();
        let flat: Vec<F> = round_polys.iter().flatten().copied().collect();
        let degree = round_polys.first().map_or(0, |r| r.len().saturating_sub(1)) as u32;
        let st = unsafe {
            sys::zk_sumcheck_verify_sop(
                sys::ctx(), handles.as_ptr(), handles.len() as u32, self.term_len.as_ptr(), self.term_factors.as_ptr(),
                self.term_len.len() as u32, sum as *const F as *const u64, sys::as_limbs(&flat), round_polys.len() as u32, degree,
            )
        };
        match st {
            sys::ZK_OK => Ok(true),
            sys::ZK_VERIFY_FALSE => Ok(false),
            _ => Err(sys::status_to_err(st)),
        }
    }

    /// Round polynomials and challenges of the sumcheck over P (the reference's loop, sumcheck/src/prover.rs:33-73).
    pub(crate) fn prove(&self, degree: u32, sum: &F, absorb: bool) -> Result<(Vec<Vec<F>>, Vec<F>), &'static str> {
        let tables = self.upload_all()?; // consumed by the library, freed on drop
        let handles: Vec<*mut sys::zk_table> = tables.iter().map(|t| t.0).collect();
        let np = degree as usize + 1;
        let mut rp = vec![F::zero(); self.n_vars * np];
        let mut ch = vec![F::zero(); self.n_vars];
        sys::check(unsafe {
            sys::zk_sumcheck_prove_sop(
                sys::ctx(), handles.as_ptr(), handles.len() as u32, self.term_len.as_ptr(), self.term_factors.as_ptr(),
                self.term_len.len() as u32, degree, sum as *const F as *const u64, absorb as i32, sys::as_limbs_mut(&mut rp),
                sys::as_limbs_mut(&mut ch), core::ptr::null_mut(),
            )
        })?;
        Ok((rp.chunks(np).map(|c| c.to_vec()).collect(), ch))
    }
}
