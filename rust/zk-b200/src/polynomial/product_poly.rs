//! polynomial/src/product_poly.rs:4-88 over the B200 library.
use crate::device::DeviceTable;
use crate::polynomial::multilinear::evaluation_form::MultiLinearPolynomial;
use ark_ff::PrimeField;
use zk_b200_sys as sys;

/// P(x) = A(x).B(x).C(x)  (:4-10)
#[derive(Clone, Debug, PartialEq)]
pub struct ProductPoly<F: PrimeField> {
    n_vars: usize,
    polynomials: Vec<MultiLinearPolynomial<F>>,
}

impl<F: PrimeField> ProductPoly<F> {
    /// :14-32
    pub fn new(polynomials: Vec<MultiLinearPolynomial<F>>) -> Result<Self, &'static str> {
        if polynomials.is_empty() {
            return Err("cannot create product polynomial from empty polynomials");
        }
        let n_vars = polynomials[0].n_vars();
        if polynomials.iter().any(|p| p.n_vars() != n_vars) {
            return Err("cannot create product polynomial from polynomial that don't share the same number of variables");
        }
        Ok(Self { n_vars, polynomials })
    }

    pub(crate) fn polynomials(&self) -> &[MultiLinearPolynomial<F>] {
        &self.polynomials
    }

    fn upload_all(&self) -> Result<Vec<DeviceTable>, &'static str> {
        self.polynomials.iter().map(|p| DeviceTable::upload(p.evaluation_slice(), p.n_vars())).collect()
    }

    /// :36-44
    pub fn evaluate(&self, assignments: &[F]) -> Result<F, &'static str> {
        if assignments.len() != self.n_vars {
            return Err("evaluate must assign to all variables");
        }
        let tables = self.upload_all()?;
        let handles: Vec<*const sys::zk_table> = tables.iter().map(|t| t.0 as *const _).collect();
        let mut out = [F::zero()];
        sys::check(unsafe {
            sys::zk_product_evaluate(
                sys::ctx(), handles.as_ptr(), handles.len() as u32, sys::as_limbs(assignments), assignments.len() as u32,
                sys::as_limbs_mut(&mut out),
            )
        })?;
        Ok(out[0])
    }

    /// :48-63
    pub fn partial_evaluate(&self, initial_var: usize, assignments: &[F]) -> Result<Self, &'static str> {
        let partial = self
            .polynomials
            .iter()
            .map(|p| p.partial_evaluate(initial_var, assignments))
            .collect::<Result<Vec<_>, _>>()?;
        Self::new(partial)
    }

    /// :66-74 — element-wise product of the factor tables.
    pub fn prod_reduce(&self) -> Vec<F> {
        let tables = self.upload_all().expect("upload");
        let handles: Vec<*const sys::zk_table> = tables.iter().map(|t| t.0 as *const _).collect();
        let mut out: *mut sys::zk_table = core::ptr::null_mut();
        let st = unsafe { sys::zk_product_prod_reduce(sys::ctx(), handles.as_ptr(), handles.len() as u32, &mut out) };
        assert_eq!(st, sys::ZK_OK, "{}", sys::status_to_err(st));
        DeviceTable(out).download().expect("download")
    }

    /// :77-83 — factor-major, index-minor.
    pub fn to_bytes(&self) -> Vec<u8> {
        self.polynomials.iter().flat_map(|p| p.to_bytes()).collect()
    }

    /// :86-88
    pub fn n_vars(&self) -> usize {
        self.n_vars
    }
}
