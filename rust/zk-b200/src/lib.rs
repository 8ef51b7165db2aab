//! SOURCE ONLY (no rustc/cargo in the build image).
//!
//! The public surface of the reference's crates for the sumcheck hot path, same module paths, type names, method
//! names, argument meaning and `&'static str` errors, executed by the B200 library:
//!
//! | reference item                                                              | here                                   |
//! |-----------------------------------------------------------------------------|----------------------------------------|
//! | `polynomial::multilinear::evaluation_form::MultiLinearPolynomial`            | `polynomial::multilinear::evaluation_form` |
//! | `polynomial::multilinear::pairing_index::{index_pair, mask}`                  | `polynomial::multilinear::pairing_index`   |
//! | `polynomial::product_poly::ProductPoly`                                      | `polynomial::product_poly`             |
//! | `sumcheck::{SumcheckProof, SubClaim, prover::SumcheckProver, verifier::SumcheckVerifier}` | `sumcheck`                |
//! | `transcript::Transcript`                                                     | `transcript`                           |
//! | `fft::{fft, ifft}`                                                           | `fft`                                  |
//! | — (the GKR layer polynomial, beyond `ProductPoly`)           | `sum_of_products`                      |
//!
//! Only `ark_bls12_381::Fr` and `ark_bls12_377::Fr` are served (`zk_b200_sys::field_id_of`); for any other `F` every
//! call returns `Err(UNSUPPORTED_FIELD)` — the reference's generic CPU code is not reproduced here (INTEGRATION.md §3
//! shows the in-tree patch that keeps it as the fall-through for other fields).
pub mod fft;
pub mod polynomial;
pub mod sum_of_products;
pub mod sumcheck;
pub mod transcript;

pub const UNSUPPORTED_FIELD: &str = "zk_b200: only ark_bls12_381::Fr and ark_bls12_377::Fr run on the GPU path";

pub(crate) mod device {
    //! RAII wrapper of a device-resident table.
    use ark_ff::PrimeField;
    use zk_b200_sys as sys;

    pub struct DeviceTable(pub *mut sys::zk_table);

    impl DeviceTable {
        pub fn upload<F: PrimeField>(evaluations: &[F], n_vars: usize) -> Result<Self, &'static str> {
            let field = sys::field_id_of::<F>().ok_or(crate::UNSUPPORTED_FIELD)?;
            let mut t: *mut sys::zk_table = core::ptr::null_mut();
            sys::check(unsafe {
                sys::zk_table_upload(sys::ctx(), field, sys::as_limbs(evaluations), evaluations.len() as u64, n_vars as u32, &mut t)
            })?;
            Ok(DeviceTable(t))
        }
        pub fn download<F: PrimeField>(&self) -> Result<Vec<F>, &'static str> {
            let len = unsafe { sys::zk_table_local_len(self.0) } as usize;
            let mut out = vec![F::zero(); len];
            sys::check(unsafe { sys::zk_table_download(sys::ctx(), self.0, sys::as_limbs_mut(&mut out)) })?;
            Ok(out)
        }
        pub fn n_vars(&self) -> usize {
            unsafe { sys::zk_table_n_vars(self.0) as usize }
        }
    }

    impl Drop for DeviceTable {
        fn drop(&mut self) {
            unsafe { sys::zk_table_free(self.0) }
        }
    }
}
