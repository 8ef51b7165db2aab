//! fft/src/lib.rs:4-19 over the B200 NTT (`zk_ntt_host`): natural order in and out, omega_N = g^((p-1)/N), the inverse
//! scaled by N^-1.  `fft_internal(values, omega)` (:21) with a caller-chosen omega has no GPU counterpart: the
//! device twiddles are the canonical roots of unity, which is what `fft` / `ifft` pass.
use ark_ff::{FftField, PrimeField};
use zk_b200_sys as sys;

/// :4-8 — panics like the reference: "values must be a power of 2", or the `unwrap` on a missing root of unity.
pub fn fft<F: FftField + PrimeField>(coefficients: Vec<F>) -> Vec<F> {
    let field = sys::field_id_of::<F>().expect(crate::UNSUPPORTED_FIELD);
    sys::ntt(field, coefficients, false)
}

/// :11-19
pub fn ifft<F: FftField + PrimeField>(evaluations: Vec<F>) -> Vec<F> {
    let field = sys::field_id_of::<F>().expect(crate::UNSUPPORTED_FIELD);
    sys::ntt(field, evaluations, true)
}
