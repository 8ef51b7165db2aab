//! SOURCE ONLY: `extern "C"` binding of include/zk_b200.h for the iammadab/zk crates, plus the few helpers every
//! caller needs (field id lookup, layout assertion, status -> `&'static str`, a per-thread context).
//! Never compiled in the build image (no rustc/cargo); the same ABI is exercised by zk_b200/_ffi.py, and
//! tests/test_abi_host.py checks that `ffi.rs` declares exactly the entry points of the header.
//! The drop-in modules with the reference's own type and method names live in ../zk-b200.
#![allow(non_camel_case_types)]
use ark_ff::PrimeField;
use core::any::TypeId;
use core::ffi::CStr;

mod ffi;
pub use ffi::*;

#[repr(C)]
pub struct zk_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct zk_table {
    _p: [u8; 0],
}
#[repr(C)]
pub struct zk_transcript {
    _p: [u8; 0],
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct zk_microbench {
    pub imad_wide_per_s: f64,
    pub imad_lo_per_s: f64,
    pub iadd3_per_s: f64,
    pub mixed_per_s: f64,
    pub fe_mul_per_s: f64,
    pub copy_gbs: f64,
    pub read_gbs: f64,
    pub sm_clock_mhz: f64,
    pub dfma_per_s: f64,
    pub fe_mul_fixed_per_s: f64,
}

// zk_status (include/zk_b200.h)
pub const ZK_OK: i32 = 0;
pub const ZK_ERR_EVAL_LEN: i32 = 1;
pub const ZK_ERR_EVALUATE_ARITY: i32 = 2;
pub const ZK_ERR_EMPTY_PRODUCT: i32 = 3;
pub const ZK_ERR_NVARS_MISMATCH: i32 = 4;
pub const ZK_ERR_PROOF_ROUNDS: i32 = 5;
pub const ZK_ERR_INITIAL_EVAL: i32 = 6;
pub const ZK_ERR_ROUND_CHECK: i32 = 7;
pub const ZK_VERIFY_FALSE: i32 = 8;
pub const ZK_ERR_NOT_POW2: i32 = 9;
pub const ZK_ERR_NO_ROOT: i32 = 10;
pub const ZK_ERR_VAR_RANGE: i32 = 11;
pub const ZK_ERR_INVALID_ARG: i32 = 12;
pub const ZK_ERR_UNSUPPORTED: i32 = 13;
pub const ZK_ERR_CUDA: i32 = 14;
pub const ZK_ERR_NCCL: i32 = 15;
pub const ZK_ERR_OOM: i32 = 16;

pub const ZK_BLS12_381_FR: i32 = 0;
pub const ZK_BLS12_377_FR: i32 = 1;

/// 0 for BLS12-381 Fr, 1 for BLS12-377 Fr, None for any other field (the reference's generic CPU code stays in charge).
pub fn field_id_of<F: PrimeField>() -> Option<i32> {
    if TypeId::of::<F>() == TypeId::of::<ark_bls12_381::Fr>() {
        Some(ZK_BLS12_381_FR)
    } else if TypeId::of::<F>() == TypeId::of::<ark_bls12_377::Fr>() {
        Some(ZK_BLS12_377_FR)
    } else {
        None
    }
}

/// ark-ff 0.5 `Fp256` = 4 LE u64 Montgomery limbs; checked once because the struct is not repr(C).
pub fn assert_layout() {
    use ark_bls12_381::Fr;
    assert_eq!(core::mem::size_of::<Fr>(), 32);
    let one = Fr::from(1u64);
    let limbs: [u64; 4] = unsafe { core::mem::transmute_copy(&one) };
    assert_eq!(limbs, [0x00000001fffffffe, 0x5884b7fa00034802, 0x998c4fefecbc4ff5, 0x1824b159acc5056f]);
}

/// The reference's literal `&'static str` for a non-zero status (the strings live in the library's static data).
pub fn status_to_err(status: i32) -> &'static str {
    unsafe { CStr::from_ptr(zk_status_string(status)) }.to_str().unwrap_or("zk_b200 error")
}

pub fn check(status: i32) -> Result<(), &'static str> {
    if status == ZK_OK {
        Ok(())
    } else {
        Err(status_to_err(status))
    }
}

thread_local! {
    // one context per thread: the library wants one host thread per zk_ctx (the reference is single-threaded)
    static CTX: *mut zk_ctx = {
        assert_layout();
        let mut c: *mut zk_ctx = core::ptr::null_mut();
        let st = unsafe { zk_ctx_create(0, &mut c) };
        assert_eq!(st, 0, "zk_b200: no CUDA device (there is no CPU fallback)");
        c
    };
}

/// This thread's single-GPU context (device 0), created on first use.
pub fn ctx() -> *mut zk_ctx {
    CTX.with(|c| *c)
}

/// `&[F]` as the `const uint64_t*` the library reads (4 limbs per element; see `assert_layout`).
pub fn as_limbs<F: PrimeField>(xs: &[F]) -> *const u64 {
    xs.as_ptr() as *const u64
}
pub fn as_limbs_mut<F: PrimeField>(xs: &mut [F]) -> *mut u64 {
    xs.as_mut_ptr() as *mut u64
}

/// Host-table form of the prover: `tables[k]` = `poly.polynomials[k].evaluation_slice()`.
/// Returns (round_polys, challenges) — sumcheck/src/prover.rs:15-30.
pub fn prove<F: PrimeField>(
    field_id: i32, tables: &[&[F]], degree: u32, sum: &F, absorb_initial_poly: bool,
) -> Result<(Vec<Vec<F>>, Vec<F>), &'static str> {
    if tables.is_empty() {
        return Err("cannot create product polynomial from empty polynomials");
    }
    let n = tables[0].len().trailing_zeros();
    let ptrs: Vec<*const u64> = tables.iter().map(|t| as_limbs(t)).collect();
    let np = degree as usize + 1;
    let mut rp = vec![F::zero(); n as usize * np];
    let mut ch = vec![F::zero(); n as usize];
    check(unsafe {
        zk_sumcheck_prove_host(
            ctx(), field_id, ptrs.as_ptr(), tables.len() as u32, n, degree, sum as *const F as *const u64,
            absorb_initial_poly as i32, as_limbs_mut(&mut rp), as_limbs_mut(&mut ch), core::ptr::null_mut(),
            core::ptr::null_mut(),
        )
    })?;
    Ok((rp.chunks(np).map(|c| c.to_vec()).collect(), ch))
}

/// fft/src/lib.rs:4-19
pub fn ntt<F: PrimeField>(field_id: i32, mut values: Vec<F>, inverse: bool) -> Vec<F> {
    let st = unsafe { zk_ntt_host(ctx(), field_id, as_limbs_mut(&mut values), values.len() as u64, inverse as i32) };
    match st {
        ZK_OK => values,
        ZK_ERR_NOT_POW2 => panic!("values must be a power of 2"),
        ZK_ERR_NO_ROOT => panic!("called `Option::unwrap()` on a `None` value"),
        _ => panic!("{}", status_to_err(st)),
    }
}
