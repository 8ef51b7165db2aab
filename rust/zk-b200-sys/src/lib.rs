//! SOURCE ONLY: thin `extern "C"` binding of include/zk_b200.h for the iammadab/zk crates.
//! Never compiled in the build image (no rustc/cargo); the same ABI is exercised by zk_b200/_ffi.py.
//! Mirrors sumcheck/src/prover.rs:15-30 (`prove`, `prove_partial`) for F = ark_bls12_381::Fr / ark_bls12_377::Fr.
use ark_ff::PrimeField;
use core::any::TypeId;
use core::ffi::{c_char, CStr};

#[repr(C)]
pub struct zk_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct zk_table {
    _p: [u8; 0],
}

extern "C" {
    pub fn zk_ctx_create(device: i32, out: *mut *mut zk_ctx) -> i32;
    pub fn zk_ctx_destroy(ctx: *mut zk_ctx);
    pub fn zk_status_string(status: i32) -> *const c_char;
    pub fn zk_sumcheck_prove_host(
        ctx: *mut zk_ctx, field: i32, host_tables: *const *const u64, m: u32, n_vars: u32, degree: u32,
        sum: *const u64, absorb_initial_poly: i32, round_polys_out: *mut u64, challenges_out: *mut u64,
        final_evals_out: *mut u64, sum_out: *mut u64,
    ) -> i32;
    pub fn zk_ntt_host(ctx: *mut zk_ctx, field: i32, data: *mut u64, len: u64, inverse: i32) -> i32;
}

/// 0 for BLS12-381 Fr, 1 for BLS12-377 Fr, None for any other field (keep the CPU path).
pub fn field_id_of<F: PrimeField + 'static>() -> Option<i32> {
    if TypeId::of::<F>() == TypeId::of::<ark_bls12_381::Fr>() {
        Some(0)
    } else if TypeId::of::<F>() == TypeId::of::<ark_bls12_377::Fr>() {
        Some(1)
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

fn status_to_err(status: i32) -> &'static str {
    // the strings live in the library's static data for the life of the process
    unsafe { CStr::from_ptr(zk_status_string(status)) }.to_str().unwrap_or("zk_b200 error")
}

thread_local! {
    static CTX: *mut zk_ctx = {
        assert_layout();
        let mut c: *mut zk_ctx = core::ptr::null_mut();
        let st = unsafe { zk_ctx_create(0, &mut c) };
        assert_eq!(st, 0, "zk_b200: no CUDA device (there is no CPU fallback)");
        c
    };
}

/// `tables[k]` = `poly.polynomials[k].evaluation_slice()`.  Returns (round_polys, challenges).
pub fn prove<F: PrimeField>(
    field_id: i32, tables: &[&[F]], degree: u32, sum: &F, absorb_initial_poly: bool,
) -> Result<(Vec<Vec<F>>, Vec<F>), &'static str> {
    if tables.is_empty() {
        return Err("cannot create product polynomial from empty polynomials");
    }
    let n = tables[0].len().trailing_zeros();
    let ptrs: Vec<*const u64> = tables.iter().map(|t| t.as_ptr() as *const u64).collect();
    let np = degree as usize + 1;
    let mut rp = vec![F::zero(); n as usize * np];
    let mut ch = vec![F::zero(); n as usize];
    let st = CTX.with(|c| unsafe {
        zk_sumcheck_prove_host(
            *c, field_id, ptrs.as_ptr(), tables.len() as u32, n, degree, sum as *const F as *const u64,
            absorb_initial_poly as i32, rp.as_mut_ptr() as *mut u64, ch.as_mut_ptr() as *mut u64,
            core::ptr::null_mut(), core::ptr::null_mut(),
        )
    });
    if st != 0 {
        return Err(status_to_err(st));
    }
    Ok((rp.chunks(np).map(|c| c.to_vec()).collect(), ch))
}

/// fft/src/lib.rs:4-19
pub fn ntt<F: PrimeField>(field_id: i32, mut values: Vec<F>, inverse: bool) -> Vec<F> {
    let st = CTX.with(|c| unsafe { zk_ntt_host(*c, field_id, values.as_mut_ptr() as *mut u64, values.len() as u64, inverse as i32) });
    match st {
        0 => values,
        9 => panic!("values must be a power of 2"),
        10 => panic!("called `Option::unwrap()` on a `None` value"),
        _ => panic!("{}", status_to_err(st)),
    }
}
