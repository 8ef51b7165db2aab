//! `polynomial` crate surface (polynomial/src/lib.rs:5-7) — the modules on the sumcheck path.
pub mod multilinear;
pub mod product_poly;
