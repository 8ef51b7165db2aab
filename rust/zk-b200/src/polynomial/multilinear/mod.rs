//! polynomial/src/multilinear/mod.rs:3-4
pub mod evaluation_form;
pub mod pairing_index;
