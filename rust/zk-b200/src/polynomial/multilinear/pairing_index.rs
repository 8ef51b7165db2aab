//! polynomial/src/multilinear/pairing_index.rs:2-26 — pure index arithmetic, kept on the host: the kernels compute
//! the same pairs in place (variable 0 is the most significant index bit).
//! `index_pair(3, 0)` yields (0,4), (1,5), (2,6), (3,7).

/// Pair k of variable `index` in an `n_vars` table: (k with a 0 inserted at bit n_vars-1-index, the same with a 1).
pub fn index_pair(n_vars: u8, index: u8) -> impl Iterator<Item = (usize, usize)> {
    let pos = n_vars - 1 - index; // debug builds panic on underflow, like the reference (:3,:6)
    let pairs = 1usize << (n_vars - 1);
    (0..pairs).map(move |k| {
        let low = k & mask(pos);
        let with_zero = ((k >> pos) << (pos + 1)) | low;
        (with_zero, with_zero | (1usize << pos))
    })
}

/// The `n` low bits set (:24-26).
pub fn mask(n: u8) -> usize {
    (1usize << n) - 1
}
