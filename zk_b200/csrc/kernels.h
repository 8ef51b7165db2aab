// kernels.h — internal launch interface between the C-ABI host layer (api.cpp) and the sm_100a
// kernels (kernels_*.cu).  Not part of the public boundary; see include/zk_b200.h for that.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include <atomic>

#include "field.cuh"

namespace zk {

// Function attributes (the opt-in to more than 48 KB of dynamic shared memory) and occupancy are PER DEVICE, and one
// process may hold contexts on several devices (zk_ctx_create(device)), each driven by its own thread: lazily computed
// values are cached per device ordinal, in atomics.  `compute` returns a positive value (blocks per SM), or a negated
// cudaError_t, which is not cached.
constexpr int kMaxDevices = 64;
struct PerDeviceCache {
    std::atomic<int> v[kMaxDevices];
};
template <class Fn>
inline int per_device(PerDeviceCache& c, Fn&& compute) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return compute();
    int x = c.v[dev].load(std::memory_order_acquire);
    if (x <= 0) {
        x = compute();
        if (x > 0) c.v[dev].store(x, std::memory_order_release);
    }
    return x;
}

constexpr int kMaxFactors = 8;   // ProductPoly factor count supported on the device
constexpr int kMaxDegree = 15;   // MAX_VAR_DEGREE supported (round polynomial has kMaxDegree+1 evaluations)
constexpr int kMaxGridBlocks = 148 * 16;

struct TablePtrs {
    Fe* t[kMaxFactors];
};

// ---- the round-sum all-reduce fused into the reducing launch (sharded contexts) ------------------------
// Every rank owns a small mailbox in its HBM; the other ranks map it through CUDA IPC.  The last block of a reducing
// launch stores its D+1 partial sums straight into slot [parity][own rank] of EVERY rank's mailbox over NVLink (peer
// stores), then a sequence flag; it waits for the flags of all senders in its own mailbox, adds the partials in rank
// order (exact modular sums, the same order on every rank: bit-identical results everywhere) and publishes like a
// single-GPU launch.  No collective call and no second launch per round.  world == 0: not in use (single GPU, or the
// NCCL fallback when IPC is unavailable).
constexpr int kMaxRanks = 8;
struct MailboxSlot {
    Fe v[16];            // kMaxDegree + 1 partial sums
    unsigned flag;       // sequence number of the round these partials belong to (written last)
    unsigned pad[7];
};
struct MailboxArgs {
    int world;                       // 0 = disabled
    int rank;
    unsigned seq;                    // this exchange's sequence number (> 0, the same on every rank)
    MailboxSlot* mine;               // [2][kMaxRanks] slots in this GPU's memory
    MailboxSlot* peer[kMaxRanks];    // the same array of every rank (peer[rank] == mine), peer-mapped
};

// Per-context scratch used by the reducing kernels.
struct ReduceScratch {
    Fe* block_partials;      // [kMaxGridBlocks * (kMaxDegree+1)] device
    unsigned* ticket;        // device, zero between launches
    Fe* result_dev;          // [(kMaxDegree+1)] device copy of the last result
    Fe* result_host;         // [(kMaxDegree+1)] pinned+mapped host memory (device-visible alias below)
    Fe* result_host_devptr;  // device pointer aliasing result_host
    unsigned* flag_host;     // pinned+mapped completion flag: the last block stores `seq` after the results
    unsigned* flag_host_devptr;
    unsigned seq;            // value the next reducing launch publishes (0 = do not publish)
    uint64_t* lanes;         // when non-null the reducing launch also widens its result into u64 lanes (sharded all-reduce input)
    int num_sms;
    MailboxArgs mbox;        // world > 1: the launch all-reduces its result through the peer mailboxes itself
};

// ---- sumcheck hot path (kernels_sumcheck.cu) ---------------------------------------------------
// Round polynomial evaluations S(t) = sum_{j<half} prod_k [A_k[j] + t (A_k[j+half] - A_k[j])], t = 0..degree.
// (sumcheck/src/prover.rs:49-56).  Result (degree+1 elements, Montgomery, reduced) lands in
// scratch.result_dev and scratch.result_host.
cudaError_t launch_round_poly(int field, const TablePtrs& tabs, int m, int degree, uint64_t half,
                              const ReduceScratch& scratch, cudaStream_t stream, int* launches);
// The same sums over the sub-range of `count` pairs (tabs.t[k][j], tabs.t[k][j + hoff]), j < count — a slice of a round
// (the pointers are pre-offset to its first pair; hoff = the distance between the two entries of a pair = half the table).
// Round sums are additive over index ranges: zk_sumcheck_prove_host overlaps round 0 with the host-to-device copies this way.
// Fused paths only (has_fused_path).
cudaError_t launch_round_poly_range(int field, const TablePtrs& tabs, int m, int degree, uint64_t count, uint64_t hoff,
                                    const ReduceScratch& scratch, cudaStream_t stream, int* launches);
// In-place fold of every factor at r (prover.rs:64 -> evaluation_form.rs:40-80 with initial_var = 0):
// T[j] = T[j] - r (T[j] - T[j+half]), j < half.
cudaError_t launch_fold(int field, const TablePtrs& tabs, int m, uint64_t half, const Fe& r, cudaStream_t stream,
                        int* launches);
// Fused: fold the n_prev-entry tables at r (halving them in place), and in the same pass produce the
// next round's polynomial over the folded tables.  n_prev >= 4.
// `claim` (optional, fused path only): this rank's share of S_prev(r) — the value S(0) + S(1) of the round being
// computed must have — lets the kernel skip the t = 1 term and publish S(1) = claim - S(0) (same field element).
cudaError_t launch_fold_round_poly(int field, const TablePtrs& tabs, int m, int degree, uint64_t n_prev, const Fe& r,
                                   const ReduceScratch& scratch, cudaStream_t stream, int* launches,
                                   const Fe* claim = nullptr);
// sum_j prod_k A_k[j]  (the claim): result in scratch.result_dev[0] / result_host[0].
cudaError_t launch_product_sum(int field, const TablePtrs& tabs, int m, uint64_t n, const ReduceScratch& scratch,
                               cudaStream_t stream, int* launches);
// true if (m, degree) has a fully fused template instantiation
bool has_fused_path(int m, int degree);

// ---- sum of products (kernels_sop.cu) -------------------------------------------------------------
// P(x) = sum_t prod_{k in term t} A_k(x): the shape of a GKR layer polynomial add.(Wb + Wc) + mul.Wb.Wc, which the
// reference's ProductPoly (polynomial/src/product_poly.rs:4-10, a pure product) cannot express.
// A table may appear in several terms (and several times in one term); it is read — and folded — once per item.
constexpr int kMaxTerms = 8;
constexpr int kMaxVirtual = 4;
struct SopSpec {
    int n_tables;                          // distinct tables, <= kMaxFactors
    int n_terms;                           // <= kMaxTerms
    uint8_t len[kMaxTerms];                // factors of term t, 1..kMaxFactors
    uint8_t fac[kMaxTerms][kMaxFactors];   // table indices of term t (>= n_tables: a virtual table)
    // Launcher-made common factors (kernels_sop.cu: sop_group): virtual table n_tables + v = table virt_a[v] + table
    // virt_b[v], so that x.a + x.b is evaluated as x.(a + b).  Never set by the C ABI's callers.
    int n_virt;
    uint8_t virt_a[kMaxVirtual], virt_b[kMaxVirtual];
};
bool sop_degree_supported(int degree);  // 1..4, like the fused product path
// S(t) = sum_{j<half} sum_terms prod_k [A_k[j] + t (A_k[j+half] - A_k[j])], t = 0..degree  -> scratch.result_*
cudaError_t launch_sop_round_poly(int field, const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t half,
                                  const ReduceScratch& scratch, cudaStream_t stream, int* launches);
// fold every table of the n_prev-entry set at r in place (n_prev >= 4), and the next round's S(t) in the same pass.
// `claim` (optional): this rank's share of S_prev(r) = S(0) + S(1) of the round being computed; the kernel then skips
// the t = 1 products and publishes S(1) = claim - S(0).  Only valid when degree >= the longest term.
cudaError_t launch_sop_fold_round_poly(int field, const TablePtrs& tabs, const SopSpec& spec, int degree,
                                       uint64_t n_prev, const Fe& r, const ReduceScratch& scratch, cudaStream_t stream,
                                       int* launches, const Fe* claim = nullptr);

// The latency kernel of the small rounds (kernels_sumcheck.cu: round_small_kernel): the fused fold + round sums of a sum of
// products over at most 4 tables with 8 lanes per item; applies when the round has few items (q = n_prev / 4).
bool small_round_applies(int n_tables, int degree, uint64_t q);
cudaError_t launch_small_fold_round(int field, const TablePtrs& tabs, const SopSpec& spec, int degree, uint64_t q, const Fe& r,
                                    const ReduceScratch& scratch, cudaStream_t stream, int* launches, const Fe* claim);

// ---- MLE utilities (kernels_mle.cu) --------------------------------------------------------------
// General partial_evaluate step for variable `initial_var` of an nv-variable table: out[k] = fold of the
// pair (insert_bit(k,pos,0), |1<<pos), pos = nv-1-initial_var; out-of-place (evaluation_form.rs:54-72).
cudaError_t launch_fold_var(int field, const Fe* in, Fe* out, unsigned nv, unsigned initial_var, const Fe& a,
                            cudaStream_t stream, int* launches);
// out[j] = prod_k A_k[j]  (product_poly.rs:66-74)
cudaError_t launch_prod_reduce(int field, const TablePtrs& tabs, int m, uint64_t n, Fe* out, cudaStream_t stream,
                               int* launches);
// Montgomery -> 32-byte big-endian canonical (evaluation_form.rs:97-103)
cudaError_t launch_to_bytes(int field, const Fe* in, uint64_t n, uint8_t* out, cudaStream_t stream, int* launches);
// Montgomery <-> canonical little-endian limbs, in place
cudaError_t launch_convert(int field, Fe* data, uint64_t n, bool to_mont, cudaStream_t stream, int* launches);
// Synthetic table: entry j of the local shard is global index first + j*stride (SURVEY.md 8d generator)
cudaError_t launch_generate(int field, Fe* out, uint64_t count, uint64_t seed, uint64_t table_id, uint64_t first,
                            uint64_t stride, cudaStream_t stream, int* launches);
// out[j*G + q] = in[q*L + j]: re-interleave the all-gathered residual shards (q-major) into global order
// (`batch` tables stored back to back, in and out, in one launch)
cudaError_t launch_interleave(const Fe* in, Fe* out, uint64_t local_len, unsigned world, cudaStream_t stream,
                              int* launches, unsigned batch = 1);
// After the exact ncclSum all-reduce of the u64 lanes (one 32-bit limb per lane, written by the reducing
// launch): carry-propagate the lane sums, reduce mod p, publish result + completion flag.
cudaError_t launch_narrow(int field, const uint64_t* lanes, Fe* out_dev, Fe* out_host_devptr, int count,
                          unsigned* flag_host_devptr, unsigned seq, cudaStream_t stream, int* launches);

// ---- NTT (kernels_ntt.cu) --------------------------------------------------------------------
struct NttPlan;
cudaError_t ntt_plan_create(int field, unsigned log_n, bool inverse, cudaStream_t stream, NttPlan** out, int* launches);
void ntt_plan_destroy(NttPlan*);
bool ntt_plan_is(const NttPlan*, int field, unsigned log_n, bool inverse);
// natural order in -> natural order out.  `data` is clobbered; *result is where the transform ends up: `data`
// itself (small sizes) or the plan's N-element scratch buffer, which the caller may keep by handing the plan
// another N-element buffer in exchange (ntt_plan_adopt_scratch).
cudaError_t ntt_execute(NttPlan* plan, Fe* data, Fe** result, cudaStream_t stream, int* launches);
void ntt_plan_adopt_scratch(NttPlan* plan, Fe* buf);

// ---- multi-GPU NTT pieces (kernels_ntt_sharded.cu; the factorisation is described in ntt_sharded_kernels.cuh) ----
// out[i] = base^(i << shift), i < count
cudaError_t launch_pow_table(int field, Fe* out, uint64_t count, const Fe& base, unsigned shift, cudaStream_t stream,
                             int* launches);
// x[k] *= t_hi[k >> lo_bits] * t_lo[k & (2^lo_bits - 1)], k < m   (the inter-rank twiddles w_N^(q k))
cudaError_t launch_twiddle_mul(int field, Fe* x, uint64_t m, const Fe* t_lo, const Fe* t_hi, unsigned lo_bits,
                               cudaStream_t stream, int* launches);
// out[c][j] = (scale *) sum_q w_G^(q c) in[q][j], j < chunk; ranks in {2, 4, 8}; w_half[i] = w_G^i, i < ranks/2
// (the inverse root for the inverse transform); scale may be null
cudaError_t launch_gdft(int field, int ranks, const Fe* in, Fe* out, uint64_t chunk, const Fe* w_half, const Fe* scale,
                        cudaStream_t stream, int* launches);
// out[q * L + j] = in[j * G + q]: strided shards of a full table, rank-major (inverse of launch_interleave)
cudaError_t launch_deinterleave(const Fe* in, Fe* out, uint64_t local_len, unsigned world, cudaStream_t stream,
                                int* launches);

// ---- micro-benchmarks (microbench.cu) ------------------------------------------------------------
struct MicrobenchResult {
    double imad_wide_per_s;     // independent IMAD.WIDE.U32 per second, whole chip
    double imad_lo_per_s;       // IMAD (32-bit lo) per second
    double iadd3_per_s;         // IADD3 per second
    double mixed_per_s;         // IMAD.WIDE + IADD3 interleaved: instructions per second
    double fe_mul_per_s;        // standalone Montgomery multiplications per second (the field-mul ceiling)
    double copy_gbs;            // 256-bit streaming copy GB/s (read+write)
    double read_gbs;            // 256-bit streaming read GB/s
    double sm_clock_mhz;        // clock64-derived SM clock during the IMAD run
    double dfma_per_s;          // independent DFMA per second, whole chip
    double fe_mul_fixed_per_s;  // standalone fe_mul_fixed_f64 (FP64-pipe fixed-multiplier product) per second
};
cudaError_t run_microbench(int field, MicrobenchResult* out, cudaStream_t stream);

}  // namespace zk
