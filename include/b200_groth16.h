/*
 * b200_groth16.h - C-ABI of the B200-native Groth16 prover hot path for MNT4753 / MNT6753.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. It is what the reference's
 * `B::` wrapper layer (libsnark/prover_reference_include/prover_reference_functions.hpp:5-162, implemented on the
 * CPU by libsnark/prover_reference_functions.cpp) binds to in this repo; each entry point cites the reference
 * interface it replaces. INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - field element: 96 bytes = 12 LE u64 limbs, Montgomery form x*2^768 mod p, canonical
 *     (libsnark/serialization.hpp:22-32).
 *   - G1 affine point: 192 B (x, y); G2 affine: 384 B (MNT4753, Fq2) / 576 B (MNT6753, Fq3); y == 0 encodes the
 *     point at infinity (serialization.hpp:43-67, 83-111).
 *   - projective point (X:Y:Z), O = (0:1:0): 3 coordinates of the same field, i.e. 288 B (G1), 576 B / 864 B (G2).
 *   - `d_` pointers are device memory of the current device, `h_` pointers are host memory.
 *   - every function returns 0 on success; on failure a negative code, and b200_last_error() describes it.
 *     There is no CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef B200_GROTH16_H
#define B200_GROTH16_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200_MNT4753 0
#define B200_MNT6753 1
#define B200_FE_BYTES 96

/* ---- runtime ------------------------------------------------------------------------------------------------ */
const char *b200_version(void);
const char *b200_last_error(void);
int b200_device_count(void);
int b200_set_device(int ordinal);
int b200_sync(void);
int b200_malloc(void **d_ptr, size_t bytes);
int b200_free(void *d_ptr);
int b200_host_alloc(void **h_ptr, size_t bytes); /* pinned */
int b200_host_free(void *h_ptr);
int b200_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes);
int b200_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes);
int b200_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes);
int b200_memset_zero(void *d_ptr, size_t bytes);

/* ---- Fr vectors (elements of the scalar field of `curve`) -------------------------------------------------- */
/* a[i] *= b[i]          replaces B::vector_Fr_muleq  (prover_reference_functions.cpp:170-180 / 485-495) */
int b200_fr_muleq(int curve, void *d_a, const void *d_b, size_t n);
/* a[i] -= b[i]          replaces B::vector_Fr_subeq  (prover_reference_functions.cpp:182-192 / 497-507) */
int b200_fr_subeq(int curve, void *d_a, const void *d_b, size_t n);

/* ---- evaluation domain (basic radix-2, libfqfft basic_radix2_domain.tcc:25-134) ------------------------------ */
typedef struct b200_domain b200_domain;
/* replaces B::get_evaluation_domain (prover_reference_functions.cpp:157-161); m must be a power of two <= 2^s */
int b200_domain_create(int curve, size_t m, b200_domain **out);
int b200_domain_destroy(b200_domain *dom);
size_t b200_domain_size(const b200_domain *dom); /* B::domain_get_m */
/* test hook: copy the first `count` entries of a precomputed table to the host.
 * which: 0 omega^i, 1 omega^-i, 2 g^i, 3 g^-i/m, 4 {1/m, 1/Z(g)} */
int b200_domain_table(const b200_domain *dom, int which, void *h_out, size_t count);
int b200_domain_fft(b200_domain *dom, void *d_a);       /* basic_radix2_domain::FFT       (:62-68)  */
int b200_domain_ifft(b200_domain *dom, void *d_a);      /* B::domain_iFFT                  (:70-82)  */
int b200_domain_coset_fft(b200_domain *dom, void *d_a); /* B::domain_cosetFFT, g = 17      (:84-89)  */
int b200_domain_icoset_fft(b200_domain *dom, void *d_a);/* B::domain_icosetFFT             (:91-96)  */
int b200_domain_divide_by_z_on_coset(b200_domain *dom, void *d_a); /* B::domain_divide_by_Z_on_coset (:125-134) */
/* whole witness map, libsnark/main.cpp:104-163 == cuda_prover_piecewise.cu:18-53. ca/cb/cc (m elements each) are
 * clobbered; d_out receives m+1 elements with out[m] = 0. */
int b200_compute_h(b200_domain *dom, void *d_ca, void *d_cb, void *d_cc, void *d_out);

/* ---- multi-scalar multiplication ---------------------------------------------------------------------------- */
/* sum_i scalars[i] * points[i]. scalars: n Fr elements (Montgomery, as on disk); points: n affine points (wire
 * format, infinity = y==0). Result: projective point written to HOST memory (288 / 576 / 864 B).
 * replaces B::multiexp_G1 / B::multiexp_G2 (prover_reference_functions.cpp:247-265 / 553-571), i.e.
 * libff::multi_exp_with_mixed_addition (multiexp.tcc:443-496). */
int b200_msm_g1(int curve, const void *d_scalars, const void *d_points, size_t n, void *h_out_proj);
int b200_msm_g2(int curve, const void *d_scalars, const void *d_points, size_t n, void *h_out_proj);
/* Fixed-base MSM context: pre-shifted base tables for ANY point set (what a proving key builds for its five queries,
 * without the key around it): create once per base set, run for every scalar vector over it. Used by the sharded size
 * sweeps (each rank owns the table of its point range). */
typedef struct b200_msm_ctx b200_msm_ctx;
int b200_msm_ctx_create(int curve, int group, const void *d_points, size_t n, b200_msm_ctx **out);
int b200_msm_ctx_run(b200_msm_ctx *ctx, const void *d_scalars, void *h_out_proj);
int b200_msm_ctx_destroy(b200_msm_ctx *ctx);
/* tuning hook: force the Pippenger window width (0 = automatic) */
int b200_msm_set_window(int c);
/* Bucket accumulation of every following MSM: 0 = XYZZ mixed additions, 1 = batched affine additions with
 * simultaneous inversion, 2 = automatic (default: batched affine where it measured faster - large G2/Fq2 MSMs).
 * Same results bit for bit; the environment variable B200_BATCH_AFFINE=0|1|2 sets the initial mode. */
int b200_msm_set_batch_affine(int on);

/* ---- O(1) group / field helpers on HOST buffers (serial tail of the prover) --------------------------------- */
int b200_g1_add(int curve, const void *h_p, const void *h_q, void *h_out);          /* B::G1_add   (:163-168) */
int b200_g2_add(int curve, const void *h_p, const void *h_q, void *h_out);
int b200_g1_scale(int curve, const void *h_fr, const void *h_p, void *h_out);       /* B::G1_scale (:150-155) */
int b200_g2_scale(int curve, const void *h_fr, const void *h_p, void *h_out);
int b200_g1_to_affine(int curve, const void *h_p, void *h_out_xy);                  /* write_g1 (serialization.hpp:43-54) */
int b200_g2_to_affine(int curve, const void *h_p, void *h_out_xy);                  /* write_g2 (serialization.hpp:56-67) */
int b200_g1_from_affine(int curve, const void *h_xy, void *h_out_proj);             /* read_g1  (serialization.hpp:83-91) */
int b200_g2_from_affine(int curve, const void *h_xy, void *h_out_proj);
/* test hook: 2^k * P (P affine, not O) through the Jacobian doubling the base-table builder uses -> affine */
int b200_host_jacobian_doublings(int curve, int group, const void *h_xy, int k, void *h_out_xy);
/* host field ops (tag 0 = modulus A, 1 = modulus B); op: 0 add 1 sub 2 mul 3 inv 4 from_mont 5 to_mont
 * 6 inv by the bitwise binary gcd (the device's routine, run on the host) 7 inv by the batched binary gcd (what op 3
 * uses on the host) 8 inv by Fermat's little theorem */
int b200_host_fp_op(int tag, int op, const void *h_a, const void *h_b, void *h_r);

/* ---- proving key resident on the device + whole-proof entry point ------------------------------------------ */
typedef struct b200_params b200_params;
/* h_image = byte image of a parameter file (libsnark/main.cpp:42-61): d, m, A[m+1], B1[m+1], B2[m+1], L[m-1], H[d].
 * replaces B::read_params (prover_reference_functions.cpp:291-345) */
int b200_params_from_host(int curve, const void *h_image, size_t bytes, b200_params **out);
/* Load a parameter file straight from disk (the reference's B::read_params takes a path too,
 * prover_reference_functions.cpp:291-345, and parses it element by element with fread, serialization.hpp:83-111: 8.0 s
 * for the 1.2 GB MNT4753 key). Here: one bulk read per 64 MB chunk into two pinned staging buffers,
 * each chunk's host->device copy running asynchronously under the read of the next one; the wire encoding IS the
 * device layout, so nothing is parsed. b200_params_load_ms: out3 = {file read, waiting for copies, total} of that load
 * (zeros for keys that were not loaded from a file). */
int b200_params_from_file(int curve, const char *path, b200_params **out);
int b200_params_load_ms(const b200_params *p, double *out3);
/* The same chunked pinned-read + asynchronous copy for any byte range of a file: d_dst[0, bytes) = file[offset, +bytes).
 * B::read_input binds to it (libsnark/main.cpp:63-83 reads the witness element by element). */
int b200_file_to_device(const char *path, size_t file_offset, void *d_dst, size_t bytes);
/* The evaluation domain of size d+1 that every key owns (built when the key is loaded, i.e. outside the reference's
 * timed region). Borrowed: valid until b200_params_destroy. B::get_evaluation_domain hands it out when the size matches,
 * so the driver's witness map does not rebuild the twiddle tables inside the timed region. */
b200_domain *b200_params_domain(const b200_params *p);
/* adopt caller-owned device arrays (synthetic keys built on the device) */
int b200_params_from_device(int curve, size_t d, size_t m, const void *d_A, const void *d_B1, const void *d_B2,
                            const void *d_L, const void *d_H, b200_params **out);
int b200_params_destroy(b200_params *p);
/* Pre-shifted base tables: for rank's slice of each query build table[j][i] = 2^(start_j) * P_i for the W windows of
 * the MSM (key-only preprocessing, 36-48x the size of the query, lives in the spare HBM). b200_prove* use them when
 * precomputation is enabled (default; B200_PRECOMPUTE=0 or b200_set_precompute(0) selects the table-free MSM) and
 * build them on first use if this was not called. Idempotent per (rank, world). */
int b200_params_precompute(b200_params *p, int rank, int world);
double b200_params_precompute_ms(const b200_params *p);
/* One throw-away proof on pseudo-random inputs: loads every kernel and sizes every workspace for this key, so that the
 * first real proof of a process does not pay first-use costs (B::read_params calls it after the tables are built). */
int b200_params_warmup(b200_params *p);
int b200_set_precompute(int on);
/* sum_i scalars[i] * query[i] over one whole query of the key (which: 0 A, 1 B1, 2 B2 (G2), 3 L, 4 H); uses the
 * pre-shifted base table when precomputation is enabled and n is the query's length. Result: projective, host memory.
 * B::multiexp_G1 / B::multiexp_G2 on vectors obtained from B::params_* bind to this. */
int b200_params_msm(b200_params *p, int which, const void *d_scalars, size_t n, void *h_out_proj);
/* The same MSM, asynchronously: returns as soon as the GPU work is enqueued (each outstanding call takes the next of
 * five workspaces / streams, so consecutive calls overlap on the device like the five MSMs of b200_prove);
 * b200_msm_wait blocks until h_out_proj - which must stay valid until then - holds the result, and frees the handle.
 * This is what lets the reference's UNMODIFIED driver (cuda_prover_piecewise.cu:71-81 issues its five multiexps back
 * to back and first looks at a result at :85-87) overlap them: B::multiexp_* return a pending point that is resolved
 * when B::G1_scale / G1_add / groth16_output_write first read it. */
typedef struct b200_msm_pending b200_msm_pending;
int b200_params_msm_async(b200_params *p, int which, const void *d_scalars, size_t n, void *h_out_proj,
                          b200_msm_pending **out);
int b200_msm_wait(b200_msm_pending *pending);
size_t b200_params_d(const b200_params *p);
size_t b200_params_m(const b200_params *p);
const void *b200_params_query(const b200_params *p, int which); /* 0 A, 1 B1, 2 B2, 3 L, 4 H (device pointers) */

typedef struct {
  double h2d_ms, compute_h_ms, msm_a_ms, msm_b1_ms, msm_b2_ms, msm_h_ms, msm_l_ms, tail_ms, total_ms;
} b200_prove_timings;
/* One proof: h_input = byte image of an input file (libsnark/main.cpp:63-83): w[m+1], ca, cb, cc [d+1 each], r.
 * h_out receives A (G1) | B (G2) | C (G1) in wire format (768 B MNT4753 / 960 B MNT6753), main.cpp:85-101,187-272.
 * The input is copied host->device inside the call; timings (optional) are wall-clock per phase.
 * range_lo/range_hi (in [0,1]) select the fraction of every MSM's point range this call sums - (0,1) for a whole
 * proof; for multi-GPU sharding each rank passes its slice and gets PARTIAL projective sums in h_partials
 * (5 points: A, B1, B2(G2), H, L) instead of a finished proof when h_out == NULL. */
int b200_prove(b200_params *p, const void *h_input, size_t input_bytes, void *h_out, size_t *out_bytes,
               b200_prove_timings *timings);
int b200_prove_partial(b200_params *p, const void *h_input, size_t input_bytes, int rank, int world,
                       void *h_partials, size_t *partial_bytes, b200_prove_timings *timings);
/* Uneven sharding: the caller owns the run [rank, rank_end) of `world` equal slices (b200_prove_partial is the run of
 * length one). A GPU that also proves another curve takes a shorter run; the partial sums of any set of runs that
 * tile [0, world) combine to the proof. b200_params_precompute_span builds the base tables for such a run. */
int b200_prove_partial_span(b200_params *p, const void *h_input, size_t input_bytes, int rank, int rank_end, int world,
                            void *h_partials, size_t *partial_bytes, b200_prove_timings *timings);
int b200_params_precompute_span(b200_params *p, int rank, int rank_end, int world);
/* The same with the witness map done elsewhere: d_h_coefficients (device, d+1 Fr elements = what b200_compute_h writes,
 * ready on the default stream) replaces this call's own compute_H; the ca / cb / cc part of the input image is then not
 * read. At N GPUs the three chains of compute_H (iFFT + cosetFFT of a, b, c: main.cpp:116-135) can run on three ranks
 * and the result be broadcast, instead of every rank repeating all seven transforms (bench.py: B200_BENCH_SPLIT_H). */
int b200_prove_partial_ext(b200_params *p, const void *h_input, size_t input_bytes, int rank, int rank_end, int world,
                           const void *d_h_coefficients, void *h_partials, size_t *partial_bytes, b200_prove_timings *timings);
/* As b200_prove_partial_ext, but the B1 slot of the partial sums holds r * (this rank's B1 sum), r read from the input
 * image: the 753 doublings of r * Bt1 (main.cpp:253) then run on every rank's host under its own GPU work instead of
 * serially on the combining rank. Such partials are combined with h_r_fr == NULL. */
int b200_prove_partial_scaled(b200_params *p, const void *h_input, size_t input_bytes, int rank, int rank_end, int world,
                              const void *d_h_coefficients, void *h_partials, size_t *partial_bytes,
                              b200_prove_timings *timings);
/* Per-query sharding: spans[2*q], spans[2*q+1] = the run [first, end) of `world` slices this call sums of query q
 * (0 A, 1 B1, 2 B2, 3 L, 4 H - the order of b200_params_query). An empty run (first == end) leaves O in that slot and, for
 * H, skips the witness map: a GPU can be given the whole H MSM (one compute_H instead of one per GPU) while others
 * split B2. Any set of calls whose runs tile [0, world) for every query combines to the proof. b1_scaled: as
 * b200_prove_partial_scaled. b200_params_precompute_queries builds the base tables for such a slicing. */
int b200_prove_partial_queries(b200_params *p, const void *h_input, size_t input_bytes, const int *spans, int world,
                               int b1_scaled, const void *d_h_coefficients, void *h_partials, size_t *partial_bytes,
                               b200_prove_timings *timings);
int b200_params_precompute_queries(b200_params *p, const int *spans, int world);
/* combine `world` partial results (rank-major, as produced by b200_prove_partial) into the final proof; h_r_fr == NULL:
 * the B1 slots are already multiplied by r (b200_prove_partial_scaled) */
int b200_prove_combine(int curve, const void *h_partials_all, int world, const void *h_r_fr, void *h_out,
                       size_t *out_bytes);

/* Several proofs in flight on the current device: job 0 runs on the calling thread, every further job on a
 * persistent worker thread with its own streams and workspaces, so a small proof (MNT6753, 2^15 constraints) runs
 * underneath a large one (MNT4753, 2^20). world <= 1: a finished proof in h_out (as b200_prove); world > 1: this
 * rank's partial sums (as b200_prove_partial; query_spans / b1_scaled select the _queries / _scaled variants). At most
 * 8 jobs; the keys must be distinct objects. Replaces running
 * the reference driver once per curve (cuda_prover_piecewise.cu:100-121). */
typedef struct {
  b200_params *key;
  const void *h_input;
  size_t input_bytes;
  void *h_out;
  size_t out_bytes; /* set by the call */
  int rank, world;
  int rank_end; /* world > 1: the job owns slices [rank, rank_end) of world; 0 means rank + 1 */
  const void *d_h_coefficients; /* world > 1, optional: see b200_prove_partial_ext */
  int b1_scaled; /* world > 1: nonzero = partial sums as b200_prove_partial_scaled writes them */
  const int *query_spans; /* world > 1, optional: 10 ints, per-query runs (b200_prove_partial_queries); overrides rank / rank_end */
  int status; /* set by the call: 0 or the job's error code */
  b200_prove_timings timings;
} b200_proof_job;
int b200_prove_batch(b200_proof_job *jobs, int count);

/* ---- complete Groth16 proof terms ------------------------------------------------------------
 * The challenge's prover stops at A = sum w_i A_i, B = sum w_i B_i (G2), C = H + L + r*B1 (main.cpp:227-253). A complete
 * r1cs_gg_ppzksnark proof (r1cs_gg_ppzksnark.tcc:457-470, and the `debug` sketch main.cpp:295-343) adds the key's
 * alpha / beta / delta elements and the second randomiser s:
 *     A' = alpha_g1 + A + r*delta_g1      B' = beta_g2 + B + s*delta_g2      C' = C + s*A' + r*beta_g1
 * (C already holds r*B1, which is what remains of r*B1' - r*s*delta_g1). O(1) host group operations on top of the
 * five MSMs. h_proof / h_out: A | B | C in wire format; h_extras: alpha_g1 | beta_g1 | delta_g1 (G1) | beta_g2 |
 * delta_g2 (G2) in wire format (the proving-key elements the challenge's parameter file leaves out). */
int b200_groth16_finalize(int curve, const void *h_proof, const void *h_r_fr, const void *h_s_fr, const void *h_extras,
                          void *h_out, size_t *out_bytes);
/* b200_prove followed by b200_groth16_finalize (r = the input image's last element, as in main.cpp:218) */
int b200_prove_full(b200_params *p, const void *h_input, size_t input_bytes, const void *h_s_fr, const void *h_extras,
                    void *h_out, size_t *out_bytes);

/* ---- key generation: fixed-base batch exponentiation ------------------------------------------
 * out[i] = scalars[i] * g for one base and n scalars - libff::batch_exp over get_window_table / windowed_exp
 * (multiexp.tcc:547-645), the inner loop of r1cs_gg_ppzksnark_generator (r1cs_gg_ppzksnark.tcc:289-342).
 * h_base_affine: g in affine wire format (HOST); d_scalars: n Fr elements (Montgomery, as everywhere); d_out_affine: n
 * affine wire-format points, i.e. exactly what a query of a parameter file holds. window: 0 = automatic, else the
 * window width in bits. ms3 (optional): milliseconds of {table build, exponentiation, conversion to affine}. */
int b200_batch_exp(int curve, int group, const void *h_base_affine, const void *d_scalars, size_t n, void *d_out_affine,
                   int window, double *ms3);

/* ---- test / bench hooks: element-wise application of the device primitives the kernels are built from ------- */
/* op: 0 add 1 sub 2 mul 3 sqr 4 from_mont 5 to_mont 6 inv ; tag: 0 = modulus A, 1 = modulus B */
int b200_dev_fp_op(int tag, int op, const void *d_a, const void *d_b, void *d_r, size_t n);
/* G2 coordinate-field op (Fq2 for MNT4753, Fq3 for MNT6753). op: 0 add 1 sub 2 mul 3 sqr 4 inv (fp2.tcc:128-142, fp3.tcc:125-143) */
int b200_dev_fqe_op(int curve, int op, const void *d_a, const void *d_b, void *d_r, size_t n);
/* group: 1 = G1, 2 = G2. op: 0 add(proj,proj) 1 dbl(proj) 2 mixed_add(proj, affine) 3 to_affine(proj)->affine */
int b200_dev_group_op(int curve, int group, int op, const void *d_p, const void *d_q, void *d_r, size_t n);
/* synthetic bases: out[i] = (first + i) * G  in affine wire format, G = the curve's G1/G2 generator */
int b200_gen_points(int curve, int group, void *d_out_affine, size_t n, uint64_t first);
/* IMAD roofline microbenchmark on every SM, runs of >= 50 ms each (shorter ones under-read: the clock is still
 * ramping). mac32_per_s[0] = independent IMAD.WIDE chains, [1] = the multiplier's (mad.lo.cc, madc.hi.cc) carry-chain
 * pattern, [2] = the NOMINAL rate 32 IMAD.WIDE / clk / SM x SMs x the device's maximum SM clock; ms[0..1] = the timed
 * runs, ms[2] = that clock in MHz; [3] = the generated Montgomery multiplication itself in a register-resident loop
 * (the production instruction mix at the accumulation kernels' occupancy) and its run time. Arrays of 4. */
int b200_imad_peak(double *mac32_per_s, double *ms);
/* the current mode (0, 1 or 2, see b200_msm_set_batch_affine) */
int b200_msm_get_batch_affine(void);
/* number of CUDA kernels this library has launched so far (bench.py: gpu_launches) */
unsigned long long b200_launch_count(void);
/* accumulated MSM phase times (ms) since the last reset: out10 = G1 {digits, sort, accumulate, reduce, host tail},
 * then the same five for G2 calls */
int b200_msm_phase_totals(double *out10, int reset);
/* window width c, number of windows W and task length T of the most recently prepared MSM */
int b200_msm_last_plan(int *out3);
/* time of the last MSM phases (ms): 0 digits, 1 sort, 2 accumulate, 3 reduce, 4 host tail */
int b200_msm_last_phase_ms(double *out5);
/* test hook (host only, no GPU): the key-load-time grouping of equal bases. members: indices grouped, representative
 * first; groups: (representative, first segment, number of segments) triples; *n_members / *n_groups: capacity in,
 * count out (u32 words); *merged: bases folded into a representative. Points at infinity are never grouped. */
int b200_host_equal_bases(const void *h_points, size_t n, size_t point_bytes, uint32_t *members, size_t *n_members,
                          uint32_t *groups, size_t *n_groups, size_t *merged);
/* diagnostics: begin != 0 marks t = 0 on the default stream; afterwards (begin == 0) out15[slot*3 + {0,1,2}] = ms at which
 * the MSM issued on stream `slot` (issue order of b200_prove: B2, A, B1, L, H) started accumulating, started its
 * bucket reduction and finished it. Single proof at a time. */
int b200_prove_timeline(int begin, double *out15);

#ifdef __cplusplus
}
#endif
#endif
