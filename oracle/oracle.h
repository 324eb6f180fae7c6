/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the arkworks hot path (see oracle.c header).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load liboracle.so.  PARITY UNPINNED at the reference boundary (no golden vectors upstream). */
#ifndef ACCMSM_ORACLE_H
#define ACCMSM_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* field ids: 0 = Pallas base field Fp (= Vesta scalar), 1 = Pallas scalar field Fq (= Vesta base).
 * curve ids: 0 = Pallas (coords in Fp, scalars in Fq), 1 = Vesta (coords in Fq, scalars in Fp).
 * Field elements: 4 x u64 little-endian limbs.  "mont" = Montgomery image, R = 2^256 (the ark-ff
 * Fp256 memory image); "canon" = BigInteger256 canonical.
 * Affine points: 8 x u64 = x[4] || y[4] (Montgomery) + separate infinity byte. */

int  oracle_num_threads(void);
void oracle_set_num_threads(int n);

/* field vectors (n elements each, Montgomery in / Montgomery out unless stated) */
void oracle_fe_mul(int field, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
void oracle_fe_add(int field, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
void oracle_fe_sub(int field, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
void oracle_fe_inv(int field, const uint64_t *a, uint64_t *out, size_t n);
void oracle_fe_to_mont(int field, const uint64_t *canon, uint64_t *mont, size_t n);
void oracle_fe_from_mont(int field, const uint64_t *mont, uint64_t *canon, size_t n);

/* seeded synthetic inputs (SplitMix64) */
void oracle_gen_scalars(int field, uint64_t seed, size_t n, int montgomery, uint64_t *out);
void oracle_gen_points(int curve, uint64_t seed, size_t n, uint64_t *out_xy);

/* group */
int  oracle_on_curve(int curve, const uint64_t *xy);
void oracle_point_mul(int curve, const uint64_t *xy, uint8_t inf, const uint64_t *scalar_canon,
                      uint64_t *out_xy, uint8_t *out_inf);
void oracle_point_add(int curve, const uint64_t *a_xy, uint8_t a_inf, const uint64_t *b_xy,
                      uint8_t b_inf, uint64_t *out_xy, uint8_t *out_inf);

/* ark-ec 0.2.0 VariableBaseMSM::multi_scalar_mul restated, windows in parallel
 * (OpenMP) like ark's rayon path; result normalised to affine.  scalars are CANONICAL. */
void oracle_msm_ark(int curve, const uint64_t *bases_xy, const uint8_t *bases_inf, size_t n_bases,
                    const uint64_t *scalars_canon, size_t n_scalars, uint64_t *out_xy,
                    uint8_t *out_inf);
/* PedersenCommitment::commit / IpaPC::cm_commit: scalars are Fp256 Montgomery
 * images; into_repr() then MSM; optional randomizer * hiding_generator. */
void oracle_commit(int curve, const uint64_t *bases_xy, size_t n_bases, const uint64_t *elems_mont,
                   size_t n_elems, const uint64_t *hiding_xy, const uint64_t *randomizer_mont,
                   uint64_t *out_xy, uint8_t *out_inf);

/* SuccinctCheckPolynomial (App. A.3) */
void oracle_compute_coeffs(int field, const uint64_t *challenges_mont, int k, uint64_t *coeffs_mont);
void oracle_succinct_evaluate(int field, const uint64_t *challenges_mont, int k,
                              const uint64_t *z_mont, uint64_t *out_mont);
void oracle_poly_evaluate(int field, const uint64_t *coeffs_mont, size_t n, const uint64_t *z_mont,
                          uint64_t *out_mont);
/* IpaPC::check tail (App. A.2): final_key = cm_commit(key, compute_coeffs(xi)) ; returns 1 iff it
 * equals expected (affine compare).  out_xy/out_inf receive final_key. */
int  oracle_ipa_check_final_key(int curve, const uint64_t *key_xy, size_t n_key,
                                const uint64_t *challenges_mont, int k,
                                const uint64_t *expected_xy, uint8_t expected_inf,
                                uint64_t *out_xy, uint8_t *out_inf);
/* key folding of IpaPC::open (App. A.2): key_l += xi_round * key_r, k rounds -> 1 point */
void oracle_ipa_fold_key(int curve, const uint64_t *key_xy, size_t n_key,
                         const uint64_t *challenges_mont, int k, uint64_t *out_xy, uint8_t *out_inf);
/* IpaPC::open, one round at a time (App. A.2; src/ipa_pc_as/mod.rs:454-462): the host squeezes the round
 * challenge from (l, r) between the two calls.  State arrays hold n entries; after fold the first n/2 are live. */
void oracle_powers(int field, const uint64_t *z_mont, size_t n, uint64_t *out_mont);
void oracle_ipa_open_round_lr(int curve, const uint64_t *key_xy, const uint64_t *coeffs_mont,
                              const uint64_t *z_vec_mont, size_t n, const uint64_t *h_prime_xy,
                              uint64_t *l_xy, uint8_t *l_inf, uint64_t *r_xy, uint8_t *r_inf);
void oracle_ipa_open_fold(int curve, uint64_t *key_xy, uint64_t *coeffs_mont, uint64_t *z_vec_mont, size_t n,
                          const uint64_t *xi_mont, const uint64_t *xi_inv_mont);
/* group equation of IpaPC::succinct_check (src/ipa_pc_as/mod.rs:198-205) with transcript values supplied */
int  oracle_ipa_succinct_check(int curve, const uint64_t *comm_xy, uint8_t comm_inf, const uint64_t *z_mont,
                               const uint64_t *v_mont, const uint64_t *l_xy, const uint64_t *r_xy, int k,
                               const uint64_t *xi_mont, const uint64_t *h_prime_xy,
                               const uint64_t *final_key_xy, const uint64_t *c_mont);

/* combine_succinct_check_polynomials (src/ipa_pc_as/mod.rs:391-404) */
void oracle_combine_check_polys(int field, const uint64_t *challenges_mont /* m x k */, int m, int k,
                                const uint64_t *alphas_mont /* m */,
                                const uint64_t *random_poly_mont /* nullable */, size_t n_random,
                                uint64_t *out_mont /* 2^k (>= n_random) */);

/* hp_as (src/hp_as/mod.rs:278-349,482-512) */
void oracle_hadamard(int field, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
void oracle_scale(int field, const uint64_t *v, const uint64_t *c, uint64_t *out, size_t n);
/* vectors: m pointers with individual lengths (ragged allowed); out length = max(len, n_hiding) */
void oracle_combine_vectors(int field, const uint64_t *const *vecs, const size_t *lens, int m,
                            const uint64_t *challenges, const uint64_t *hiding, size_t n_hiding,
                            uint64_t *out, size_t out_len);
/* t-vectors: n inputs; a_vecs/b_vecs pointer arrays with lens; mu has >= n (+1 if hiding) entries;
 * out = (2n-1) x len row-major */
void oracle_tvecs(int field, const uint64_t *const *a_vecs, const size_t *a_lens,
                  const uint64_t *const *b_vecs, const size_t *b_lens, int n, const uint64_t *mu,
                  size_t len, const uint64_t *hiding_a, size_t n_ha, const uint64_t *hiding_b,
                  size_t n_hb, uint64_t *out);

/* r1cs_nark matrix_vec_mul (src/r1cs_nark_as/r1cs_nark/mod.rs:443-462) over CSR */
void oracle_csr_matvec(int field, const uint32_t *row_ptr, const uint32_t *cols,
                       const uint64_t *coeffs_mont, size_t n_rows, const uint64_t *input,
                       size_t n_input, const uint64_t *witness, size_t n_witness, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
