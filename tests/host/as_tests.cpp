// C++ parity harness over accumulation_b200/host/ark_mirror.hpp -- the compiled-caller view of the drop-in.
// Scenarios follow the reference's own tests: commitments equal the expected group elements, a valid accumulator /
// proof is accepted by `decide` / `check`, and any corruption is rejected (src/lib.rs:334-395 test_template;
// src/hp_as/mod.rs:894-925; src/ipa_pc_as/mod.rs:820-848; src/r1cs_nark_as/r1cs_nark/mod.rs:335-419).
// Expected values come from the CPU oracle (oracle/oracle.h, test infrastructure).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include "../../accumulation_b200/host/ark_mirror.hpp"
#include "../../oracle/oracle.h"

using namespace accmsm_host;
static int failures = 0;
#define CHECK(name, cond) do { bool ok_ = (cond); std::printf("%s %s\n", ok_ ? "PASS" : "FAIL", name); if (!ok_) failures++; } while (0)

static std::vector<Fe> gen_scalars(int field, uint64_t seed, size_t n) {
    std::vector<Fe> v(n);
    if (n) oracle_gen_scalars(field, seed, n, 1, v[0].data());
    return v;
}
static std::vector<Affine> gen_points(int curve, uint64_t seed, size_t n) {
    std::vector<uint64_t> xy(n * 8);
    oracle_gen_points(curve, seed, n, xy.data());
    std::vector<Affine> out(n);
    for (size_t i = 0; i < n; i++) out[i] = affine_from(xy.data() + 8 * i, 0);
    return out;
}
static std::vector<uint64_t> flat(const std::vector<Affine> &p) {
    std::vector<uint64_t> xy(p.size() * 8);
    for (size_t i = 0; i < p.size(); i++) affine_to(p[i], xy.data() + 8 * i);
    return xy;
}
static Affine oracle_commit_(int curve, const std::vector<Affine> &gens, const std::vector<Fe> &elems, const Affine *h = nullptr, const Fe *r = nullptr) {
    auto xy = flat(gens); uint64_t hx[8]; if (h) affine_to(*h, hx);
    uint64_t out[8]; uint8_t inf = 0;
    oracle_commit(curve, xy.data(), gens.size(), elems.empty() ? nullptr : elems[0].data(), elems.size(), h ? hx : nullptr, r ? r->data() : nullptr, out, &inf);
    return affine_from(out, inf);
}

int main() {
    // ACCMSM_TEST_DEVICES="0,0" (or "0,1"): the same scenarios on a device-group ctx (accmsm_init_multi), keys sharded
    // by point range inside the library
    std::shared_ptr<Context> ctx;
    try {
        const char *devs = getenv("ACCMSM_TEST_DEVICES");
        if (devs && *devs) {
            std::vector<int> d;
            for (const char *p = devs; *p;) { d.push_back(atoi(p)); while (*p && *p != ',') p++; if (*p == ',') p++; }
            ctx = std::make_shared<Context>(d, 16);
            printf("device group of %d\n", accmsm_device_count(ctx->raw()));
        } else ctx = std::make_shared<Context>(0);
    }
    catch (const AccmsmError &e) { std::printf("NO_GPU %s\n", e.what()); return 3; }   // no CPU fallback: refuse to run

    for (int curve = 0; curve < 2; curve++) {
        const int sf = scalar_field(curve);
        const size_t L = 1 << 12;
        auto pts = gen_points(curve, 500 + curve, L + 1);
        std::vector<Affine> gens(pts.begin(), pts.begin() + L);
        Affine hgen = pts[L];
        CommitterKey ck(ctx, curve, gens, hgen);
        std::string tag = curve == 0 ? "pallas " : "vesta ";

        // --- PedersenCommitment::commit
        auto a = gen_scalars(sf, 1, L), b = gen_scalars(sf, 2, L);
        Fe r1 = gen_scalars(sf, 3, 1)[0], r2 = gen_scalars(sf, 4, 1)[0], r3 = gen_scalars(sf, 5, 1)[0];
        CHECK((tag + "commit(a)").c_str(), PedersenCommitment::commit(ck, a) == oracle_commit_(curve, gens, a));
        CHECK((tag + "commit(a, Some(r))").c_str(), PedersenCommitment::commit(ck, a, r1) == oracle_commit_(curve, gens, a, &hgen, &r1));
        auto longer = gen_scalars(sf, 6, L + 50);
        std::vector<Fe> trunc(longer.begin(), longer.begin() + L);
        CHECK((tag + "commit truncates to the key length").c_str(), PedersenCommitment::commit(ck, longer) == oracle_commit_(curve, gens, trunc));
        CHECK((tag + "commit(empty) is the identity").c_str(), PedersenCommitment::commit(ck, {}).infinity);
        auto batch = PedersenCommitment::commit_batch(ck, {a, b});
        CHECK((tag + "commit_batch").c_str(), batch[0] == oracle_commit_(curve, gens, a) && batch[1] == oracle_commit_(curve, gens, b));

        // --- ASForHadamardProducts::decide on a valid accumulator, then corrupted ones
        std::vector<Fe> prod(L);
        oracle_hadamard(sf, a[0].data(), b[0].data(), prod[0].data(), L);
        CHECK((tag + "compute_hp").c_str(), ASForHadamardProducts::compute_hp(*ctx, sf, a, b) == prod);
        ASForHadamardProducts::InputInstance inst{oracle_commit_(curve, gens, a, &hgen, &r1), oracle_commit_(curve, gens, b, &hgen, &r2),
                                                  oracle_commit_(curve, gens, prod, &hgen, &r3)};
        ASForHadamardProducts::InputWitness wit{a, b, ASForHadamardProducts::Randomness{r1, r2, r3}};
        CHECK((tag + "hp_as decide accepts a valid accumulator").c_str(), ASForHadamardProducts::decide(ck, inst, wit));
        auto bad_wit = wit; bad_wit.a_vec[7][0] ^= 1;
        CHECK((tag + "hp_as decide rejects a corrupted witness").c_str(), !ASForHadamardProducts::decide(ck, inst, bad_wit));
        auto bad_inst = inst; bad_inst.comm_3 = inst.comm_1;
        CHECK((tag + "hp_as decide rejects a corrupted instance").c_str(), !ASForHadamardProducts::decide(ck, bad_inst, wit));
        auto bad_rand = wit; bad_rand.randomness->rand_2 = r3;
        CHECK((tag + "hp_as decide rejects wrong randomness").c_str(), !ASForHadamardProducts::decide(ck, inst, bad_rand));
        ASForHadamardProducts::InputInstance inst_nozk{oracle_commit_(curve, gens, a), oracle_commit_(curve, gens, b), oracle_commit_(curve, gens, prod)};
        CHECK((tag + "hp_as decide without zk").c_str(), ASForHadamardProducts::decide(ck, inst_nozk, {a, b, std::nullopt}));

        // --- combine_vectors / scale_vector (ragged)
        std::vector<Fe> a_short(a.begin(), a.begin() + L - 9), ch = gen_scalars(sf, 7, 2), hid = gen_scalars(sf, 8, 5), exp(L);
        { const uint64_t *vp[2] = {a_short[0].data(), b[0].data()}; size_t ln[2] = {a_short.size(), b.size()};
          oracle_combine_vectors(sf, vp, ln, 2, ch[0].data(), hid[0].data(), hid.size(), exp[0].data(), L); }
        CHECK((tag + "combine_vectors (ragged, hiding)").c_str(), ASForHadamardProducts::combine_vectors(*ctx, sf, {&a_short, &b}, ch, &hid) == exp);
        oracle_scale(sf, a[0].data(), ch[1].data(), exp[0].data(), L);
        CHECK((tag + "scale_vector").c_str(), ASForHadamardProducts::scale_vector(*ctx, sf, a, ch[1]) == exp);

        // --- IpaPC: open (k = 8), succinct-check equation, decider's final-key check
        const int k = 8; const size_t D = size_t(1) << k;
        std::vector<Affine> key(gens.begin(), gens.begin() + D);
        CommitterKey ipa_ck(ctx, curve, key, hgen);
        auto coeffs = gen_scalars(sf, 9, D);
        Fe z = gen_scalars(sf, 10, 1)[0];
        auto squeeze = [&](const Affine &l, const Affine &r) {
            Fe xi = gen_scalars(sf, l.x[0] ^ (r.x[1] << 1) ^ 0x5eed, 1)[0], inv;
            oracle_fe_inv(sf, xi.data(), inv.data(), 1);
            return std::make_pair(xi, inv);
        };
        IpaProofCore proof = InnerProductArgPC::open(ipa_ck, coeffs, k, z, hgen, squeeze);
        Affine comm = oracle_commit_(curve, key, coeffs);
        Fe v; oracle_poly_evaluate(sf, coeffs[0].data(), D, z.data(), v.data());
        auto lx = flat(proof.l_vec), rx = flat(proof.r_vec), cx = flat({comm}), hx = flat({hgen}), fx = flat({proof.final_comm_key});
        CHECK((tag + "ipa open: proof satisfies succinct_check").c_str(),
              oracle_ipa_succinct_check(curve, cx.data(), 0, z.data(), v.data(), lx.data(), rx.data(), k, proof.round_challenges[0].data(), hx.data(),
                                        fx.data(), proof.c.data()) == 1);
        {   // same opening with h' given as xi_0 * (hiding generator of the key)
            Fe xi0 = gen_scalars(sf, 15, 1)[0], xi0c;
            oracle_fe_from_mont(sf, xi0.data(), xi0c.data(), 1);
            uint64_t hx2[8], hpo[8]; uint8_t hpi = 0; affine_to(hgen, hx2);
            oracle_point_mul(curve, hx2, 0, xi0c.data(), hpo, &hpi);
            Affine hprime = affine_from(hpo, hpi);
            IpaProofCore p1 = InnerProductArgPC::open(ipa_ck, coeffs, k, z, hprime, squeeze);
            IpaProofCore p2 = InnerProductArgPC::open(ipa_ck, coeffs, k, z, hprime, squeeze, xi0);
            CHECK((tag + "ipa open: h' as a point == h' as (hiding index, xi_0)").c_str(),
                  p1.l_vec == p2.l_vec && p1.r_vec == p2.r_vec && p1.final_comm_key == p2.final_comm_key && p1.c == p2.c);
        }
        {   // one library call per round (accmsm_ipa_open_fold_round, the challenge's inverse taken inside the library): same proof
            IpaProofCore p3 = InnerProductArgPC::open_one_call_per_round(ipa_ck, coeffs, k, z, hgen, [&](const Affine &l, const Affine &r) { return squeeze(l, r).first; });
            CHECK((tag + "ipa open: one call per round (fold_round) gives the same proof").c_str(),
                  p3.l_vec == proof.l_vec && p3.r_vec == proof.r_vec && p3.final_comm_key == proof.final_comm_key && p3.c == proof.c &&
                  p3.round_challenges == proof.round_challenges);
        }
        {   // VariableBaseMSM::multi_scalar_mul on unregistered bases (2k + 3 = 19 terms), alone and three of them in one batched call
            const size_t nt = 2 * k + 3;
            std::vector<std::vector<Affine>> bb(3); std::vector<std::vector<Fe>> ss(3); std::vector<Affine> expd(3);
            for (int j = 0; j < 3; j++) {
                bb[j].assign(gens.begin() + 100 * j, gens.begin() + 100 * j + nt);
                auto m = gen_scalars(sf, 700 + j, nt); ss[j].resize(nt);
                oracle_fe_from_mont(sf, m[0].data(), ss[j][0].data(), nt);
                if (j == 1) { ss[j][0] = Fe{0, 0, 0, 0}; ss[j][1] = Fe{1, 0, 0, 0}; bb[j][nt - 1] = bb[j][0]; }      // zero, one, a repeated base
                auto xy = flat(bb[j]); uint64_t o[8]; uint8_t oi = 0;
                oracle_msm_ark(curve, xy.data(), nullptr, nt, ss[j][0].data(), nt, o, &oi);
                expd[j] = affine_from(o, oi);
            }
            CHECK((tag + "multi_scalar_mul on unregistered bases").c_str(), VariableBaseMSM::multi_scalar_mul(*ctx, curve, bb[1], ss[1]) == expd[1]);
            CHECK((tag + "three one-shot MSMs in one batched call").c_str(), VariableBaseMSM::multi_scalar_mul_batch(*ctx, curve, bb, ss) == expd);
        }
        {   // the materialised folded key (accmsm_set_ipa_fold) changes nothing in the proof: forced every 2 rounds here
            ctx->check(accmsm_set_ipa_fold(ctx->raw(), 2, 2), "set_ipa_fold");
            IpaProofCore pf = InnerProductArgPC::open(ipa_ck, coeffs, k, z, hgen, squeeze);
            ctx->check(accmsm_set_ipa_fold(ctx->raw(), 5, 11), "set_ipa_fold");
            CHECK((tag + "ipa open: same proof with the folded key materialised every 2 rounds").c_str(),
                  pf.l_vec == proof.l_vec && pf.r_vec == proof.r_vec && pf.final_comm_key == proof.final_comm_key && pf.c == proof.c);
        }
        SuccinctCheckPolynomial h{proof.round_challenges};
        CHECK((tag + "ipa check: final_comm_key == cm_commit(key, h.compute_coeffs())").c_str(), InnerProductArgPC::check_final_key(ipa_ck, h, proof.final_comm_key));
        Affine bad_key = proof.final_comm_key; bad_key.y[2] ^= 4;
        CHECK((tag + "ipa check rejects a corrupted final_comm_key").c_str(), !InnerProductArgPC::check_final_key(ipa_ck, h, bad_key));
        auto hc = h.compute_coeffs(*ctx, sf);
        std::vector<Fe> hc_exp(D); oracle_compute_coeffs(sf, h.challenges[0].data(), k, hc_exp[0].data());
        CHECK((tag + "compute_coeffs").c_str(), hc == hc_exp);
        CHECK((tag + "cm_commit(key, coeffs) == final_comm_key").c_str(), InnerProductArgPC::cm_commit(ipa_ck, hc) == proof.final_comm_key);

        // --- r1cs_nark: Matrix<F> = Vec<Vec<(F, usize)>> as in examples/scaling-nark.rs (one empty row, unit coefficients)
        const size_t M = L;
        Fe one = gen_scalars(sf, 11, 1)[0]; { uint64_t c1[4] = {1, 0, 0, 0}; oracle_fe_to_mont(sf, c1, one.data(), 1); }
        auto rnd = gen_scalars(sf, 12, 3 * M);
        r1cs_nark::Matrix A(M), B(M), C(M);
        for (size_t i = 0; i + 1 < M; i++) {
            A[i] = {{one, 6}, {rnd[3 * i], i % 50}};
            B[i] = {{one, 7}};
            C[i] = {{rnd[3 * i + 1], 8 + i % (M - 10)}, {one, 3}, {rnd[3 * i + 2], (i * 7) % M}};
        }
        auto input = gen_scalars(sf, 13, 6), witness = gen_scalars(sf, 14, M - 6);
        auto mv = r1cs_nark::matrix_vec_mul(*ctx, sf, {&A, &B, &C}, input, witness);
        bool mv_ok = true;
        const r1cs_nark::Matrix *mats[3] = {&A, &B, &C};
        std::vector<std::vector<Fe>> mv_exp(3, std::vector<Fe>(M));
        for (int m = 0; m < 3; m++) {
            auto csr = r1cs_nark::to_csr(*mats[m]);
            oracle_csr_matvec(sf, csr.row_ptr.data(), csr.cols.data(), csr.coeffs.data(), M, input[0].data(), input.size(), witness[0].data(), witness.size(), mv_exp[m][0].data());
            mv_ok = mv_ok && mv[m] == mv_exp[m];
        }
        CHECK((tag + "matrix_vec_mul (A, B, C)").c_str(), mv_ok);
        r1cs_nark::IndexMatrices index(ck, A, B, C);
        std::array<Fe, 3> blinders{r1, r2, r3};
        auto fused = index.matvec_commit(input, witness, blinders);
        bool fc_ok = true;
        for (int m = 0; m < 3; m++) fc_ok = fc_ok && fused.first[m] == mv_exp[m] && fused.second[m] == oracle_commit_(curve, gens, mv_exp[m], &hgen, &blinders[m]);
        CHECK((tag + "nark: mat-vec + commitments on the device").c_str(), fc_ok);
    }
    // threading: every export is blocking and the ctx serialises its calls, so callers on several host
    // threads (rayon workers in the reference's deployment) may share one context
    {
        const size_t L = 3000;
        auto pts = gen_points(0, 77, L);
        CommitterKey ck(ctx, 0, pts);
        std::vector<std::vector<Fe>> in(4);
        std::vector<Affine> got(4), exp(4);
        for (int t = 0; t < 4; t++) { in[t] = gen_scalars(1, 100 + t, L - 10 * t); exp[t] = oracle_commit_(0, std::vector<Affine>(pts.begin(), pts.begin() + (L - 10 * t)), in[t]); }
        std::vector<std::thread> th;
        for (int t = 0; t < 4; t++) th.emplace_back([&, t] { for (int r = 0; r < 5; r++) got[t] = PedersenCommitment::commit(ck, in[t]); });
        for (auto &x : th) x.join();
        bool ok = true;
        for (int t = 0; t < 4; t++) ok = ok && got[t] == exp[t];
        CHECK("four host threads sharing one context", ok);
    }
    // error behaviour: failures surface as exceptions, never as wrong answers
    try {
        auto pts = gen_points(0, 1, 4);
        CommitterKey ck(ctx, 0, pts);                       // no hiding generator
        PedersenCommitment::commit(ck, gen_scalars(1, 1, 4), gen_scalars(1, 2, 1)[0]);
        CHECK("commit with a randomizer but no hiding generator throws", false);
    } catch (const AccmsmError &) { CHECK("commit with a randomizer but no hiding generator throws", true); }
    std::printf("%s: %d failure(s)\n", failures ? "FAILED" : "OK", failures);
    return failures ? 1 : 0;
}
