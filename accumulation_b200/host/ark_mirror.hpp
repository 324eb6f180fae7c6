// C++ host side above the C-ABI (include/accmsm.h): a mirror of the interface through which
// arkworks-rs/accumulation reaches its hot path, with the reference's names, argument meaning and error
// behaviour, so that compiled callers (and the Rust `[patch]` of ark-poly-commit) have a
// one-to-one template.  The reference is Rust; this image has no cargo, hence C++.
//
//   reference (ark-poly-commit / ark-accumulation)                     here
//   trivial_pc::CommitterKey{generators, hiding_generator}             accmsm_host::CommitterKey
//   trivial_pc::PedersenCommitment::commit(ck, elems, randomizer)      PedersenCommitment::commit
//   ipa_pc::InnerProductArgPC::cm_commit / check (final key) / open    InnerProductArgPC::{cm_commit, check_final_key, open}
//   ipa_pc::SuccinctCheckPolynomial(Vec<F>)::compute_coeffs            SuccinctCheckPolynomial::compute_coeffs
//   hp_as: compute_hp, compute_t_vecs, combine_vectors, scale_vector,  ASForHadamardProducts::...
//          compute_product_poly_comm, decide   (src/hp_as/mod.rs:278-512, 894-925)
//   r1cs_nark: Matrix<F> = Vec<Vec<(F, usize)>>, matrix_vec_mul        r1cs_nark::{Matrix, matrix_vec_mul, IndexMatrices}
//                                        (src/r1cs_nark_as/r1cs_nark/mod.rs:443-462)
// Every call computes on the GPU; there is no CPU path.  ark's commit is infallible -> failures throw AccmsmError
// (the Rust wrapper panics / maps to PCError at the same places).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/accmsm.h"

namespace accmsm_host {

using Fe = std::array<uint64_t, 4>;          // Fp256 Montgomery image, little-endian limbs (ark-ff memory image)
struct Affine {                               // ark_ec::short_weierstrass_jacobian::GroupAffine
    Fe x{}, y{};
    bool infinity = false;
    bool operator==(const Affine &o) const {  // GroupAffine's PartialEq
        if (infinity || o.infinity) return infinity == o.infinity;
        return x == o.x && y == o.y;
    }
    bool operator!=(const Affine &o) const { return !(*this == o); }
};
enum Curve { PALLAS = 0, VESTA = 1 };
inline int scalar_field(int curve) { return curve == PALLAS ? 1 : 0; }

struct AccmsmError : std::runtime_error { using std::runtime_error::runtime_error; };

class Context {
public:
    explicit Context(int device = 0) {
        int rc = accmsm_init(&ctx_, device);
        if (rc) throw AccmsmError(std::string("accmsm_init: ") + accmsm_strerror(rc) + " (a CUDA device is required; there is no CPU path)");
    }
    // a group of GPUs behind the same calls (accmsm_init_multi): keys are sharded by point range inside the library
    explicit Context(const std::vector<int> &devices, size_t min_shard = size_t(1) << 16) {
        int rc = accmsm_init_multi(&ctx_, devices.data(), (int)devices.size());
        if (rc) throw AccmsmError(std::string("accmsm_init_multi: ") + accmsm_strerror(rc) + " (a CUDA device is required; there is no CPU path)");
        accmsm_set_min_shard(ctx_, min_shard);
    }
    ~Context() { accmsm_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    accmsm_ctx *raw() const { return ctx_; }
    void check(int rc, const char *what) const {
        if (rc) throw AccmsmError(std::string(what) + ": " + accmsm_strerror(rc) + " (" + accmsm_last_error(ctx_) + ")");
    }
private:
    accmsm_ctx *ctx_ = nullptr;
};

// trivial_pc::CommitterKey / the comm_key part of ipa_pc::CommitterKey.  The hiding generator is registered as
// the base after the last generator, so commit(elems, Some(r)) is one pass of the device MSM.
class CommitterKey {
public:
    CommitterKey(std::shared_ptr<Context> ctx, int curve, const std::vector<Affine> &generators,
                 const std::optional<Affine> &hiding_generator = std::nullopt, bool precompute = true)
        : ctx_(std::move(ctx)), curve_(curve), n_(generators.size()), has_hiding_(hiding_generator.has_value()) {
        std::vector<uint64_t> xy;
        std::vector<uint8_t> inf;
        xy.reserve((n_ + 1) * 8);
        auto push = [&](const Affine &g) {
            xy.insert(xy.end(), g.x.begin(), g.x.end());
            xy.insert(xy.end(), g.y.begin(), g.y.end());
            inf.push_back(g.infinity ? 1 : 0);
        };
        for (const auto &g : generators) push(g);
        if (hiding_generator) push(*hiding_generator);
        ctx_->check(accmsm_register_bases(ctx_->raw(), curve, xy.data(), inf.data(), inf.size(), &handle_), "register_bases");
        if (precompute) ctx_->check(accmsm_precompute_bases(ctx_->raw(), handle_, 0), "precompute_bases");
    }
    ~CommitterKey() { accmsm_release_bases(ctx_->raw(), handle_); }
    CommitterKey(const CommitterKey &) = delete;
    CommitterKey &operator=(const CommitterKey &) = delete;
    size_t supported_num_elems() const { return n_; }          // src/hp_as/mod.rs:129,681
    int curve() const { return curve_; }
    uint64_t handle() const { return handle_; }
    size_t hiding_index() const { return n_; }
    bool has_hiding() const { return has_hiding_; }
    const std::shared_ptr<Context> &ctx() const { return ctx_; }
private:
    std::shared_ptr<Context> ctx_;
    int curve_;
    size_t n_;
    bool has_hiding_;
    uint64_t handle_ = 0;
};

inline Affine affine_from(const uint64_t xy[8], uint8_t inf) {
    Affine a;
    std::memcpy(a.x.data(), xy, 32); std::memcpy(a.y.data(), xy + 4, 32);
    a.infinity = inf != 0;
    return a;
}
inline void affine_to(const Affine &a, uint64_t xy[8]) { std::memcpy(xy, a.x.data(), 32); std::memcpy(xy + 4, a.y.data(), 32); }

struct PedersenCommitment {
    // PedersenCommitment::commit(ck, elems, randomizer) -> G  (src/hp_as/mod.rs:196; ark-ec truncates to min(len))
    static Affine commit(const CommitterKey &ck, const std::vector<Fe> &elems, const std::optional<Fe> &randomizer = std::nullopt) {
        if (randomizer && !ck.has_hiding()) throw AccmsmError("commit: the key has no hiding generator");
        size_t n = std::min(elems.size(), ck.supported_num_elems());
        uint64_t xy[8]; uint8_t inf = 0;
        ck.ctx()->check(accmsm_commit(ck.ctx()->raw(), ck.handle(), n, n ? elems[0].data() : nullptr, ck.hiding_index(),
                                      randomizer ? randomizer->data() : nullptr, xy, &inf), "commit");
        return affine_from(xy, inf);
    }
    // several vectors of equal length over one key in shared passes (hp_as::decide, NARK prove)
    static std::vector<Affine> commit_batch(const CommitterKey &ck, const std::vector<std::vector<Fe>> &vecs) {
        std::vector<Affine> out;
        if (vecs.empty()) return out;
        size_t n = std::min(vecs[0].size(), ck.supported_num_elems());
        std::vector<uint64_t> flat(vecs.size() * n * 4), xy(vecs.size() * 8);
        std::vector<uint8_t> inf(vecs.size());
        for (size_t j = 0; j < vecs.size(); j++) {
            if (vecs[j].size() < n) throw AccmsmError("commit_batch: vectors of different length");
            if (n) std::memcpy(flat.data() + j * n * 4, vecs[j][0].data(), n * 32);
        }
        ck.ctx()->check(accmsm_msm_batch(ck.ctx()->raw(), ck.handle(), 0, n, vecs.size(), flat.data(), 1, xy.data(), inf.data()), "msm_batch");
        for (size_t j = 0; j < vecs.size(); j++) out.push_back(affine_from(xy.data() + 8 * j, inf[j]));
        return out;
    }
};

// ipa_pc::SuccinctCheckPolynomial(pub Vec<F>)  (src/ipa_pc_as/mod.rs:245,400,418)
struct SuccinctCheckPolynomial {
    std::vector<Fe> challenges;   // .0, xi_1 first
    std::vector<Fe> compute_coeffs(const Context &ctx, int field) const {
        std::vector<Fe> out(size_t(1) << challenges.size());
        ctx.check(accmsm_compute_coeffs(ctx.raw(), field, challenges.empty() ? nullptr : challenges[0].data(), (int)challenges.size(),
                                        out[0].data()), "compute_coeffs");
        return out;
    }
};

struct IpaProofCore {             // the fields of ipa_pc::Proof produced by the opening loop
    std::vector<Affine> l_vec, r_vec;
    Affine final_comm_key;
    Fe c{};
    std::vector<Fe> round_challenges;
};

struct InnerProductArgPC {
    static Affine cm_commit(const CommitterKey &comm_key, const std::vector<Fe> &scalars, const std::optional<Fe> &randomizer = std::nullopt) {
        return PedersenCommitment::commit(comm_key, scalars, randomizer);
    }
    // tail of check_individual_opening_challenges (AS decide, src/ipa_pc_as/mod.rs:836-845):
    // final_key = cm_commit(vk.comm_key, h.compute_coeffs());  Ok(final_key == proof.final_comm_key)
    static bool check_final_key(const CommitterKey &vk, const SuccinctCheckPolynomial &h, const Affine &final_comm_key) {
        uint64_t exp[8], xy[8]; uint8_t inf = 0; int accept = 0;
        affine_to(final_comm_key, exp);
        vk.ctx()->check(accmsm_ipa_check_final_key(vk.ctx()->raw(), vk.handle(), h.challenges.empty() ? nullptr : h.challenges[0].data(),
                                                   (int)h.challenges.size(), exp, final_comm_key.infinity, &accept, xy, &inf), "ipa_check_final_key");
        return accept != 0;
    }
    // the opening loop of open_individual_opening_challenges (src/ipa_pc_as/mod.rs:454-462); the caller is the host
    // transcript: round_challenge(l, r) -> (xi, xi^-1)
    // h' = xi_0 * h: pass the point h_prime, or (faster) xi_0 when h is the hiding generator registered with `ck`
    static IpaProofCore open(const CommitterKey &ck, const std::vector<Fe> &combined_coeffs, int log_d, const Fe &point,
                             const Affine &h_prime, const std::function<std::pair<Fe, Fe>(const Affine &, const Affine &)> &round_challenge,
                             const std::optional<Fe> &xi0 = std::nullopt) {
        uint64_t hp[8]; affine_to(h_prime, hp);
        uint64_t sess = 0;
        if (xi0 && !ck.has_hiding()) throw AccmsmError("open: the key has no hiding generator");
        ck.ctx()->check(accmsm_ipa_open_begin(ck.ctx()->raw(), ck.handle(), combined_coeffs.empty() ? nullptr : combined_coeffs[0].data(),
                                              combined_coeffs.size(), log_d, point.data(), xi0 ? nullptr : hp, &sess), "ipa_open_begin");
        if (xi0) ck.ctx()->check(accmsm_ipa_open_use_hiding_generator(ck.ctx()->raw(), sess, ck.hiding_index(), xi0->data()), "ipa_open_use_hiding_generator");
        IpaProofCore p;
        for (int r = 0; r < log_d; r++) {
            uint64_t l[8], rr[8]; uint8_t li = 0, ri = 0;
            ck.ctx()->check(accmsm_ipa_open_round(ck.ctx()->raw(), sess, l, &li, rr, &ri), "ipa_open_round");
            p.l_vec.push_back(affine_from(l, li)); p.r_vec.push_back(affine_from(rr, ri));
            auto xi = round_challenge(p.l_vec.back(), p.r_vec.back());
            ck.ctx()->check(accmsm_ipa_open_fold(ck.ctx()->raw(), sess, xi.first.data(), xi.second.data()), "ipa_open_fold");
            p.round_challenges.push_back(xi.first);
        }
        uint64_t fk[8];
        ck.ctx()->check(accmsm_ipa_open_finish(ck.ctx()->raw(), sess, fk, p.c.data()), "ipa_open_finish");
        p.final_comm_key = affine_from(fk, 0);
        return p;
    }
    // The same loop with ONE library call per round (accmsm_ipa_open_fold_round: fold with the challenge squeezed from the
    // previous (l, r) -- its inverse is computed inside the library, on the host, like upstream's round_challenge.inverse() --
    // and run the next round): what the patched open_individual_opening_challenges uses.  round_challenge(l, r) -> xi.
    static IpaProofCore open_one_call_per_round(const CommitterKey &ck, const std::vector<Fe> &combined_coeffs, int log_d, const Fe &point,
                                                const Affine &h_prime, const std::function<Fe(const Affine &, const Affine &)> &round_challenge,
                                                const std::optional<Fe> &xi0 = std::nullopt) {
        uint64_t hp[8]; affine_to(h_prime, hp);
        uint64_t sess = 0;
        if (xi0 && !ck.has_hiding()) throw AccmsmError("open: the key has no hiding generator");
        ck.ctx()->check(accmsm_ipa_open_begin(ck.ctx()->raw(), ck.handle(), combined_coeffs.empty() ? nullptr : combined_coeffs[0].data(),
                                              combined_coeffs.size(), log_d, point.data(), xi0 ? nullptr : hp, &sess), "ipa_open_begin");
        if (xi0) ck.ctx()->check(accmsm_ipa_open_use_hiding_generator(ck.ctx()->raw(), sess, ck.hiding_index(), xi0->data()), "ipa_open_use_hiding_generator");
        IpaProofCore p;
        uint64_t l[8], rr[8]; uint8_t li = 0, ri = 0; int done = log_d == 0;
        if (!done) ck.ctx()->check(accmsm_ipa_open_round(ck.ctx()->raw(), sess, l, &li, rr, &ri), "ipa_open_round");
        while (!done) {
            p.l_vec.push_back(affine_from(l, li)); p.r_vec.push_back(affine_from(rr, ri));
            Fe xi = round_challenge(p.l_vec.back(), p.r_vec.back());
            p.round_challenges.push_back(xi);
            ck.ctx()->check(accmsm_ipa_open_fold_round(ck.ctx()->raw(), sess, xi.data(), l, &li, rr, &ri, &done), "ipa_open_fold_round");
        }
        uint64_t fk[8];
        ck.ctx()->check(accmsm_ipa_open_finish(ck.ctx()->raw(), sess, fk, p.c.data()), "ipa_open_finish");
        p.final_comm_key = affine_from(fk, 0);
        return p;
    }
};

// ark_ec::msm::VariableBaseMSM::multi_scalar_mul(&bases, &scalars).into_affine() on bases that are NOT a registered key: the short
// linear combinations of commitments (src/hp_as/mod.rs:391-406, src/ipa_pc_as/mod.rs:322-343) and the 2k + 3 term group equation of
// IpaPC::succinct_check (:198-205).  Scalars are BigInteger256 images (canonical, `into_repr()` done by the caller, as upstream).
struct VariableBaseMSM {
    static Affine multi_scalar_mul(const Context &ctx, int curve, const std::vector<Affine> &bases, const std::vector<Fe> &scalars_canonical) {
        const size_t n = std::min(bases.size(), scalars_canonical.size());      // ark-ec truncates to the shorter slice
        std::vector<uint64_t> xy(n * 8); std::vector<uint8_t> inf(n);
        for (size_t i = 0; i < n; i++) { affine_to(bases[i], xy.data() + 8 * i); inf[i] = bases[i].infinity; }
        uint64_t out[8]; uint8_t oi = 0;
        ctx.check(accmsm_msm_oneshot(ctx.raw(), curve, xy.data(), inf.data(), n ? scalars_canonical[0].data() : nullptr, 0, n, out, &oi), "msm_oneshot");
        return affine_from(out, oi);
    }
    // m such MSMs of equal length in shared passes (the succinct checks of all inputs and accumulators of one prove / verify)
    static std::vector<Affine> multi_scalar_mul_batch(const Context &ctx, int curve, const std::vector<std::vector<Affine>> &bases,
                                                      const std::vector<std::vector<Fe>> &scalars_canonical) {
        const size_t m = std::min(bases.size(), scalars_canonical.size());
        size_t n = m ? SIZE_MAX : 0;
        for (size_t j = 0; j < m; j++) n = std::min(n, std::min(bases[j].size(), scalars_canonical[j].size()));
        std::vector<uint64_t> xy(m * n * 8), sc(m * n * 4); std::vector<uint8_t> inf(m * n);
        for (size_t j = 0; j < m; j++) for (size_t i = 0; i < n; i++) {
            affine_to(bases[j][i], xy.data() + 8 * (j * n + i)); inf[j * n + i] = bases[j][i].infinity;
            std::copy(scalars_canonical[j][i].begin(), scalars_canonical[j][i].end(), sc.begin() + 4 * (j * n + i));
        }
        std::vector<uint64_t> out(m * 8); std::vector<uint8_t> oi(m);
        ctx.check(accmsm_msm_oneshot_batch(ctx.raw(), curve, xy.data(), inf.data(), sc.data(), 0, n, m, out.data(), oi.data()), "msm_oneshot_batch");
        std::vector<Affine> res(m);
        for (size_t j = 0; j < m; j++) res[j] = affine_from(out.data() + 8 * j, oi[j]);
        return res;
    }
};

struct ASForHadamardProducts {
    struct InputInstance { Affine comm_1, comm_2, comm_3; };                       // src/hp_as/data_structures.rs:14-23
    struct Randomness { Fe rand_1, rand_2, rand_3; };
    struct InputWitness { std::vector<Fe> a_vec, b_vec; std::optional<Randomness> randomness; };   // :54-63

    static std::vector<Fe> compute_hp(const Context &ctx, int field, const std::vector<Fe> &a, const std::vector<Fe> &b) {   // :278-285
        size_t n = std::min(a.size(), b.size());
        std::vector<Fe> out(n);
        if (n) ctx.check(accmsm_vec_hadamard(ctx.raw(), field, a[0].data(), b[0].data(), n, out[0].data()), "vec_hadamard");
        return out;
    }
    static std::vector<Fe> scale_vector(const Context &ctx, int field, const std::vector<Fe> &v, const Fe &coeff) {          // :482-489
        std::vector<Fe> out(v.size());
        if (!v.empty()) ctx.check(accmsm_vec_scale(ctx.raw(), field, v[0].data(), v.size(), coeff.data(), out[0].data()), "vec_scale");
        return out;
    }
    static std::vector<Fe> combine_vectors(const Context &ctx, int field, const std::vector<const std::vector<Fe> *> &vectors,
                                           const std::vector<Fe> &challenges, const std::vector<Fe> *hiding_vec = nullptr) { // :492-512
        std::vector<const uint64_t *> ptrs; std::vector<size_t> lens; size_t cap = hiding_vec ? hiding_vec->size() : 0;
        for (auto *v : vectors) { ptrs.push_back(v->empty() ? nullptr : (*v)[0].data()); lens.push_back(v->size()); cap = std::max(cap, v->size()); }
        std::vector<Fe> out(cap);
        size_t olen = 0;
        ctx.check(accmsm_vec_lincomb(ctx.raw(), field, ptrs.data(), lens.data(), (int)vectors.size(), challenges.empty() ? nullptr : challenges[0].data(),
                                     hiding_vec && !hiding_vec->empty() ? (*hiding_vec)[0].data() : nullptr, hiding_vec ? hiding_vec->size() : 0,
                                     cap ? out[0].data() : nullptr, cap, &olen), "vec_lincomb");
        out.resize(olen);
        return out;
    }
    // decide (:894-925): one fused device call
    static bool decide(const CommitterKey &ck, const InputInstance &instance, const InputWitness &witness) {
        size_t n = std::min({witness.a_vec.size(), witness.b_vec.size(), ck.supported_num_elems()});
        uint64_t exp[24]; uint8_t einf[3] = {instance.comm_1.infinity, instance.comm_2.infinity, instance.comm_3.infinity};
        affine_to(instance.comm_1, exp); affine_to(instance.comm_2, exp + 8); affine_to(instance.comm_3, exp + 16);
        uint64_t r[12];
        if (witness.randomness) {
            if (!ck.has_hiding()) throw AccmsmError("decide: the key has no hiding generator");
            std::memcpy(r, witness.randomness->rand_1.data(), 32); std::memcpy(r + 4, witness.randomness->rand_2.data(), 32);
            std::memcpy(r + 8, witness.randomness->rand_3.data(), 32);
        }
        int accept = 0;
        ck.ctx()->check(accmsm_hp_decide(ck.ctx()->raw(), ck.handle(), n ? witness.a_vec[0].data() : nullptr, n ? witness.b_vec[0].data() : nullptr, n,
                                         ck.hiding_index(), witness.randomness ? r : nullptr, exp, einf, &accept, nullptr, nullptr), "hp_decide");
        return accept != 0;
    }
};

namespace r1cs_nark {
// ark_relations::r1cs::Matrix<F> = Vec<Vec<(F, usize)>>  (src/r1cs_nark_as/r1cs_nark/data_structures.rs:17-29)
using Matrix = std::vector<std::vector<std::pair<Fe, size_t>>>;

struct Csr { std::vector<uint32_t> row_ptr, cols; std::vector<uint64_t> coeffs; };
inline Csr to_csr(const Matrix &m) {
    Csr c; c.row_ptr.push_back(0);
    for (const auto &row : m) {
        for (const auto &e : row) { c.cols.push_back((uint32_t)e.second); c.coeffs.insert(c.coeffs.end(), e.first.begin(), e.first.end()); }
        c.row_ptr.push_back((uint32_t)c.cols.size());
    }
    return c;
}
// matrix_vec_mul(matrix, input, witness) for A, B, C at once (src/r1cs_nark_as/r1cs_nark/mod.rs:443-447)
inline std::vector<std::vector<Fe>> matrix_vec_mul(const Context &ctx, int field, const std::vector<const Matrix *> &mats,
                                                   const std::vector<Fe> &input, const std::vector<Fe> &witness) {
    std::vector<Csr> csr; for (auto *m : mats) csr.push_back(to_csr(*m));
    size_t n_rows = mats.empty() ? 0 : mats[0]->size();
    std::vector<std::vector<Fe>> out(mats.size(), std::vector<Fe>(n_rows));
    std::vector<const uint32_t *> rp, cl; std::vector<const uint64_t *> cf; std::vector<uint64_t *> op;
    for (size_t i = 0; i < mats.size(); i++) { rp.push_back(csr[i].row_ptr.data()); cl.push_back(csr[i].cols.data()); cf.push_back(csr[i].coeffs.data()); op.push_back(n_rows ? out[i][0].data() : nullptr); }
    if (n_rows) ctx.check(accmsm_csr_matvec(ctx.raw(), field, (int)mats.size(), rp.data(), cl.data(), cf.data(), n_rows, input.empty() ? nullptr : input[0].data(),
                                            input.size(), witness.empty() ? nullptr : witness[0].data(), witness.size(), op.data()), "csr_matvec");
    return out;
}
// IndexProverKey{a, b, c, ck} with the matrices resident on the device (registered at index time)
class IndexMatrices {
public:
    IndexMatrices(const CommitterKey &ck, const Matrix &a, const Matrix &b, const Matrix &c) : ck_(ck), n_rows_(a.size()) {
        Csr ca = to_csr(a), cb = to_csr(b), cc = to_csr(c);
        const uint32_t *rp[3] = {ca.row_ptr.data(), cb.row_ptr.data(), cc.row_ptr.data()};
        const uint32_t *cl[3] = {ca.cols.data(), cb.cols.data(), cc.cols.data()};
        const uint64_t *cf[3] = {ca.coeffs.data(), cb.coeffs.data(), cc.coeffs.data()};
        ck.ctx()->check(accmsm_register_csr(ck.ctx()->raw(), scalar_field(ck.curve()), 3, rp, cl, cf, n_rows_, &handle_), "register_csr");
    }
    ~IndexMatrices() { accmsm_release_csr(ck_.ctx()->raw(), handle_); }
    // z_M = M (input || witness), comm_M = Commit(z_M, blinder_M)   (prove :183-185 + :216-218; decide src/r1cs_nark_as/mod.rs:1052-1097)
    std::pair<std::vector<std::vector<Fe>>, std::vector<Affine>> matvec_commit(const std::vector<Fe> &input, const std::vector<Fe> &witness,
                                                                              const std::optional<std::array<Fe, 3>> &blinders = std::nullopt) const {
        std::vector<std::vector<Fe>> vecs(3, std::vector<Fe>(n_rows_));
        uint64_t *op[3] = {vecs[0][0].data(), vecs[1][0].data(), vecs[2][0].data()};
        uint64_t xy[24]; uint8_t inf[3]; uint64_t bl[12];
        if (blinders) for (int m = 0; m < 3; m++) std::memcpy(bl + 4 * m, (*blinders)[m].data(), 32);
        ck_.ctx()->check(accmsm_csr_matvec_commit(ck_.ctx()->raw(), ck_.handle(), handle_, input.empty() ? nullptr : input[0].data(), input.size(),
                                                  witness.empty() ? nullptr : witness[0].data(), witness.size(), ck_.hiding_index(),
                                                  blinders ? bl : nullptr, op, xy, inf), "csr_matvec_commit");
        std::vector<Affine> comms;
        for (int m = 0; m < 3; m++) comms.push_back(affine_from(xy + 8 * m, inf[m]));
        return {vecs, comms};
    }
private:
    const CommitterKey &ck_;
    size_t n_rows_;
    uint64_t handle_ = 0;
};
}  // namespace r1cs_nark

}  // namespace accmsm_host
