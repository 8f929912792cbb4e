// elastic_elgamal_b200.hpp -- header-only C++17 mirror of the reference's API surface for the batch path, over the
// C ABI of eg_b200.h.  Names follow slowli/elastic-elgamal: PublicKey (src/keys/mod.rs:122), Ciphertext
// (src/encryption.rs:96), RingProof (src/proofs/ring.rs:282), LogEqualityProof (src/proofs/log_equality.rs:96),
// EncryptedChoice / ChoiceParams (src/app/choice.rs:276,132).  Objects are the reference's `to_bytes` forms.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "eg_b200.h"

namespace elastic_elgamal_b200 {

struct Error : std::runtime_error {
    eg_status status;
    Error(eg_status s, const std::string &m) : std::runtime_error(m), status(s) {}
};

// per-item outcome, mirroring VerificationError / ChoiceVerificationError / QuadraticVotingError
enum class Verdict : uint8_t {
    Ok = EG_V_OK, Malformed = EG_V_MALFORMED, ChallengeMismatch = EG_V_CHALLENGE_MISMATCH, ChoiceSum = EG_V_CHOICE_SUM,
    ChoiceRange = EG_V_CHOICE_RANGE, QvCreditRange = EG_V_QV_CREDIT_RANGE, QvCreditEquivalence = EG_V_QV_CREDIT_EQUIV,
    QvVariantBase = EG_V_QV_VARIANT_BASE
};

using Element = std::array<uint8_t, 32>;
using Scalar = std::array<uint8_t, 32>;
struct Ciphertext { Element random_element, blinded_element; };                       // to_bytes: R || B
struct LogEqualityProof { Scalar challenge, response; };                              // to_bytes: c || s
struct RingProof { Scalar common_challenge; std::vector<Scalar> ring_responses; };    // to_bytes: e0 || s...
struct PublicKey { Element bytes; };
// src/proofs/commitment.rs:126-135 (serde only in the reference; bytes = struct field order)
struct CommitmentEquivalenceProof { Scalar challenge, randomness_response, value_response, commitment_response; };
// src/proofs/possession.rs:71-76
struct ProofOfPossession { Scalar challenge; std::vector<Scalar> responses; };
// src/proofs/range.rs:446-450 + the ciphertext it speaks about
struct RangeProof { std::vector<Ciphertext> partial_ciphertexts; RingProof inner; };

struct ChoiceParams { PublicKey receiver; uint32_t options_count; bool single; };     // ChoiceParams::{single, multi}
struct EncryptedChoice { std::vector<Ciphertext> choices; RingProof range_proof; LogEqualityProof sum_proof; };

struct ChoiceBatchResult { std::vector<Verdict> verdicts; std::vector<Ciphertext> tally; };

class Engine {
  public:
    explicit Engine(int device = 0) {
        eg_status st = eg_ctx_create(device, &ctx_);
        if (st != EG_SUCCESS) throw Error(st, "eg_ctx_create failed: no CUDA device (the engine has no CPU fallback)");
    }
    ~Engine() { eg_ctx_destroy(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    // PublicKey::from_bytes
    void set_receiver(const PublicKey &key) { check(eg_ctx_set_receiver(ctx_, key.bytes.data())); }

    // keys.iter().map(|(ct, proof)| receiver.verify_bool(ct, proof))
    std::vector<Verdict> verify_bool_batch(const std::vector<Ciphertext> &cts, const std::vector<RingProof> &proofs) {
        if (cts.size() != proofs.size()) throw Error(EG_ERR_LEN_MISMATCH, "ciphertexts / proofs size mismatch");
        std::vector<uint8_t> c, p, v(cts.size());
        for (size_t i = 0; i < cts.size(); i++) {
            if (proofs[i].ring_responses.size() != 2) throw Error(EG_ERR_LEN_MISMATCH, "items in all rings");   // ring.rs:310-315
            append(c, cts[i]);
            append(p, proofs[i]);
        }
        check(eg_verify_bool_batch(ctx_, cts.size(), c.data(), p.data(), v.data()));
        return to_verdicts(v);
    }

    // ballots.iter().map(|b| b.verify(&params)) + the tally fold (examples/voting.rs:189-204)
    ChoiceBatchResult verify_batch(const ChoiceParams &params, const std::vector<EncryptedChoice> &ballots) {
        const uint32_t m = params.options_count;
        std::vector<uint8_t> c, r, s, v(ballots.size()), t(64 * (size_t)m);
        for (const auto &b : ballots) {
            if (b.choices.size() != m) throw Error(EG_ERR_LEN_MISMATCH, "number of options in the ballot");        // choice.rs:149-158
            if (b.range_proof.ring_responses.size() != 2 * (size_t)m) throw Error(EG_ERR_LEN_MISMATCH, "items in all rings");
            for (const auto &ct : b.choices) append(c, ct);
            append(r, b.range_proof);
            if (params.single) { s.insert(s.end(), b.sum_proof.challenge.begin(), b.sum_proof.challenge.end());
                                 s.insert(s.end(), b.sum_proof.response.begin(), b.sum_proof.response.end()); }
        }
        check(eg_verify_choice_batch(ctx_, ballots.size(), m, params.single ? 1 : 0, c.data(), r.data(),
                                     params.single ? s.data() : nullptr, v.data(), t.data()));
        ChoiceBatchResult out;
        out.verdicts = to_verdicts(v);
        out.tally.resize(m);
        for (uint32_t k = 0; k < m; k++) {
            std::copy(t.begin() + 64 * k, t.begin() + 64 * k + 32, out.tally[k].random_element.begin());
            std::copy(t.begin() + 64 * k + 32, t.begin() + 64 * k + 64, out.tally[k].blinded_element.begin());
        }
        return out;
    }

    // proofs.iter().map(|(ct, proof)| receiver.verify_range(&range, ct, proof))   (keys/impls.rs:143-151)
    std::vector<Verdict> verify_range_batch(const eg_range &range, const std::vector<Ciphertext> &cts, const std::vector<RangeProof> &proofs,
                                            const std::string &label = "ciphertext_range") {
        if (cts.size() != proofs.size()) throw Error(EG_ERR_LEN_MISMATCH, "ciphertexts / proofs size mismatch");
        size_t total = 0;
        for (uint32_t r = 0; r < range.n_rings; r++) total += range.size[r];
        std::vector<uint8_t> c, pc, r, v(cts.size());
        for (size_t i = 0; i < cts.size(); i++) {
            if (proofs[i].partial_ciphertexts.size() + 1 != range.n_rings) throw Error(EG_ERR_LEN_MISMATCH, "number of rings");  // range.rs:555-559
            if (proofs[i].inner.ring_responses.size() != total) throw Error(EG_ERR_LEN_MISMATCH, "items in all rings");
            append(c, cts[i]);
            for (const auto &ct : proofs[i].partial_ciphertexts) append(pc, ct);
            append(r, proofs[i].inner);
        }
        check(eg_verify_range_batch(ctx_, &range, label.c_str(), cts.size(), c.data(), range.n_rings > 1 ? pc.data() : nullptr, r.data(), v.data()));
        return to_verdicts(v);
    }

    // CommitmentEquivalenceProof::verify over a batch (commitment.rs:198-248); `blinding_base` = H
    void set_blinding_base(const Element &base) { check(eg_ctx_set_blinding_base(ctx_, base.data())); }
    std::vector<Verdict> verify_commitment_equivalence_batch(const std::vector<Ciphertext> &cts, const std::vector<Element> &commitments,
                                                             const std::vector<CommitmentEquivalenceProof> &proofs, const std::string &label) {
        if (cts.size() != proofs.size() || cts.size() != commitments.size()) throw Error(EG_ERR_LEN_MISMATCH, "batch size mismatch");
        std::vector<uint8_t> c, k, p, v(cts.size());
        for (size_t i = 0; i < cts.size(); i++) {
            append(c, cts[i]);
            k.insert(k.end(), commitments[i].begin(), commitments[i].end());
            for (const Scalar *s : {&proofs[i].challenge, &proofs[i].randomness_response, &proofs[i].value_response, &proofs[i].commitment_response})
                p.insert(p.end(), s->begin(), s->end());
        }
        check(eg_verify_commitment_equiv_batch(ctx_, label.c_str(), cts.size(), c.data(), k.data(), p.data(), v.data()));
        return to_verdicts(v);
    }

    // ProofOfPossession::verify over a batch of proofs for `keys_per_proof` keys each (possession.rs:137-163)
    std::vector<Verdict> verify_possession_batch(const std::vector<std::vector<PublicKey>> &keys, const std::vector<ProofOfPossession> &proofs,
                                                 const std::string &label) {
        if (keys.size() != proofs.size()) throw Error(EG_ERR_LEN_MISMATCH, "batch size mismatch");
        if (keys.empty()) return {};
        const size_t kpp = keys[0].size();
        std::vector<uint8_t> k, p, v(keys.size());
        for (size_t i = 0; i < keys.size(); i++) {
            if (keys[i].size() != kpp || proofs[i].responses.size() != kpp) throw Error(EG_ERR_LEN_MISMATCH, "public keys");   // possession.rs:147
            for (const auto &key : keys[i]) k.insert(k.end(), key.bytes.begin(), key.bytes.end());
            p.insert(p.end(), proofs[i].challenge.begin(), proofs[i].challenge.end());
            for (const auto &s : proofs[i].responses) p.insert(p.end(), s.begin(), s.end());
        }
        check(eg_verify_possession_batch(ctx_, label.c_str(), (uint32_t)kpp, keys.size(), k.data(), p.data(), v.data()));
        return to_verdicts(v);
    }

    eg_ctx *raw() { return ctx_; }

  private:
    eg_ctx *ctx_ = nullptr;
    void check(eg_status st) { if (st != EG_SUCCESS) throw Error(st, eg_last_error(ctx_)); }
    static void append(std::vector<uint8_t> &o, const Ciphertext &ct) {
        o.insert(o.end(), ct.random_element.begin(), ct.random_element.end());
        o.insert(o.end(), ct.blinded_element.begin(), ct.blinded_element.end());
    }
    static void append(std::vector<uint8_t> &o, const RingProof &p) {
        o.insert(o.end(), p.common_challenge.begin(), p.common_challenge.end());
        for (const auto &s : p.ring_responses) o.insert(o.end(), s.begin(), s.end());
    }
    static std::vector<Verdict> to_verdicts(const std::vector<uint8_t> &v) {
        std::vector<Verdict> out(v.size());
        for (size_t i = 0; i < v.size(); i++) out[i] = static_cast<Verdict>(v[i]);
        return out;
    }
};

}  // namespace elastic_elgamal_b200
