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
