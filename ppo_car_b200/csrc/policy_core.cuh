// policy_core.cuh — the actor/critic forward pass and the categorical sampling that sit between two
// environment steps in the rollout loop (reference: lib/model.py:10-40, train.py:180-195), fused into the
// rollout kernel so that a whole n_steps rollout is ONE launch.
//
//   actor  : 18 -> 256 (ReLU) -> 9 logits          critic : 18 -> 256 (ReLU) -> 1 value
//   action ~ Categorical(softmax(logits)) by inverse CDF on a counter-based uniform (Philox4x32-10 keyed by
//            the seed, counter = global env id and global step), logprob = log_softmax(logits)[action]
//
// Float32 throughout (the reference trains in float32; the rollout's logprob must agree with the update's
// recomputation to ~1e-6 so that the first PPO ratio is 1).  Two hidden units share one FFMA2: the weight
// pairs are laid out by pack_policy_weights() (host, ppo_car_b200/policy.py) so that every operand is a
// broadcast 16-byte shared-memory load.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace carenv {

constexpr int kHidden = 256;
constexpr int kPolicyObs = 18;
constexpr int kActions = 9;
constexpr int kActorPairFloats = 60;    // 36 w1 + 2 b1 + 2 pad + 20 w2
constexpr int kCriticPairFloats = 44;   // 36 w1 + 2 b1 + 2 pad + 2 w2 + 2 pad
constexpr int kPairFloats = kActorPairFloats + kCriticPairFloats;           // 104
constexpr int kPolicyTail = 12;         // b2[0..9] (b2[9] = 0), b2c, pad
constexpr int kPolicyFloats = (kHidden / 2) * kPairFloats + kPolicyTail;    // 13,324 floats = 53,296 B

struct PolicyOut {
    float logit[10];     // logit[9] is padding (weights and bias zero)
    float value;
};

// obs[18] in registers, w = packed weights in shared memory.
__device__ __forceinline__ void policy_forward(const float (&obs)[kPolicyObs], const float *__restrict__ w,
                                               PolicyOut &out) {
    float2 L[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) L[q] = make_float2(0.0f, 0.0f);
    float2 V = make_float2(0.0f, 0.0f);
#pragma unroll 2
    for (int p = 0; p < kHidden / 2; ++p) {
        const float4 *a4 = reinterpret_cast<const float4 *>(w + p * kPairFloats);
        const float4 *c4 = reinterpret_cast<const float4 *>(w + p * kPairFloats + kActorPairFloats);
        // ---- actor hidden pair
        float4 t = a4[9];                                   // (b1[j], b1[j+1], pad, pad)
        float2 H = make_float2(t.x, t.y);
#pragma unroll
        for (int i = 0; i < 9; ++i) {                       // two inputs per 16-byte load
            const float4 ww = a4[i];
            H = __ffma2_rn(make_float2(obs[2 * i], obs[2 * i]), make_float2(ww.x, ww.y), H);
            H = __ffma2_rn(make_float2(obs[2 * i + 1], obs[2 * i + 1]), make_float2(ww.z, ww.w), H);
        }
        const float h0 = fmaxf(H.x, 0.0f), h1 = fmaxf(H.y, 0.0f);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float4 ww = a4[10 + q];                   // (W2[2q][j], W2[2q+1][j], W2[2q][j+1], W2[2q+1][j+1])
            L[q] = __ffma2_rn(make_float2(h0, h0), make_float2(ww.x, ww.y), L[q]);
            L[q] = __ffma2_rn(make_float2(h1, h1), make_float2(ww.z, ww.w), L[q]);
        }
        // ---- critic hidden pair
        t = c4[9];
        float2 Hc = make_float2(t.x, t.y);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float4 ww = c4[i];
            Hc = __ffma2_rn(make_float2(obs[2 * i], obs[2 * i]), make_float2(ww.x, ww.y), Hc);
            Hc = __ffma2_rn(make_float2(obs[2 * i + 1], obs[2 * i + 1]), make_float2(ww.z, ww.w), Hc);
        }
        t = c4[10];                                         // (W2c[j], W2c[j+1], pad, pad)
        V = __ffma2_rn(make_float2(fmaxf(Hc.x, 0.0f), fmaxf(Hc.y, 0.0f)), make_float2(t.x, t.y), V);
    }
    const float *tail = w + (kHidden / 2) * kPairFloats;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        out.logit[2 * q] = L[q].x + tail[2 * q];
        out.logit[2 * q + 1] = L[q].y + tail[2 * q + 1];
    }
    out.value = (V.x + V.y) + tail[10];
}

// ---- Philox4x32-10 (Salmon et al. 2011), one 128-bit block per (env, step) ----------------------------
__device__ __forceinline__ uint32_t philox_uniform_bits(uint32_t key0, uint32_t key1, uint32_t c0, uint32_t c1,
                                                        uint32_t c2, uint32_t c3) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ key0, n1 = lo1, n2 = hi0 ^ c3 ^ key1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        key0 += W0; key1 += W1;
    }
    return c0;
}

// Inverse-CDF categorical sample.  u in [0,1).  Returns the action; logp = log_softmax(logits)[action].
__device__ __forceinline__ int sample_action(const PolicyOut &p, float u, float &logp, float &u_scaled) {
    float m = p.logit[0];
#pragma unroll
    for (int i = 1; i < kActions; ++i) m = fmaxf(m, p.logit[i]);
    float e[kActions], s = 0.0f;
#pragma unroll
    for (int i = 0; i < kActions; ++i) { e[i] = expf(p.logit[i] - m); s += e[i]; }
    const float target = u * s;
    u_scaled = target;
    int a = kActions - 1;
    float c = 0.0f;
    bool found = false;
#pragma unroll
    for (int i = 0; i < kActions; ++i) {
        c += e[i];
        if (!found && target < c) { a = i; found = true; }
    }
    float la = p.logit[0];
#pragma unroll
    for (int i = 1; i < kActions; ++i) la = (a == i) ? p.logit[i] : la;
    logp = (la - m) - logf(s);
    return a;
}

}  // namespace carenv
