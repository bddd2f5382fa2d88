// ppo_update.cuh — one clipped-surrogate PPO minibatch update of the reference network in three launches
// (the caller of the hot path, train.py:223-261; network lib/model.py:10-26).
//
//   k_ppo_forward   thread = sample: both nets' forward, log-softmax, ratio, minibatch-normalised advantage,
//                   clipped surrogate, value loss, entropy; writes d loss / d logits (actor) and d loss / d value
//                   (critic) per sample plus the gathered observation rows to a scratch buffer
//   k_ppo_backward  thread = (hidden unit, sample slice): recomputes its pre-activation per sample and accumulates
//                   the gradients of ITS rows of W1 / b1 and ITS column of W2 in registers — no atomics, the
//                   summation order is fixed, results are deterministic
//   k_ppo_adam      (after the optional NCCL all-reduce of the 12,298 gradients) global-norm clip
//                   (torch.nn.utils.clip_grad_norm_), Adam step (torch.optim.Adam, eps inside sqrt(v)/sqrt(bc2) + eps),
//                   running loss statistics
//
// A PyTorch autograd update of the same minibatch is ~60 kernels (0.3 ms even inside a CUDA graph); with the
// reference's schedule of 80 minibatch updates per epoch that was 24 ms of a 40 ms epoch at 32,768 envs per GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "policy_core.cuh"

namespace carenv {
namespace ppo {

constexpr int kIn = kPolicyObs;                 // 18
constexpr int kH = kHidden;                     // 256
constexpr int kQ = kActions;                    // 9
constexpr int kDz = 12;                         // d loss / d logits per sample, padded to 3 float4
// flat parameter / gradient layout: W1a, b1a, W2a, b2a, W1c, b1c, W2c, b2c
constexpr int kOffW1a = 0, kOffB1a = kOffW1a + kH * kIn, kOffW2a = kOffB1a + kH, kOffB2a = kOffW2a + kQ * kH;
constexpr int kOffW1c = kOffB2a + kQ, kOffB1c = kOffW1c + kH * kIn, kOffW2c = kOffB1c + kH, kOffB2c = kOffW2c + kH;
constexpr int kNumParams = kOffB2c + 1;         // 12,298
constexpr int kMaxBatch = 1024;
constexpr int kFwdThreads = 64;                 // samples per forward CTA
constexpr int kBwdUnits = 32, kBwdSlices = 8;   // backward CTA: 32 hidden units x 8 sample slices = 256 threads

struct Params { const float *w1a, *b1a, *w2a, *b2a, *w1c, *b1c, *w2c, *b2c; };

// scratch layout (floats): xs [B][20] | dz [B][12] | dv [B] | partial sums [2 * ceil(B / 64)][2]
__host__ __device__ inline int scratch_xs() { return 0; }
__host__ __device__ inline int scratch_dz(int B) { return B * 20; }
__host__ __device__ inline int scratch_dv(int B) { return scratch_dz(B) + B * kDz; }
__host__ __device__ inline int scratch_part(int B) { return scratch_dv(B) + B; }
__host__ __device__ inline int scratch_floats(int B) { return scratch_part(B) + 4 * ((B + kFwdThreads - 1) / kFwdThreads) + 8; }

__device__ __forceinline__ float block_sum_64(float v, float *red) {      // 64 threads = 2 warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return red[0] + red[1];
}

// grid (ceil(B / 64), 2): blockIdx.y = 0 actor, 1 critic.
__global__ void __launch_bounds__(kFwdThreads)
k_ppo_forward(Params P, const float *__restrict__ obs, int obs_is_gathered, const long long *__restrict__ idx,
              const float *__restrict__ act, const float *__restrict__ old_logp, const float *__restrict__ adv,
              const float *__restrict__ ret, int B, float clip_ratio, float vf_coef, float ent_coef,
              float *__restrict__ scratch) {
    extern __shared__ __align__(16) float sm[];
    const bool critic = blockIdx.y == 1;
    const int Q = critic ? 1 : kQ;
    float *sW1 = sm;                              // [256][20]: 18 weights, bias, pad  (16-byte rows)
    float *sW2 = sW1 + kH * 20;                   // [256][12]: W2[q][j] for q < Q, zero padding
    float *red = sW2 + kH * kDz;                  // 4 floats
    {   // stage the weights: every thread copies whole rows with all of a row's loads in flight at once
        const float *w1 = critic ? P.w1c : P.w1a, *b1 = critic ? P.b1c : P.b1a, *w2 = critic ? P.w2c : P.w2a;
#pragma unroll 2
        for (int j = threadIdx.x; j < kH; j += blockDim.x) {
            const float2 *r = reinterpret_cast<const float2 *>(w1 + j * kIn);     // rows are 72 bytes: 8-byte aligned
            float2 v[kIn / 2];
#pragma unroll
            for (int c = 0; c < kIn / 2; ++c) v[c] = r[c];
            const float b = b1[j];
            float wq[kDz];
#pragma unroll
            for (int q = 0; q < kDz; ++q) wq[q] = q < Q ? w2[q * kH + j] : 0.0f;
            float4 *d1 = reinterpret_cast<float4 *>(sW1 + j * 20);
            d1[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
            d1[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
            d1[2] = make_float4(v[4].x, v[4].y, v[5].x, v[5].y);
            d1[3] = make_float4(v[6].x, v[6].y, v[7].x, v[7].y);
            d1[4] = make_float4(v[8].x, v[8].y, b, 0.0f);
            float4 *d2 = reinterpret_cast<float4 *>(sW2 + j * kDz);
            d2[0] = make_float4(wq[0], wq[1], wq[2], wq[3]);
            d2[1] = make_float4(wq[4], wq[5], wq[6], wq[7]);
            d2[2] = make_float4(wq[8], wq[9], wq[10], wq[11]);
        }
    }
    // advantage statistics of the whole minibatch (train.py:236-237: mean, unbiased std clamped at 1e-5)
    float a_mean = 0.0f, a_inv = 0.0f;
    if (!critic) {
        constexpr int kPer = kMaxBatch / kFwdThreads;        // 16 values per thread, gathered with all loads in flight
        long long r[kPer];
        float al[kPer];
#pragma unroll
        for (int c = 0; c < kPer; ++c) { const int i = threadIdx.x + c * kFwdThreads; r[c] = i < B ? idx[i] : 0; }
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < kPer; ++c) { al[c] = threadIdx.x + c * kFwdThreads < B ? adv[r[c]] : 0.0f; s += al[c]; }
        a_mean = block_sum_64(s, red) / (float)B;
        float s2 = 0.0f;
#pragma unroll
        for (int c = 0; c < kPer; ++c) {
            const float d = threadIdx.x + c * kFwdThreads < B ? al[c] - a_mean : 0.0f;
            s2 += d * d;
        }
        const float var = block_sum_64(s2, red) / (float)(B > 1 ? B - 1 : 1);
        a_inv = 1.0f / fmaxf(sqrtf(var), 1.0e-5f);
    }
    __syncthreads();

    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < B;
    const long long row = live ? idx[s] : 0;
    float x[kIn];
    {
        const float *src = obs + (size_t)(obs_is_gathered ? (live ? s : 0) : row) * kIn;
#pragma unroll
        for (int k = 0; k < kIn; ++k) x[k] = src[k];
    }
    float z[kDz];
#pragma unroll
    for (int q = 0; q < kDz; ++q) z[q] = 0.0f;
#pragma unroll 2
    for (int j = 0; j < kH; ++j) {
        const float4 *w = reinterpret_cast<const float4 *>(sW1 + j * 20);
        float pre;
        {
            const float4 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
            float pa = w4.z, pb = 0.0f, pc = 0.0f;                            // bias; three chains instead of one
            pa = fmaf(w0.x, x[0], pa); pb = fmaf(w0.y, x[1], pb); pc = fmaf(w0.z, x[2], pc);
            pa = fmaf(w0.w, x[3], pa); pb = fmaf(w1.x, x[4], pb); pc = fmaf(w1.y, x[5], pc);
            pa = fmaf(w1.z, x[6], pa); pb = fmaf(w1.w, x[7], pb); pc = fmaf(w2.x, x[8], pc);
            pa = fmaf(w2.y, x[9], pa); pb = fmaf(w2.z, x[10], pb); pc = fmaf(w2.w, x[11], pc);
            pa = fmaf(w3.x, x[12], pa); pb = fmaf(w3.y, x[13], pb); pc = fmaf(w3.z, x[14], pc);
            pa = fmaf(w3.w, x[15], pa); pb = fmaf(w4.x, x[16], pb); pc = fmaf(w4.y, x[17], pc);
            pre = (pa + pb) + pc;
        }
        const float h = fmaxf(pre, 0.0f);
        const float4 *v = reinterpret_cast<const float4 *>(sW2 + j * kDz);
        if (critic) {
            z[0] = fmaf(v[0].x, h, z[0]);
        } else {
            const float4 v0 = v[0], v1 = v[1], v2 = v[2];
            z[0] = fmaf(v0.x, h, z[0]); z[1] = fmaf(v0.y, h, z[1]); z[2] = fmaf(v0.z, h, z[2]); z[3] = fmaf(v0.w, h, z[3]);
            z[4] = fmaf(v1.x, h, z[4]); z[5] = fmaf(v1.y, h, z[5]); z[6] = fmaf(v1.z, h, z[6]); z[7] = fmaf(v1.w, h, z[7]);
            z[8] = fmaf(v2.x, h, z[8]);
        }
    }
    const float invB = 1.0f / (float)B;
    float part0 = 0.0f, part1 = 0.0f;             // actor: policy loss, entropy; critic: value loss
    if (critic) {
        const float v = z[0] + P.b2c[0];
        const float d = v - (live ? ret[row] : 0.0f);
        if (live) {
            scratch[scratch_dv(B) + s] = vf_coef * d * invB;
            part0 = 0.5f * d * d;
        }
    } else {
        float m = -1.0e30f;
#pragma unroll
        for (int q = 0; q < kQ; ++q) { z[q] += P.b2a[q]; m = fmaxf(m, z[q]); }
        float e[kQ], se = 0.0f;
#pragma unroll
        for (int q = 0; q < kQ; ++q) { e[q] = expf(z[q] - m); se += e[q]; }
        const float lse = logf(se);
        float H = 0.0f, p[kQ], lp[kQ];
#pragma unroll
        for (int q = 0; q < kQ; ++q) { lp[q] = (z[q] - m) - lse; p[q] = expf(lp[q]); H -= p[q] * lp[q]; }
        const int a = live ? (int)act[row] : 0;
        float new_lp = lp[0];
#pragma unroll
        for (int q = 1; q < kQ; ++q) new_lp = (a == q) ? lp[q] : new_lp;
        const float ratio = expf(new_lp - (live ? old_logp[row] : 0.0f));
        const float an = ((live ? adv[row] : 0.0f) - a_mean) * a_inv;
        const float lo = 1.0f - clip_ratio, hi = 1.0f + clip_ratio;
        const float t1 = -an * ratio, t2 = -an * fminf(fmaxf(ratio, lo), hi);
        const bool inside = ratio >= lo && ratio <= hi;
        const float g_ratio = (inside || t1 > t2) ? -an : 0.0f;          // torch.max / clamp subgradients
        const float g_lp = g_ratio * ratio * invB;
        if (live) {
            float4 *dst = reinterpret_cast<float4 *>(scratch + scratch_dz(B) + (size_t)s * kDz);
            float dz[kDz];
#pragma unroll
            for (int q = 0; q < kQ; ++q)
                dz[q] = g_lp * ((a == q ? 1.0f : 0.0f) - p[q]) + ent_coef * invB * p[q] * (lp[q] + H);
            dz[9] = dz[10] = dz[11] = 0.0f;
            dst[0] = make_float4(dz[0], dz[1], dz[2], dz[3]);
            dst[1] = make_float4(dz[4], dz[5], dz[6], dz[7]);
            dst[2] = make_float4(dz[8], 0.0f, 0.0f, 0.0f);
            float4 *xd = reinterpret_cast<float4 *>(scratch + scratch_xs() + (size_t)s * 20);
            xd[0] = make_float4(x[0], x[1], x[2], x[3]);
            xd[1] = make_float4(x[4], x[5], x[6], x[7]);
            xd[2] = make_float4(x[8], x[9], x[10], x[11]);
            xd[3] = make_float4(x[12], x[13], x[14], x[15]);
            xd[4] = make_float4(x[16], x[17], 1.0f, 0.0f);               // the 1 multiplies the bias
            part0 = fmaxf(t1, t2);
            part1 = H;
        }
    }
    const float s0 = block_sum_64(part0, red), s1 = block_sum_64(part1, red);
    if (threadIdx.x == 0) {
        float *dst = scratch + scratch_part(B) + 2 * (blockIdx.y * gridDim.x + blockIdx.x);
        dst[0] = s0; dst[1] = s1;
    }
}

// grid (256 / 32, 2): CTA = 32 hidden units of one net; thread = (unit, sample slice).
__global__ void __launch_bounds__(kBwdUnits * kBwdSlices)
k_ppo_backward(Params P, int B, const float *__restrict__ scratch, float *__restrict__ grads) {
    extern __shared__ __align__(16) float sm[];
    const bool critic = blockIdx.y == 1;
    float *sx = sm;                               // [B][20]
    float *sd = sx + (size_t)B * 20;              // actor: [B][12]; critic: [B]
    float *sacc = sd + (size_t)B * kDz;           // [8 slices][32 units][29] partial gradients
    {
        const float4 *src = reinterpret_cast<const float4 *>(scratch + scratch_xs());
#pragma unroll 4
        for (int i = threadIdx.x; i < B * 5; i += blockDim.x) reinterpret_cast<float4 *>(sx)[i] = src[i];
        if (critic) {
#pragma unroll 2
            for (int i = threadIdx.x; i < B; i += blockDim.x) sd[i] = scratch[scratch_dv(B) + i];
        } else {
            const float4 *sz = reinterpret_cast<const float4 *>(scratch + scratch_dz(B));
#pragma unroll 4
            for (int i = threadIdx.x; i < B * 3; i += blockDim.x) reinterpret_cast<float4 *>(sd)[i] = sz[i];
        }
    }
    const int unit = threadIdx.x % kBwdUnits, slice = threadIdx.x / kBwdUnits;
    const int j = blockIdx.x * kBwdUnits + unit;
    const float *w1 = (critic ? P.w1c : P.w1a) + j * kIn;
    float w[kIn], w2[kQ];
#pragma unroll
    for (int k = 0; k < kIn; ++k) w[k] = w1[k];
    const float b1 = (critic ? P.b1c : P.b1a)[j];
#pragma unroll
    for (int q = 0; q < kQ; ++q) w2[q] = critic ? (q == 0 ? P.w2c[j] : 0.0f) : P.w2a[q * kH + j];
    float gw1[kIn], gb1 = 0.0f, gw2[kQ];
#pragma unroll
    for (int k = 0; k < kIn; ++k) gw1[k] = 0.0f;
#pragma unroll
    for (int q = 0; q < kQ; ++q) gw2[q] = 0.0f;
    __syncthreads();
    for (int s = slice; s < B; s += kBwdSlices) {
        const float4 *xv = reinterpret_cast<const float4 *>(sx + (size_t)s * 20);
        float x[20];
#pragma unroll
        for (int c = 0; c < 5; ++c) { const float4 t = xv[c]; x[4 * c] = t.x; x[4 * c + 1] = t.y; x[4 * c + 2] = t.z; x[4 * c + 3] = t.w; }
        float pa = b1, pb = 0.0f, pc = 0.0f;       // the forward kernel's summation order: same ReLU mask
#pragma unroll
        for (int k = 0; k < kIn; k += 3) { pa = fmaf(w[k], x[k], pa); pb = fmaf(w[k + 1], x[k + 1], pb); pc = fmaf(w[k + 2], x[k + 2], pc); }
        const float pre = (pa + pb) + pc;
        const float h = fmaxf(pre, 0.0f);
        float dh = 0.0f;
        if (critic) {
            const float dv = sd[s];
            gw2[0] = fmaf(dv, h, gw2[0]);
            dh = dv * w2[0];
        } else {
            const float4 *dzv = reinterpret_cast<const float4 *>(sd + (size_t)s * kDz);
            const float4 d0 = dzv[0], d1 = dzv[1], d2 = dzv[2];
            const float dz[kQ] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x};
#pragma unroll
            for (int q = 0; q < kQ; ++q) { gw2[q] = fmaf(dz[q], h, gw2[q]); dh = fmaf(dz[q], w2[q], dh); }
        }
        dh = pre > 0.0f ? dh : 0.0f;                                   // ReLU'(pre), zero at pre == 0 like torch
#pragma unroll
        for (int k = 0; k < kIn; ++k) gw1[k] = fmaf(dh, x[k], gw1[k]);
        gb1 += dh;
    }
    // fold the 8 sample slices in a fixed order
    constexpr int kAcc = kIn + 1 + kQ + 1;        // 29 floats per (slice, unit): odd stride, conflict-free
    float *mine = sacc + (size_t)(slice * kBwdUnits + unit) * kAcc;
#pragma unroll
    for (int k = 0; k < kIn; ++k) mine[k] = gw1[k];
    mine[kIn] = gb1;
#pragma unroll
    for (int q = 0; q < kQ; ++q) mine[kIn + 1 + q] = gw2[q];
    __syncthreads();
    const int Q = critic ? 1 : kQ;
    for (int i = threadIdx.x; i < kBwdUnits * (kIn + 1 + Q); i += blockDim.x) {
        const int u = i / (kIn + 1 + Q), c = i % (kIn + 1 + Q);
        float acc = 0.0f;
#pragma unroll
        for (int sl = 0; sl < kBwdSlices; ++sl) acc += sacc[(size_t)(sl * kBwdUnits + u) * kAcc + c];
        const int jj = blockIdx.x * kBwdUnits + u;
        int off;
        if (c < kIn) off = (critic ? kOffW1c : kOffW1a) + jj * kIn + c;
        else if (c == kIn) off = (critic ? kOffB1c : kOffB1a) + jj;
        else off = critic ? kOffW2c + jj : kOffW2a + (c - kIn - 1) * kH + jj;
        grads[off] = acc;
    }
    // second-layer bias: sum over the samples (one CTA per net does it; warp q folds output q, fixed order)
    if (blockIdx.x == 0) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int q = warp; q < Q; q += (kBwdUnits * kBwdSlices) / 32) {
            float acc = 0.0f;
            for (int s = lane; s < B; s += 32) acc += critic ? sd[s] : sd[(size_t)s * kDz + q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) grads[(critic ? kOffB2c : kOffB2a) + q] = acc;
        }
    }
}

struct MutableParams { float *w1a, *b1a, *w2a, *b2a, *w1c, *b1c, *w2c, *b2c; };

__device__ __forceinline__ float *param_ptr(const MutableParams &P, int i) {
    if (i < kOffB1a) return P.w1a + i;
    if (i < kOffW2a) return P.b1a + (i - kOffB1a);
    if (i < kOffB2a) return P.w2a + (i - kOffW2a);
    if (i < kOffW1c) return P.b2a + (i - kOffB2a);
    if (i < kOffB1c) return P.w1c + (i - kOffW1c);
    if (i < kOffW2c) return P.b1c + (i - kOffB1c);
    if (i < kOffB2c) return P.w2c + (i - kOffW2c);
    return P.b2c;
}

// One CTA of 1,024 threads: clip_grad_norm_ + Adam over the 12,298 parameters, statistics of the minibatch.
__global__ void __launch_bounds__(1024)
k_ppo_adam(MutableParams P, float *__restrict__ grads, float grad_scale, float *__restrict__ m, float *__restrict__ v,
           const float *__restrict__ lr, int *__restrict__ step, float beta1, float beta2, float eps,
           float max_grad_norm, const float *__restrict__ scratch, int B, float vf_coef, float ent_coef,
           float *__restrict__ sums4) {
    __shared__ float red[32];
    __shared__ float s_coef;
    constexpr int kPer = (kNumParams + 1023) / 1024;         // 13 parameters per thread, all loads in flight at once
    float g[kPer], mo[kPer], vo[kPer], po[kPer];
    float ss = 0.0f;
#pragma unroll
    for (int c = 0; c < kPer; ++c) {
        const int i = threadIdx.x + c * 1024;
        const bool ok = i < kNumParams;
        g[c] = ok ? grads[i] * grad_scale : 0.0f;
        mo[c] = ok ? m[i] : 0.0f;
        vo[c] = ok ? v[i] : 0.0f;
        po[c] = ok ? *param_ptr(P, i) : 0.0f;
        ss += g[c] * g[c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int w = 0; w < 32; ++w) tot += red[w];
        const float norm = sqrtf(tot);
        s_coef = fminf(max_grad_norm / (norm + 1.0e-6f), 1.0f);          // clip_grad_norm_
        // minibatch statistics: partial sums of the forward CTAs in a fixed order
        const int nb = (B + kFwdThreads - 1) / kFwdThreads;
        float pol = 0.0f, ent = 0.0f, vl = 0.0f;
        for (int b = 0; b < nb; ++b) {
            pol += scratch[scratch_part(B) + 2 * b];
            ent += scratch[scratch_part(B) + 2 * b + 1];
            vl += scratch[scratch_part(B) + 2 * (nb + b)];
        }
        pol /= (float)B; ent /= (float)B; vl /= (float)B;
        sums4[0] += pol; sums4[1] += vl; sums4[2] += ent; sums4[3] += pol + vf_coef * vl - ent_coef * ent;
    }
    __syncthreads();
    const int t = *step + 1;
    const float bc1 = 1.0f - powf(beta1, (float)t), bc2 = 1.0f - powf(beta2, (float)t);
    const float step_size = lr[0] / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const float coef = s_coef;
#pragma unroll
    for (int c = 0; c < kPer; ++c) {
        const int i = threadIdx.x + c * 1024;
        if (i < kNumParams) {
            const float gc = g[c] * coef;
            const float mi = beta1 * mo[c] + (1.0f - beta1) * gc;
            const float vi = beta2 * vo[c] + (1.0f - beta2) * gc * gc;
            m[i] = mi; v[i] = vi;
            *param_ptr(P, i) = po[c] - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *step = t;
}

}  // namespace ppo
}  // namespace carenv
