// ppo_epoch.cuh — ALL minibatch updates of one PPO epoch in ONE persistent cooperative launch, with the gradient
// all-reduce across GPUs done inside the kernel over NVLink peer memory (the caller of the hot path,
// train.py:223-261; network lib/model.py:10-26).
//
// The three-launch update of ppo_update.cuh (k_ppo_forward / k_ppo_backward / k_ppo_adam + an NCCL all-reduce
// between two launches) costs ~55 us per minibatch and the reference schedule has 80 minibatches per epoch.  Here a
// grid of G CTAs x 256 threads stays resident for the whole epoch and walks over the updates:
//
//   phase A   CTA c owns samples [c*S, c*S + S) of the minibatch (S = ceil(B / G)); thread j owns hidden unit j of
//             BOTH nets (its W1 rows, b1 and W2 columns live in registers: 48 parameters).  Forward for the CTA's
//             samples, logits folded over the 256 threads (transposing warp butterfly + 8-way smem fold, fixed
//             order), loss terms and d loss / d logits by thread = sample, then backward: every thread accumulates
//             the gradients of ITS 48 parameters over the CTA's samples in registers and writes them as one
//             coalesced row of the partial-gradient matrix [G][12,298].
//   barrier 1 (grid)
//   phase B   CTA c owns parameter slice c (193 parameters at G = 64): sums the G partials in a fixed order; with
//             several GPUs it publishes the slice in ITS exchange buffer (IPC-mapped into every peer), raises a
//             per-slice flag on every peer (st.release.sys), waits for the peers' flags (ld.acquire.sys) and sums
//             the peers' slices in RANK order — every rank gets the same bits.  Slice sum of squares -> global.
//   barrier 2 (grid)
//   phase C   global norm from the G slice sums (fixed order), clip coefficient, Adam step on the slice
//             (torch.optim.Adam, eps outside the square root as in ppo_update.cuh), loss statistics.
//   barrier 3 (grid): the next update's forward reads every parameter.
//
// No atomics on data, fixed summation orders: results are deterministic and identical on every rank.  Every wait is
// bounded (a missing peer sets *err and the grid drains instead of hanging the GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ppo_update.cuh"

namespace carenv {
namespace ppo {

constexpr int kEpochThreads = kH;                    // 256: thread j = hidden unit j of both nets
constexpr int kLocalK = 48;                          // parameters owned by a thread
constexpr int kLocal = kLocalK * kH + kQ + 1;        // 12,298 (= kNumParams), "local" order: e = k * 256 + j, then b2
constexpr int kLocalPad = 12320;                     // row stride of the partial-gradient matrix (multiple of 32)
constexpr int kXStride = 12416;                      // floats per exchange-buffer half
constexpr int kMaxCtas = 160;
constexpr int kMinCtas = 49;                         // a slice must fit one CTA: ceil(12,298 / G) <= 256
constexpr int kMaxWorld = 8;
static_assert(kLocal == kNumParams, "local layout is a permutation of the flat parameter layout");

// One per rank, in device memory that every peer maps (cudaIpc): the rank's reduced slices, double-buffered by
// update parity, and the flags its peers raise ("slice c of update seq is published in MY buffer").
struct Exchange {
    float x[2][kXStride];
    unsigned long long flag[kMaxWorld][kMaxCtas];
};

struct EpochArgs {
    MutableParams P;
    const float *obs;                 // [.][18]
    const long long *idx;             // [n_updates][B] rows of obs / act / ...
    const float *act, *old_logp, *adv, *ret;
    int B, n_updates;
    float clip_ratio, vf_coef, ent_coef, max_grad_norm, beta1, beta2, eps;
    float *m, *v;                     // Adam moments, flat parameter layout
    const float *lr;
    int *step;
    float *sums4;
    float *partial;                   // [gridDim.x][kLocalPad]
    float *ssq;                       // [2][kMaxCtas]
    float *stat;                      // [2][kMaxCtas][4]: policy loss, entropy, value loss partial sums
    unsigned int *bar;                // grid barrier counter, zero at launch
    int *err;                         // set to 1 by a wait that ran out of time
    int world, rank;
    Exchange *peer[kMaxWorld];        // peer[rank] is the local buffer
    unsigned long long seq_base;      // updates completed by earlier launches (same on every rank)
};

// local element e -> index in the flat layout W1a b1a W2a b2a W1c b1c W2c b2c
__host__ __device__ inline int local_to_flat(int e) {
    if (e >= kLocalK * kH) { const int q = e - kLocalK * kH; return q < kQ ? kOffB2a + q : kOffB2c; }
    const int k = e / kH, j = e % kH;
    if (k < kIn) return kOffW1a + j * kIn + k;
    if (k == kIn) return kOffB1a + j;
    if (k < kIn + 1 + kQ) return kOffW2a + (k - kIn - 1) * kH + j;
    if (k < 2 * kIn + 1 + kQ) return kOffW1c + j * kIn + (k - kIn - 1 - kQ);
    if (k == 2 * kIn + 1 + kQ) return kOffB1c + j;
    return kOffW2c + j;
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_volatile_int(const int *p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr long long kWaitCycles = 4000000000ll;      // ~2 s at 1.9 GHz: a wait this long means a peer is gone

// Grid-wide barrier on a monotonic counter (the launch is cooperative: all CTAs are resident).  Returns non-zero
// when the grid must drain (timeout here or an error raised elsewhere).
__device__ __forceinline__ int grid_barrier(unsigned int *bar, unsigned int &target, int *err, int *s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        const long long t0 = clock64();
        int bad = 0;
        while (ld_acquire_gpu(bar) < target) {
            if (ld_volatile_int(err) != 0) { bad = 1; break; }
            if (clock64() - t0 > kWaitCycles) { atomicExch(err, 1); bad = 1; break; }
        }
        if (!bad) bad = ld_volatile_int(err);
        *s_flag = bad;
    }
    __syncthreads();
    return *s_flag;
}

// 256-thread block sum in a fixed order (warp butterfly, then the 8 warp sums in order).
__device__ __forceinline__ float block_sum_256(float v, float *red8) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red8[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red8[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red8[w];
    return t;
}

// Transposing warp fold: every lane holds N values; afterwards the warp-wide sums are spread over the lanes, lane
// l holding fold_rem<N,16>() consecutive sums starting at fold_base<N,16>(l).  N/2 + N/4 + ... shuffles instead
// of 5 N.
template <int N, int OFF, int NMAX>
__device__ __forceinline__ void warp_fold(float (&v)[NMAX], const int lane) {
    if constexpr (OFF >= 1) {
        if constexpr (N % 2 == 0) {
            constexpr int H = N / 2;
            const bool up = (lane & OFF) != 0;
#pragma unroll
            for (int i = 0; i < H; ++i) {
                const float send = up ? v[i] : v[i + H];
                const float keep = up ? v[i + H] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
            }
            warp_fold<H, OFF / 2, NMAX>(v, lane);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], OFF);
            warp_fold<N, OFF / 2, NMAX>(v, lane);
        }
    }
}
template <int N, int OFF> __device__ __forceinline__ int fold_base(const int lane) {
    if constexpr (OFF >= 1) {
        if constexpr (N % 2 == 0) return ((lane & OFF) ? N / 2 : 0) + fold_base<N / 2, OFF / 2>(lane);
        else return fold_base<N, OFF / 2>(lane);
    } else {
        return 0;
    }
}
template <int N, int OFF> __host__ __device__ constexpr int fold_rem() {
    if constexpr (OFF >= 1) {
        if constexpr (N % 2 == 0) return fold_rem<N / 2, OFF / 2>();
        else return fold_rem<N, OFF / 2>();
    } else {
        return N;
    }
}

constexpr int kEpochSamples = 8;                     // samples per CTA: G = max(64, ceil(B / 8)) CTAs cover B <= 1,024

template <int SMAX>
__global__ void __launch_bounds__(kEpochThreads, 1) k_ppo_epoch(const EpochArgs A) {
    static_assert(SMAX * kIn <= kEpochThreads, "one thread per staged observation element");
    constexpr int kZ = 10;                               // 9 logits + the value per sample
    constexpr int kFold = kZ * SMAX;
    __shared__ __align__(16) float s_x[SMAX][20];        // observation rows of the CTA's samples
    __shared__ float s_red[8][kFold];                    // per-warp sums of the second layer
    __shared__ __align__(16) float s_z[SMAX][12];        // logits | value, then d loss / d logits | d loss / d value
    __shared__ float s_part[SMAX][3];                    // per-sample policy loss, entropy, value loss
    __shared__ float s_red8[8];
    __shared__ float s_ssq[kMaxCtas];
    __shared__ float s_coef;
    __shared__ int s_flag;

    const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
    const int c = blockIdx.x, G = gridDim.x;
    const int B = A.B;
    const int S = (B + G - 1) / G;                       // samples per CTA (<= SMAX, checked by the launcher)
    const int s0 = c * S;
    const int nS = max(0, min(S, B - s0));
    const int slice = (kLocal + G - 1) / G;              // parameters per CTA in phases B / C (<= 256)
    const int e_mine = c * slice + j;
    const bool own = j < slice && e_mine < kLocal;
    const int flat = own ? local_to_flat(e_mine) : 0;
    float *const my_param = param_ptr(A.P, flat);
    const float invB = 1.0f / (float)B;
    const int step0 = *A.step;
    const float lr = A.lr[0];
    unsigned int bar_target = 0;
    for (int i = j; i < SMAX * 20; i += kEpochThreads) (&s_x[0][0])[i] = 0.0f;   // rows >= nS and the padding stay zero
    __syncthreads();

    for (int u = 0; u < A.n_updates; ++u) {
        const long long *idx = A.idx + (size_t)u * (size_t)B;
        const int par = u & 1;

        // ---- parameters of hidden unit j (written by other CTAs in the previous update: read through L2)
        float w1a[kIn], w2a[kQ], w1c[kIn], b1a, b1c, w2c;
#pragma unroll
        for (int k = 0; k < kIn; ++k) { w1a[k] = __ldcg(A.P.w1a + j * kIn + k); w1c[k] = __ldcg(A.P.w1c + j * kIn + k); }
#pragma unroll
        for (int q = 0; q < kQ; ++q) w2a[q] = __ldcg(A.P.w2a + q * kH + j);
        b1a = __ldcg(A.P.b1a + j); b1c = __ldcg(A.P.b1c + j); w2c = __ldcg(A.P.w2c + j);

        // ---- this CTA's observation rows
        if (j < nS * kIn) {
            const int s = j / kIn, k = j % kIn;
            s_x[s][k] = A.obs[(size_t)idx[s0 + s] * kIn + k];
        }

        // ---- advantage statistics of the whole minibatch (train.py:236-237), the same bits in every CTA
        float a_mean, a_inv;
        {
            constexpr int kPer = kMaxBatch / kEpochThreads;          // 4
            float al[kPer];
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int t = j + i * kEpochThreads;
                al[i] = t < B ? A.adv[idx[t]] : 0.0f;
                s += al[i];
            }
            a_mean = block_sum_256(s, s_red8) * invB;
            float s2 = 0.0f;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const float d = j + i * kEpochThreads < B ? al[i] - a_mean : 0.0f;
                s2 += d * d;
            }
            const float var = block_sum_256(s2, s_red8) / (float)(B > 1 ? B - 1 : 1);
            a_inv = 1.0f / fmaxf(sqrtf(var), 1.0e-5f);
        }
        __syncthreads();                                              // s_x complete

        // ---- forward: pre-activations of unit j for the CTA's samples, second-layer products
        float pre_a[SMAX], pre_c[SMAX], contrib[kFold];
#pragma unroll
        for (int s = 0; s < SMAX; ++s) {
            const float4 *xv = reinterpret_cast<const float4 *>(s_x[s]);
            float x[20];
#pragma unroll
            for (int i = 0; i < 5; ++i) { const float4 t = xv[i]; x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w; }
            float pa = b1a, pb = 0.0f, pc = 0.0f, qa = b1c, qb = 0.0f, qc = 0.0f;   // ppo_update.cuh's summation order
#pragma unroll
            for (int k = 0; k < kIn; k += 3) {
                pa = fmaf(w1a[k], x[k], pa); pb = fmaf(w1a[k + 1], x[k + 1], pb); pc = fmaf(w1a[k + 2], x[k + 2], pc);
                qa = fmaf(w1c[k], x[k], qa); qb = fmaf(w1c[k + 1], x[k + 1], qb); qc = fmaf(w1c[k + 2], x[k + 2], qc);
            }
            const bool on = s < nS;
            pre_a[s] = on ? (pa + pb) + pc : 0.0f;
            pre_c[s] = on ? (qa + qb) + qc : 0.0f;
            const float ha = fmaxf(pre_a[s], 0.0f), hc = fmaxf(pre_c[s], 0.0f);
#pragma unroll
            for (int q = 0; q < kQ; ++q) contrib[s * kZ + q] = w2a[q] * ha;
            contrib[s * kZ + kQ] = w2c * hc;
        }
        warp_fold<kFold, 16, kFold>(contrib, lane);
        {
            constexpr int kRem = fold_rem<kFold, 16>();
            const int base = fold_base<kFold, 16>(lane);
#pragma unroll
            for (int i = 0; i < kRem; ++i) s_red[warp][base + i] = contrib[i];   // lanes holding copies write the same bits
        }
        __syncthreads();
        if (j < kFold) {
            float t = s_red[0][j];
#pragma unroll
            for (int w = 1; w < 8; ++w) t += s_red[w][j];
            s_z[j / kZ][j % kZ] = t;
        }
        __syncthreads();

        // ---- loss terms and their derivatives, thread = sample (ppo_update.cuh: k_ppo_forward)
        if (j < SMAX) {
            float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
            if (j < nS) {
                const long long row = idx[s0 + j];
                float z[kQ];
                float mx = -1.0e30f;
#pragma unroll
                for (int q = 0; q < kQ; ++q) { z[q] = s_z[j][q] + __ldcg(A.P.b2a + q); mx = fmaxf(mx, z[q]); }
                float ex[kQ], se = 0.0f;
#pragma unroll
                for (int q = 0; q < kQ; ++q) { ex[q] = expf(z[q] - mx); se += ex[q]; }
                const float lse = logf(se);
                float H = 0.0f, p[kQ], lp[kQ];
#pragma unroll
                for (int q = 0; q < kQ; ++q) { lp[q] = (z[q] - mx) - lse; p[q] = expf(lp[q]); H -= p[q] * lp[q]; }
                const int a = (int)A.act[row];
                float new_lp = lp[0];
#pragma unroll
                for (int q = 1; q < kQ; ++q) new_lp = (a == q) ? lp[q] : new_lp;
                const float ratio = expf(new_lp - A.old_logp[row]);
                const float an = (A.adv[row] - a_mean) * a_inv;
                const float lo = 1.0f - A.clip_ratio, hi = 1.0f + A.clip_ratio;
                const float t1 = -an * ratio, t2 = -an * fminf(fmaxf(ratio, lo), hi);
                const bool inside = ratio >= lo && ratio <= hi;
                const float g_ratio = (inside || t1 > t2) ? -an : 0.0f;     // torch.max / clamp subgradients
                const float g_lp = g_ratio * ratio * invB;
                const float val = s_z[j][kQ] + __ldcg(A.P.b2c);
                const float d = val - A.ret[row];
#pragma unroll
                for (int q = 0; q < kQ; ++q)
                    s_z[j][q] = g_lp * ((a == q ? 1.0f : 0.0f) - p[q]) + A.ent_coef * invB * p[q] * (lp[q] + H);
                s_z[j][kQ] = A.vf_coef * d * invB;
                p0 = fmaxf(t1, t2); p1 = H; p2 = 0.5f * d * d;
            } else {
#pragma unroll
                for (int q = 0; q < kZ; ++q) s_z[j][q] = 0.0f;
            }
            s_part[j][0] = p0; s_part[j][1] = p1; s_part[j][2] = p2;
        }
        __syncthreads();

        // ---- backward: gradients of the 48 parameters of unit j over the CTA's samples
        {
            float g[kLocalK];
#pragma unroll
            for (int k = 0; k < kLocalK; ++k) g[k] = 0.0f;
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
                const float4 *xv = reinterpret_cast<const float4 *>(s_x[s]);
                float x[20];
#pragma unroll
                for (int i = 0; i < 5; ++i) { const float4 t = xv[i]; x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w; }
                const float4 *zv = reinterpret_cast<const float4 *>(s_z[s]);
                const float4 d0 = zv[0], d1 = zv[1], d2 = zv[2];
                const float dz[kQ] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x};
                const float dv = d2.y;
                const float ha = fmaxf(pre_a[s], 0.0f), hc = fmaxf(pre_c[s], 0.0f);
                float dh = 0.0f;
#pragma unroll
                for (int q = 0; q < kQ; ++q) { g[kIn + 1 + q] = fmaf(dz[q], ha, g[kIn + 1 + q]); dh = fmaf(dz[q], w2a[q], dh); }
                dh = pre_a[s] > 0.0f ? dh : 0.0f;                    // ReLU'(pre), zero at pre == 0 like torch
                float dc = pre_c[s] > 0.0f ? dv * w2c : 0.0f;
                g[kLocalK - 1] = fmaf(dv, hc, g[kLocalK - 1]);
#pragma unroll
                for (int k = 0; k < kIn; ++k) {
                    g[k] = fmaf(dh, x[k], g[k]);
                    g[kIn + 1 + kQ + k] = fmaf(dc, x[k], g[kIn + 1 + kQ + k]);
                }
                g[kIn] += dh;
                g[2 * kIn + 1 + kQ] += dc;
            }
            float *row = A.partial + (size_t)c * kLocalPad;
#pragma unroll
            for (int k = 0; k < kLocalK; ++k) __stcg(row + k * kH + j, g[k]);
            if (j < kZ) {                                            // second-layer biases: sum of dz over the samples
                float t = 0.0f;
#pragma unroll
                for (int s = 0; s < SMAX; ++s) t += s_z[s][j];
                __stcg(row + kLocalK * kH + j, t);
            }
            if (j == 32) {                                           // loss partial sums of this CTA
                float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
#pragma unroll
                for (int s = 0; s < SMAX; ++s) { p0 += s_part[s][0]; p1 += s_part[s][1]; p2 += s_part[s][2]; }
                float *dst = A.stat + ((size_t)par * kMaxCtas + c) * 4;
                __stcg(dst, p0); __stcg(dst + 1, p1); __stcg(dst + 2, p2);
            }
        }
        if (grid_barrier(A.bar, bar_target, A.err, &s_flag)) return;

        // ---- phase B: slice c of the gradient, summed over the CTAs (and over the GPUs)
        float gsum = 0.0f;
        if (own) {
            const float *col = A.partial + e_mine;
#pragma unroll 8
            for (int c2 = 0; c2 < G; ++c2) gsum += __ldcg(col + (size_t)c2 * kLocalPad);
        }
        if (A.world > 1) {
            Exchange *me = A.peer[A.rank];
            if (own) me->x[par][e_mine] = gsum;
            __threadfence_system();
            __syncthreads();
            const unsigned long long seq = A.seq_base + (unsigned long long)u + 1ull;
            if (j < A.world && j != A.rank) {
                st_release_sys(&A.peer[j]->flag[A.rank][c], seq);
                const long long t0 = clock64();
                while (ld_acquire_sys(&me->flag[j][c]) < seq) {
                    if (ld_volatile_int(A.err) != 0) break;
                    if (clock64() - t0 > kWaitCycles) { atomicExch(A.err, 1); break; }
                }
            }
            __syncthreads();
            if (own) {
                float t = 0.0f;
                for (int r = 0; r < A.world; ++r) t += ld_relaxed_sys(&A.peer[r]->x[par][e_mine]);   // rank order
                gsum = t * (1.0f / (float)A.world);
            }
        }
        {
            const float ss = block_sum_256(own ? gsum * gsum : 0.0f, s_red8);
            if (j == 0) __stcg(A.ssq + par * kMaxCtas + c, ss);
        }
        if (grid_barrier(A.bar, bar_target, A.err, &s_flag)) return;

        // ---- phase C: clip_grad_norm_ + Adam on the slice, statistics
        if (j < G) s_ssq[j] = __ldcg(A.ssq + par * kMaxCtas + j);
        __syncthreads();
        if (j == 0) {
            float tot = 0.0f;
            for (int c2 = 0; c2 < G; ++c2) tot += s_ssq[c2];
            s_coef = fminf(A.max_grad_norm / (sqrtf(tot) + 1.0e-6f), 1.0f);
        }
        if (c == 0 && j == 32) {
            float pol = 0.0f, ent = 0.0f, vl = 0.0f;
            for (int c2 = 0; c2 < G; ++c2) {
                const float *src = A.stat + ((size_t)par * kMaxCtas + c2) * 4;
                pol += __ldcg(src); ent += __ldcg(src + 1); vl += __ldcg(src + 2);
            }
            pol *= invB; ent *= invB; vl *= invB;
            A.sums4[0] += pol; A.sums4[1] += vl; A.sums4[2] += ent; A.sums4[3] += pol + A.vf_coef * vl - A.ent_coef * ent;
        }
        __syncthreads();
        if (own) {
            const int t = step0 + u + 1;
            const float bc1 = 1.0f - powf(A.beta1, (float)t), bc2 = 1.0f - powf(A.beta2, (float)t);
            const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
            const float gc = gsum * s_coef;
            const float mi = A.beta1 * A.m[flat] + (1.0f - A.beta1) * gc;
            const float vi = A.beta2 * A.v[flat] + (1.0f - A.beta2) * gc * gc;
            A.m[flat] = mi; A.v[flat] = vi;
            __stcg(my_param, __ldcg(my_param) - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + A.eps));
        }
        if (grid_barrier(A.bar, bar_target, A.err, &s_flag)) return;
    }
    if (c == 0 && j == 0) *A.step = step0 + A.n_updates;
}

}  // namespace ppo
}  // namespace carenv
