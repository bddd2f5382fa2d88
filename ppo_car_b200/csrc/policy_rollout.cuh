// policy_rollout.cuh — the fused policy + environment rollout kernels and the tcgen05 bring-up test.
// Textually included by carenv_kernels.cu inside its anonymous namespace (after the step kernels, whose helpers
// stage_tables / store_pose / env_step it uses); kept in its own file so that the step path (carenv_core.cuh,
// carenv_tables.h, carenv_kernels.cu) can be hashed on its own by bench.py / profiles/make_traffic.py.
#pragma once
// tcgen05 bring-up / unit test: D[128,256] = A[128,24] * B[256,24]^T with kind::tf32, accumulators in TMEM.
__global__ void __launch_bounds__(128) k_tc_gemm_test(const float *__restrict__ A, const float *__restrict__ B,
                                                      float *__restrict__ D) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char *tsm = smem + ((128u - (tc::smem_u32(smem) & 127u)) & 127u);
    unsigned char *sA = tsm, *sB = tsm + tc::kABytes;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < tc::kTileM * tc::kK; i += blockDim.x)
        *reinterpret_cast<float *>(sA + tc::operand_offset(i / tc::kK, i % tc::kK)) = A[i];
    for (int i = tid; i < tc::kTileN * tc::kK; i += blockDim.x)
        *reinterpret_cast<float *>(sB + tc::operand_offset(i / tc::kK, i % tc::kK)) = B[i];
    if (tid == 0) tc::mbar_init(tc::smem_u32(&mbar), 1);
    if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, 256);
        for (int ks = 0; ks < tc::kK / 8; ++ks) {
            const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA) + ks * 2 * tc::kLBO);
            const uint64_t db = tc::make_smem_desc(tc::smem_u32(sB) + ks * 2 * tc::kLBO);
            tc::mma_tf32(tbase, da, db, idesc, ks > 0);
        }
        tc::mma_commit(tc::smem_u32(&mbar));
    }
    tc::mbar_wait(tc::smem_u32(&mbar), 0);
    tc::tc_fence_after();
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 256; c += 16) {
        float v[16];
        tc::tmem_ld16(lane_base + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) D[(size_t)tid * 256 + c + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tbase);
}

// Fused rollout: actor/critic forward, categorical sampling, CarEnv.step and the Buffer
// row writes of train.py:173-195 for n_steps steps in ONE launch.  One thread per environment; the packed
// policy weights (53 KB) and the per-thread-indexed track tables live in shared memory.
#ifndef CARENV_POLICY_BLOCK
#define CARENV_POLICY_BLOCK 128
#endif
#ifndef CARENV_POLICY_MIN_BLOCKS
#define CARENV_POLICY_MIN_BLOCKS 3
#endif
constexpr int kPolicyBlock = CARENV_POLICY_BLOCK;
template <int U>
__global__ void __launch_bounds__(kPolicyBlock, CARENV_POLICY_MIN_BLOCKS)
k_policy_rollout(const __grid_constant__ TrackParams P, const Tables G, const float *__restrict__ weights,
                 int n_envs, int n_steps, int env_offset, unsigned long long seed, unsigned long long step0,
                 double2 *__restrict__ pos, double2 *__restrict__ vel, int4 *__restrict__ ints,
                 float *__restrict__ cur_obs, float *__restrict__ cur_term, float *__restrict__ cur_trunc,
                 double reward_scale, float *__restrict__ obs_buf, float *__restrict__ act_buf,
                 float *__restrict__ rew_buf, float *__restrict__ val_buf, float *__restrict__ term_buf,
                 float *__restrict__ trunc_buf, float *__restrict__ logp_buf, float *__restrict__ last_val,
                 float *__restrict__ u_dbg, unsigned long long *stats, int table_bytes, int obs_mode) {
    extern __shared__ __align__(16) unsigned char smem[];
    float *sw = reinterpret_cast<float *>(smem + table_bytes);
    for (int i = threadIdx.x; i < kPolicyFloats / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(sw)[i] = reinterpret_cast<const float4 *>(weights)[i];
    const Tables T = stage_tables(G, P.n_gates, P.n_seg, smem);      // ends with __syncthreads()
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;

    EnvState s;
    {
        const double2 p = pos[e], v = vel[e];
        const int4 q = ints[e];
        s.px = p.x; s.py = p.y; s.vx = v.x; s.vy = v.y;
        s.k = q.x; s.t = q.y; s.next_gate = q.z; s.passed = q.w;
    }
    float obs[kObsDim];
    {
        const float2 *src = reinterpret_cast<const float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
        for (int i = 0; i < kObsDim / 2; ++i) { const float2 v = src[i]; obs[2 * i] = v.x; obs[2 * i + 1] = v.y; }
    }
    float tc = cur_term[e], uc = cur_trunc[e];
    const uint32_t gid = (uint32_t)(env_offset + e);
    PolicyOut po;
    for (int t = 0; t < n_steps; ++t) {
        const size_t idx = (size_t)t * (size_t)n_envs + (size_t)e;
        policy_forward(obs, sw, po);
        const unsigned long long gs = step0 + (unsigned long long)t;
        const uint32_t bits = philox_uniform_bits((uint32_t)seed, (uint32_t)(seed >> 32), gid, (uint32_t)gs,
                                                  (uint32_t)(gs >> 32), 0x43415245u);
        const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);
        float logp, us;
        const int a = sample_action(po, u, logp, us);
        if (obs_mode == kObsPose) {
            store_pose(reinterpret_cast<PoseRec *>(obs_buf) + idx, s, obs[2], obs[3]);
        } else {
            float2 *dst = reinterpret_cast<float2 *>(obs_buf + idx * kObsDim);
#pragma unroll
            for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
        }
        act_buf[idx] = (float)a;
        val_buf[idx] = po.value;
        logp_buf[idx] = logp;
        term_buf[idx] = tc;
        trunc_buf[idx] = uc;
        if (u_dbg) u_dbg[idx] = u;
        StepResult o;
        env_step<U>(s, a, reward_scale, P, T, o, stats);
        rew_buf[idx] = o.reward;
#pragma unroll
        for (int i = 0; i < kObsDim; ++i) obs[i] = o.obs[i];
        tc = o.terminated ? 1.0f : 0.0f;
        uc = o.truncated ? 1.0f : 0.0f;
    }
    pos[e] = make_double2(s.px, s.py);
    vel[e] = make_double2(s.vx, s.vy);
    ints[e] = make_int4(s.k, s.t, s.next_gate, s.passed);
    {
        float2 *dst = reinterpret_cast<float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
        for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
    }
    cur_term[e] = tc;
    cur_trunc[e] = uc;
    if (last_val) {                                          // bootstrap value of the state after the rollout
        policy_forward(obs, sw, po);
        last_val[e] = po.value;
    }
}

// ---- fused rollout, tensor-core version ---------------------------------------------------------------------
// Same contract as k_policy_rollout, but the two 18->256 layers (23.6 of the 29 kflop per env-step) run on the
// 5th-gen tensor cores: per step every thread writes its observation row (hi/lo TF32 split, a constant 1 in
// column 18 carries the bias) into the K-major UMMA operand tile of its 128-env group, one elected thread issues
// tcgen05.mma kind::tf32 for all kTcTiles groups — D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (3xTF32: float32-level
// accuracy) — into tensor memory, and after the commit barrier every thread reads ITS row of pre-activations
// back with tcgen05.ld, applies ReLU and the small second layer (FFMA2) and goes on to sampling and the env step.
// TMEM: 512 columns = kTcTiles x kTcCols; the 256 hidden units of a net are produced in 256 / kTcCols rounds.
// TILES = 128-env groups per CTA: 4 (512 environment threads, TMEM 4 x 128 columns, two rounds per net) for large
// shards, 2 (256 threads, 2 x 256 columns, one round per net) for shards that would otherwise leave SMs empty.
constexpr int kTcBFloats = tc::kTileN * tc::kK;              // 6,144 floats per B operand
constexpr int kTcW2Off = 4 * kTcBFloats;                     // [actor hi | actor lo | critic hi | critic lo]
constexpr int kTcW2cOff = kTcW2Off + kHidden * 10;           // W2 actor as [j][5] pairs, then w2c[j]
constexpr int kTcTailOff = kTcW2cOff + kHidden;              // b2[0..9], b2c, pad
constexpr int kTcWeightFloats = kTcTailOff + 12;             // 27,404 floats

// Packing of the reference network's parameters (nn.Linear layout: weight [out][in], lib/model.py:10-26) into the
// two layouts above, one thread per packed float: runs after every optimiser epoch, so it is one launch and not
// forty small tensor operations (13 ms per epoch in PyTorch, as much as the rollout of 32,768 envs x 1,024 steps).
template <bool TC>
__global__ void __launch_bounds__(256)
k_pack_policy(const float *__restrict__ w1a, const float *__restrict__ b1a, const float *__restrict__ w2a,
              const float *__restrict__ b2a, const float *__restrict__ w1c, const float *__restrict__ b1c,
              const float *__restrict__ w2c, const float *__restrict__ b2c, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.0f;
    if (TC) {
        if (i >= kTcWeightFloats) return;
        if (i < 4 * kTcBFloats) {                            // [actor hi | actor lo | critic hi | critic lo], UMMA layout
            const int which = i / kTcBFloats, r = i % kTcBFloats;
            const int j = (r / 192) * 8 + (r % 32) / 4, k = ((r % 192) / 32) * 4 + r % 4;
            const float *w1 = which < 2 ? w1a : w1c, *b1 = which < 2 ? b1a : b1c;
            const float x = k < kObsDim ? w1[j * kObsDim + k] : (k == kObsDim ? b1[j] : 0.0f);
            const float hi = tc::to_tf32(x);
            v = (which & 1) ? tc::to_tf32(x - hi) : hi;
        } else if (i < kTcW2cOff) {                          // actor second layer as [j][10] (q = 9 is padding)
            const int r = i - kTcW2Off, j = r / 10, q = r % 10;
            v = q < kActions ? w2a[q * kHidden + j] : 0.0f;
        } else if (i < kTcTailOff) {
            v = w2c[i - kTcW2cOff];
        } else {
            const int r = i - kTcTailOff;                    // b2[0..8], 0, b2c, 0
            v = r < kActions ? b2a[r] : (r == 10 ? b2c[0] : 0.0f);
        }
    } else {
        if (i >= kPolicyFloats) return;
        constexpr int kBody = (kHidden / 2) * kPairFloats;
        if (i < kBody) {
            const int p = i / kPairFloats, r = i % kPairFloats;   // pair of hidden units (2p, 2p + 1)
            const bool critic = r >= kActorPairFloats;
            const int c = critic ? r - kActorPairFloats : r;
            const float *w1 = critic ? w1c : w1a, *b1 = critic ? b1c : b1a;
            if (c < 36) v = w1[(2 * p + (c & 1)) * kObsDim + (c >> 1)];
            else if (c < 38) v = b1[2 * p + (c - 36)];
            else if (c >= 40) {
                const int d = c - 40;
                if (!critic) {                               // (W2[2q][j], W2[2q+1][j]) for j = 2p, 2p + 1
                    const int q = d >> 2, sft = (d >> 1) & 1, rr = d & 1, row = 2 * q + rr;
                    v = row < kActions ? w2a[row * kHidden + 2 * p + sft] : 0.0f;
                } else if (d < 2) {
                    v = w2c[2 * p + d];
                }
            }
        } else {
            const int r = i - kBody;
            v = r < kActions ? b2a[r] : (r == 10 ? b2c[0] : 0.0f);
        }
    }
    out[i] = v;
}

template <int U, int TILES>
__global__ void __launch_bounds__(TILES * 128 + 32, 1)
k_policy_rollout_tc(const __grid_constant__ TrackParams P, const Tables G, const float *__restrict__ weights,
                    int n_envs, int n_steps, int env_offset, unsigned long long seed, unsigned long long step0,
                    double2 *__restrict__ pos, double2 *__restrict__ vel, int4 *__restrict__ ints,
                    float *__restrict__ cur_obs, float *__restrict__ cur_term, float *__restrict__ cur_trunc,
                    double reward_scale, float *__restrict__ obs_buf, float *__restrict__ act_buf,
                    float *__restrict__ rew_buf, float *__restrict__ val_buf, float *__restrict__ term_buf,
                    float *__restrict__ trunc_buf, float *__restrict__ logp_buf, float *__restrict__ last_val,
                    float *__restrict__ u_dbg, unsigned long long *stats, int table_bytes, int obs_mode) {
    constexpr int kTcTiles = TILES;                          // 128-env groups per CTA (+ one MMA-issuing warp)
    constexpr int kTcCols = 512 / TILES;                     // accumulator columns per group and round
    constexpr int kTcRounds = kHidden / kTcCols;             // rounds per net
    static_assert(TILES == 2 || TILES == 4, "TMEM: 512 columns = TILES x (256 or 128)");
    extern __shared__ __align__(16) unsigned char smem[];
    // [tables | pad to a 128-byte boundary | weights | A tiles (hi, lo per group)]
    const int w_off = (int)((tc::smem_u32(smem) + (uint32_t)table_bytes + 127u) / 128u * 128u - tc::smem_u32(smem));
    float *sw = reinterpret_cast<float *>(smem + w_off);
    unsigned char *sA = smem + w_off + ((kTcWeightFloats * 4 + 127) / 128 * 128);
    // per 128-env group: "accumulator ready" (MMA -> group), "accumulator consumed" and "operand rows written"
    // (group -> MMA issuer).  Groups never wait for each other: no CTA-wide barrier inside the step loop.
    __shared__ __align__(8) unsigned long long mbar_full[kTcTiles], mbar_cons[kTcTiles], mbar_rows[kTcTiles];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kTcWeightFloats / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(sw)[i] = reinterpret_cast<const float4 *>(weights)[i];
    if (tid == 0)
        for (int g = 0; g < kTcTiles; ++g) {
            tc::mbar_init(tc::smem_u32(&mbar_full[g]), 1);
            tc::mbar_init(tc::smem_u32(&mbar_cons[g]), 128);
            tc::mbar_init(tc::smem_u32(&mbar_rows[g]), 128);
        }
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    const Tables T = stage_tables(G, P.n_gates, P.n_seg, smem);      // ends with __syncthreads()
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t sB = tc::smem_u32(sw);
    const uint32_t idesc = tc::make_idesc_tf32(128, kTcCols);
    const int n_forward = n_steps + (last_val ? 1 : 0);      // forward passes per thread

    if (warp == kTcTiles * 4) {
        // ===== MMA issuer (one elected lane): runs ahead of the groups, throttled only by the barriers =====
        if ((tid & 31) == 0) {
            uint32_t p_rows = 0, p_cons = 0;                 // same phase for every group: all advance in lock step here
            for (int f = 0; f < n_forward; ++f) {
#pragma unroll 1
                for (int rnd = 0; rnd < 2 * kTcRounds; ++rnd) {
                    const int net = rnd / kTcRounds, part = rnd % kTcRounds;
                    const uint32_t bh = sB + (uint32_t)((2 * net) * kTcBFloats * 4 + part * (kTcCols / 8) * tc::kSBO);
                    const uint32_t bl = bh + (uint32_t)(kTcBFloats * 4);
#pragma unroll 1
                    for (int g = 0; g < kTcTiles; ++g) {
                        if (rnd == 0) tc::mbar_wait(tc::smem_u32(&mbar_rows[g]), p_rows);      // this step's rows
                        if (f > 0 || rnd > 0) tc::mbar_wait(tc::smem_u32(&mbar_cons[g]), p_cons);  // D_g read out
                        tc::tc_fence_after();
                        const uint32_t ah = tc::smem_u32(sA + g * 2 * tc::kABytes), al = ah + tc::kABytes;
                        const uint32_t d = tbase + (uint32_t)(g * kTcCols);
#pragma unroll
                        for (int pr = 0; pr < 3; ++pr) {
                            const uint32_t a0 = (pr == 1) ? al : ah, b0 = (pr == 2) ? bl : bh;
#pragma unroll
                            for (int ks = 0; ks < tc::kK / 8; ++ks)
                                tc::mma_tf32(d, tc::make_smem_desc(a0 + ks * 2 * tc::kLBO),
                                             tc::make_smem_desc(b0 + ks * 2 * tc::kLBO), idesc, (pr | ks) != 0);
                        }
                        tc::mma_commit(tc::smem_u32(&mbar_full[g]));
                    }
                    if (rnd == 0) p_rows ^= 1u;
                    if (f > 0 || rnd > 0) p_cons ^= 1u;
                }
            }
        }
    } else {
        // ===== environment threads: one thread = one environment = one TMEM lane of its group =====
        const int group = tid >> 7, row = tid & 127;
        const uint32_t my_tmem = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(group * kTcCols);
        unsigned char *myA = sA + group * 2 * tc::kABytes;   // hi tile, then lo tile
        const uint32_t full_bar = tc::smem_u32(&mbar_full[group]), cons_bar = tc::smem_u32(&mbar_cons[group]);
        const uint32_t rows_bar = tc::smem_u32(&mbar_rows[group]);
        uint32_t parity = 0;

        const int e_raw = blockIdx.x * (kTcTiles * 128) + tid;
        const bool active = e_raw < n_envs;
        const int e = active ? e_raw : n_envs - 1;           // idle threads shadow the last env (no stores)

        EnvState s;
        {
            const double2 p = pos[e], v = vel[e];
            const int4 q = ints[e];
            s.px = p.x; s.py = p.y; s.vx = v.x; s.vy = v.y;
            s.k = q.x; s.t = q.y; s.next_gate = q.z; s.passed = q.w;
        }
        float obs[kObsDim];
        {
            const float2 *src = reinterpret_cast<const float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
            for (int i = 0; i < kObsDim / 2; ++i) { const float2 v = src[i]; obs[2 * i] = v.x; obs[2 * i + 1] = v.y; }
        }
        float tc_ = cur_term[e], uc = cur_trunc[e];
        const uint32_t gid = (uint32_t)(env_offset + e);

        // one forward pass of both nets for the observation in `obs`
        auto forward = [&](PolicyOut &po) {
            // this thread's operand row: obs | 1 | 0...  split into TF32 hi and lo.  The previous pass's MMAs have
            // completed (their last "accumulator ready" was awaited), so the tile may be overwritten.
#pragma unroll
            for (int c = 0; c < tc::kKChunks; ++c) {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = 4 * c + j;
                    const float x = k < kObsDim ? obs[k < kObsDim ? k : 0] : (k == kObsDim ? 1.0f : 0.0f);
                    hi[j] = tc::to_tf32(x);
                    lo[j] = tc::to_tf32(x - hi[j]);
                }
                const int off = tc::operand_offset(row, 4 * c);
                *reinterpret_cast<float4 *>(myA + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(myA + tc::kABytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            tc::fence_async_smem();
            tc::mbar_arrive(rows_bar);
            float2 L[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) L[q] = make_float2(0.0f, 0.0f);
            float2 V = make_float2(0.0f, 0.0f);
#pragma unroll 1
            for (int rnd = 0; rnd < 2 * kTcRounds; ++rnd) {
                const int net = rnd / kTcRounds, part = rnd % kTcRounds;   // net 0 = actor, 1 = critic
                tc::mbar_wait(full_bar, parity);
                parity ^= 1u;
                tc::tc_fence_after();
                if (net == 0) {
                    const float2 *w2 = reinterpret_cast<const float2 *>(sw + kTcW2Off) + (size_t)(part * kTcCols) * 5;
#pragma unroll 1
                    for (int c = 0; c < kTcCols; c += 16) {
                        float v[16];
                        tc::tmem_ld16(my_tmem + c, v);
                        if (c + 16 == kTcCols) { tc::tc_fence_before(); tc::mbar_arrive(cons_bar); }   // D_g is free again
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float h = fmaxf(v[i], 0.0f);
#pragma unroll
                            for (int q = 0; q < 5; ++q) L[q] = __ffma2_rn(make_float2(h, h), w2[(c + i) * 5 + q], L[q]);
                        }
                    }
                } else {
                    const float4 *wc = reinterpret_cast<const float4 *>(sw + kTcW2cOff + part * kTcCols);
#pragma unroll 1
                    for (int c = 0; c < kTcCols; c += 16) {
                        float v[16];
                        tc::tmem_ld16(my_tmem + c, v);
                        if (c + 16 == kTcCols) { tc::tc_fence_before(); tc::mbar_arrive(cons_bar); }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 w4 = wc[c / 4 + i];
                            V = __ffma2_rn(make_float2(fmaxf(v[4 * i], 0.0f), fmaxf(v[4 * i + 1], 0.0f)),
                                           make_float2(w4.x, w4.y), V);
                            V = __ffma2_rn(make_float2(fmaxf(v[4 * i + 2], 0.0f), fmaxf(v[4 * i + 3], 0.0f)),
                                           make_float2(w4.z, w4.w), V);
                        }
                    }
                }
            }
            const float *tail = sw + kTcTailOff;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                po.logit[2 * q] = L[q].x + tail[2 * q];
                po.logit[2 * q + 1] = L[q].y + tail[2 * q + 1];
            }
            po.value = (V.x + V.y) + tail[10];
        };

        PolicyOut po;
        for (int t = 0; t < n_steps; ++t) {
            const size_t idx = (size_t)t * (size_t)n_envs + (size_t)e;
            forward(po);
            const unsigned long long gs = step0 + (unsigned long long)t;
            const uint32_t bits = philox_uniform_bits((uint32_t)seed, (uint32_t)(seed >> 32), gid, (uint32_t)gs,
                                                      (uint32_t)(gs >> 32), 0x43415245u);
            const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);
            float logp, us;
            const int a = sample_action(po, u, logp, us);
            if (active) {
                if (obs_mode == kObsPose) {
                    store_pose(reinterpret_cast<PoseRec *>(obs_buf) + idx, s, obs[2], obs[3]);
                } else {
                    float2 *dst = reinterpret_cast<float2 *>(obs_buf + idx * kObsDim);
#pragma unroll
                    for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
                }
                act_buf[idx] = (float)a;
                val_buf[idx] = po.value;
                logp_buf[idx] = logp;
                term_buf[idx] = tc_;
                trunc_buf[idx] = uc;
                if (u_dbg) u_dbg[idx] = u;
            }
            StepResult o;
            env_step<U>(s, a, reward_scale, P, T, o, active ? stats : nullptr);
            if (active) rew_buf[idx] = o.reward;
#pragma unroll
            for (int i = 0; i < kObsDim; ++i) obs[i] = o.obs[i];
            tc_ = o.terminated ? 1.0f : 0.0f;
            uc = o.truncated ? 1.0f : 0.0f;
        }
        if (last_val) forward(po);                           // bootstrap value of the final observation
        if (active) {
            pos[e] = make_double2(s.px, s.py);
            vel[e] = make_double2(s.vx, s.vy);
            ints[e] = make_int4(s.k, s.t, s.next_gate, s.passed);
            float2 *dst = reinterpret_cast<float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
            for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
            cur_term[e] = tc_;
            cur_trunc[e] = uc;
            if (last_val) last_val[e] = po.value;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

// ---- fused rollout, tensor cores, TWO threads per environment ---------------------------------------------------
// k_policy_rollout_tc walks one dependent chain of ~6,900 instructions per environment and step (operand row ->
// MMA -> 256 + 256 hidden units through the CUDA-core second layer -> sampling -> env step); with one 256-env CTA
// per SM (shards of 32,768 envs, BASELINE config 5) that chain IS the step time.  Measured per step (cycle stamps,
// benchmarks/tc2_phase_profile.py): env step 9.4k cycles, second layer 13k — of which 11k are the shared-memory
// pipe delivering the broadcast weights (16 bytes x 32 lanes per LDS.128 = two return cycles; both 128-env groups
// run their second layer at the same time because their MMAs complete together).  Here
//   * every environment has an ENVIRONMENT thread (operand row, Buffer rows, Philox, hidden units 0..127 of the
//     actor, sampling, env step) and a POLICY thread on the same TMEM lane (hidden units 128..255 of the actor —
//     partial logits to the environment thread through shared memory — then the whole critic and the value row,
//     off the critical path: it runs while the environment thread samples and steps);
//   * the second-layer weight loads are software-pipelined by hand (second_layer_chunk16);
//   * the issuer lane is a polling state machine per 128-env group (actor MMAs as two N = 128 halves with their
//     own commit barriers, the critic's N = 256 MMAs as soon as the actor columns are read out), so the groups are
//     independent (option tc_stagger delays group 1's start so that one group's second layer would run beside the
//     other group's env step; measured: no effect on the step time, default 0).
// The critic sums its 256 products in the order of the other kernels (same value bits); the logits are the sum of two
// 128-unit partial chains, so they differ from the other kernels' in the last bits.
// (Second-layer weights from the constant bank as uniform LDCU operands were measured too: 65 cycles per hidden
// unit instead of 27 — the uniform load path delivers one 8-byte operand per ~12 cycles per scheduler.)
// Two hidden units (h0, h1) of the actor through the 256 -> 9 layer: w = their ten weight pairs as five float4
// ([unit][5] float2 layout), L[q] = logits (2q, 2q + 1).
__device__ __forceinline__ void second_layer_pair(float h0, float h1, const float4 (&w)[5], float2 (&L)[5]) {
    L[0] = __ffma2_rn(make_float2(h0, h0), make_float2(w[0].x, w[0].y), L[0]);
    L[1] = __ffma2_rn(make_float2(h0, h0), make_float2(w[0].z, w[0].w), L[1]);
    L[2] = __ffma2_rn(make_float2(h0, h0), make_float2(w[1].x, w[1].y), L[2]);
    L[3] = __ffma2_rn(make_float2(h0, h0), make_float2(w[1].z, w[1].w), L[3]);
    L[4] = __ffma2_rn(make_float2(h0, h0), make_float2(w[2].x, w[2].y), L[4]);
    L[0] = __ffma2_rn(make_float2(h1, h1), make_float2(w[2].z, w[2].w), L[0]);
    L[1] = __ffma2_rn(make_float2(h1, h1), make_float2(w[3].x, w[3].y), L[1]);
    L[2] = __ffma2_rn(make_float2(h1, h1), make_float2(w[3].z, w[3].w), L[2]);
    L[3] = __ffma2_rn(make_float2(h1, h1), make_float2(w[4].x, w[4].y), L[3]);
    L[4] = __ffma2_rn(make_float2(h1, h1), make_float2(w[4].z, w[4].w), L[4]);
}

// Sixteen hidden units (TMEM columns already in v, weights at shared address a = [unit][5] float2).  Two hidden units
// = five 16-byte weight loads + ten FFMA2.  Left to itself the compiler issues each load right before its first use
// (one load in flight: 30 cycles of shared-memory latency per two FFMA2, 71 cycles per hidden unit measured); here
// the loads of the NEXT pair are issued — as ordered volatile loads — before the current pair is consumed.
__device__ __forceinline__ void second_layer_chunk16(uint32_t a, const float (&v)[16], float2 (&L)[5]) {
    float4 wa[5], wb[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) wa[r] = tc::lds128(a + 16u * r);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
#pragma unroll
        for (int r = 0; r < 5; ++r) wb[r] = tc::lds128(a + 80u * (i + 1) + 16u * r);
        second_layer_pair(fmaxf(v[2 * i], 0.0f), fmaxf(v[2 * i + 1], 0.0f), wa, L);
        if (i + 2 < 8) {
#pragma unroll
            for (int r = 0; r < 5; ++r) wa[r] = tc::lds128(a + 80u * (i + 2) + 16u * r);
        }
        second_layer_pair(fmaxf(v[2 * i + 2], 0.0f), fmaxf(v[2 * i + 3], 0.0f), wb, L);
    }
}

#ifdef CARENV_TC2_PROF          // per-phase cycle counts of one environment / policy thread (kernel tuning builds only)
#define TC2_T(i) do { const long long now_ = clock64(); prof_[i] += now_ - last_; last_ = now_; } while (0)
#else
#define TC2_T(i) do { } while (0)
#endif
template <int U>
__global__ void __launch_bounds__(2 * 256 + 32, 1)
k_policy_rollout_tc2(const __grid_constant__ TrackParams P, const Tables G, const float *__restrict__ weights,
                     int n_envs, int n_steps, int env_offset, unsigned long long seed, unsigned long long step0,
                     double2 *__restrict__ pos, double2 *__restrict__ vel, int4 *__restrict__ ints,
                     float *__restrict__ cur_obs, float *__restrict__ cur_term, float *__restrict__ cur_trunc,
                     double reward_scale, float *__restrict__ obs_buf, float *__restrict__ act_buf,
                     float *__restrict__ rew_buf, float *__restrict__ val_buf, float *__restrict__ term_buf,
                     float *__restrict__ trunc_buf, float *__restrict__ logp_buf, float *__restrict__ last_val,
                     float *__restrict__ u_dbg, unsigned long long *stats, int table_bytes, int obs_mode,
                     int stagger_cycles) {
    constexpr int kGroups = 2, kCols = 256, kHalf = 128, kEnvThreads = kGroups * 128;
    extern __shared__ __align__(16) unsigned char smem[];
    // [tables | pad to a 128-byte boundary | weights | A tiles (hi, lo per group) | logits [group][128][12]]
    const int w_off = (int)((tc::smem_u32(smem) + (uint32_t)table_bytes + 127u) / 128u * 128u - tc::smem_u32(smem));
    float *sw = reinterpret_cast<float *>(smem + w_off);
    unsigned char *sA = smem + w_off + ((kTcWeightFloats * 4 + 127) / 128 * 128);
    float *s_logit = reinterpret_cast<float *>(sA + kGroups * 2 * tc::kABytes);
    __shared__ __align__(8) unsigned long long mb_rows[kGroups], mb_full_lo[kGroups], mb_full_hi[kGroups],
        mb_cons_a[kGroups], mb_full_c[kGroups], mb_cons_c[kGroups], mb_logit[kGroups];
    __shared__ uint32_t tmem_slot;
    // the warp index through a shuffle: the compiler then knows that it — and every role branch below — is warp-uniform
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < kTcWeightFloats / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(sw)[i] = reinterpret_cast<const float4 *>(weights)[i];
    if (tid == 0)
        for (int g = 0; g < kGroups; ++g) {
            tc::mbar_init(tc::smem_u32(&mb_rows[g]), 128);
            tc::mbar_init(tc::smem_u32(&mb_full_lo[g]), 1);
            tc::mbar_init(tc::smem_u32(&mb_full_hi[g]), 1);
            tc::mbar_init(tc::smem_u32(&mb_cons_a[g]), 256);
            tc::mbar_init(tc::smem_u32(&mb_full_c[g]), 1);
            tc::mbar_init(tc::smem_u32(&mb_cons_c[g]), 128);
            tc::mbar_init(tc::smem_u32(&mb_logit[g]), 128);
        }
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    const Tables T = stage_tables(G, P.n_gates, P.n_seg, smem);      // ends with __syncthreads()
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const int n_forward = n_steps + (last_val ? 1 : 0);

    if (warp == 2 * kEnvThreads / 32) {
        // ===== MMA issuer (one lane): a polling state machine per group, so the groups never wait for each other =====
        if ((tid & 31) == 0) {
            const uint32_t sB = tc::smem_u32(sw);
            const uint32_t idesc_half = tc::make_idesc_tf32(128, kHalf), idesc_full = tc::make_idesc_tf32(128, kCols);
            int f[kGroups] = {0, 0}, stage[kGroups] = {0, 0};
            int open = n_forward > 0 ? kGroups : 0;
            auto issue = [&](uint32_t d, uint32_t ah, uint32_t bh, uint32_t idesc) {     // D = Ah*Bh + Al*Bh + Ah*Bl
                const uint32_t al = ah + tc::kABytes, bl = bh + (uint32_t)(kTcBFloats * 4);
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) {
                    const uint32_t a0 = (pr == 1) ? al : ah, b0 = (pr == 2) ? bl : bh;
#pragma unroll
                    for (int ks = 0; ks < tc::kK / 8; ++ks)
                        tc::mma_tf32(d, tc::make_smem_desc(a0 + ks * 2 * tc::kLBO), tc::make_smem_desc(b0 + ks * 2 * tc::kLBO),
                                     idesc, (pr | ks) != 0);
                }
            };
            for (long long spin = 0; open > 0; ++spin) {
                if (spin > (1ll << 28)) __trap();                    // a protocol error must not hang the GPU box
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    if (f[g] == n_forward) continue;
                    const uint32_t par = (uint32_t)(f[g] & 1);
                    const uint32_t ah = tc::smem_u32(sA + g * 2 * tc::kABytes);
                    const uint32_t d = tbase + (uint32_t)(g * kCols);
                    if (stage[g] == 0) {
                        // this pass's operand rows are written and the previous pass's critic columns are read out
                        if (!tc::mbar_test(tc::smem_u32(&mb_rows[g]), par)) continue;
                        if (f[g] > 0 && !tc::mbar_test(tc::smem_u32(&mb_cons_c[g]), par ^ 1u)) continue;
                        tc::tc_fence_after();
                        issue(d, ah, sB, idesc_half);                                        // actor units 0..127
                        tc::mma_commit(tc::smem_u32(&mb_full_lo[g]));
                        issue(d + kHalf, ah, sB + (uint32_t)((kHalf / 8) * tc::kSBO), idesc_half);   // actor units 128..255
                        tc::mma_commit(tc::smem_u32(&mb_full_hi[g]));
                        stage[g] = 1;
                    } else {
                        if (!tc::mbar_test(tc::smem_u32(&mb_cons_a[g]), par)) continue;      // the actor columns are read out
                        tc::tc_fence_after();
                        issue(d, ah, sB + (uint32_t)(2 * kTcBFloats * 4), idesc_full);       // critic, 256 units
                        tc::mma_commit(tc::smem_u32(&mb_full_c[g]));
                        stage[g] = 0;
                        if (++f[g] == n_forward) --open;
                    }
                }
            }
        }
    } else {
        const bool policy = warp >= kEnvThreads / 32;
        const int lt = policy ? tid - kEnvThreads : tid;
        const int group = lt >> 7, row = lt & 127;
        const int e_raw = blockIdx.x * kEnvThreads + lt;
        const bool active = e_raw < n_envs;
        const int e = active ? e_raw : n_envs - 1;           // idle threads shadow the last env (no stores)
        float *my_logit = s_logit + (size_t)(group * 128 + row) * 12;
        const uint32_t my_tmem = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(group * kCols);
        const uint32_t w2s = tc::smem_u32(sw + kTcW2Off);            // [unit][5] float2 pairs
        const uint32_t b_cons_a = tc::smem_u32(&mb_cons_a[group]);
        const uint32_t b_logit = tc::smem_u32(&mb_logit[group]), b_full_c = tc::smem_u32(&mb_full_c[group]);

        if (policy) {
            // ===== policy threads: the second layers of both nets for the environment on this TMEM lane =====
            // The per-thread allocation of a 17-warp CTA is 96 registers (five warps on one scheduler's file); the
            // policy threads take 80 (two pairs of hidden units' weights in flight), the environment threads get 112.
            asm volatile("setmaxnreg.dec.sync.aligned.u32 80;\n");
            const uint32_t b_full_hi = tc::smem_u32(&mb_full_hi[group]), b_cons_c = tc::smem_u32(&mb_cons_c[group]);
            const float *tail = sw + kTcTailOff;
            const float4 *wc = reinterpret_cast<const float4 *>(sw + kTcW2cOff);
#ifdef CARENV_TC2_PROF
            long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last_ = clock64();
#endif
            for (int f = 0; f < n_forward; ++f) {
                const uint32_t par = (uint32_t)(f & 1);
                float2 L[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) L[q] = make_float2(0.0f, 0.0f);
                tc::mbar_wait(b_full_hi, par);
                tc::tc_fence_after();
                TC2_T(0);
#pragma unroll 1
                for (int c = kHalf; c < kCols; c += 16) {               // hidden units 128..255
                    float v[16];
                    tc::tmem_ld16(my_tmem + c, v);
                    if (c + 16 == kCols) { tc::tc_fence_before(); tc::mbar_arrive(b_cons_a); }   // this half is read out
                    second_layer_chunk16(w2s + (uint32_t)c * 40u, v, L);
                }
                reinterpret_cast<float4 *>(my_logit)[0] = make_float4(L[0].x, L[0].y, L[1].x, L[1].y);
                reinterpret_cast<float4 *>(my_logit)[1] = make_float4(L[2].x, L[2].y, L[3].x, L[3].y);
                reinterpret_cast<float2 *>(my_logit)[4] = L[4];
                tc::mbar_arrive(b_logit);                            // release: the partial logits are visible
                TC2_T(1);
                tc::mbar_wait(b_full_c, par);
                tc::tc_fence_after();
                TC2_T(2);
                float2 V = make_float2(0.0f, 0.0f);
#pragma unroll 1
                for (int c = 0; c < kCols; c += 16) {
                    float v[16];
                    tc::tmem_ld16(my_tmem + c, v);
                    if (c + 16 == kCols) { tc::tc_fence_before(); tc::mbar_arrive(b_cons_c); }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 w = wc[c / 4 + i];
                        V = __ffma2_rn(make_float2(fmaxf(v[4 * i], 0.0f), fmaxf(v[4 * i + 1], 0.0f)), make_float2(w.x, w.y), V);
                        V = __ffma2_rn(make_float2(fmaxf(v[4 * i + 2], 0.0f), fmaxf(v[4 * i + 3], 0.0f)), make_float2(w.z, w.w), V);
                    }
                }
                const float value = (V.x + V.y) + tail[10];
                if (active) {
                    if (f < n_steps) val_buf[(size_t)f * (size_t)n_envs + (size_t)e] = value;
                    else last_val[e] = value;                        // bootstrap value of the final observation
                }
                TC2_T(3);
            }
#ifdef CARENV_TC2_PROF
            if (blockIdx.x == 0 && lt == 0)
                printf("tc2 policy thread cycles/step: wait_actor %lld actor %lld wait_critic %lld critic %lld\n",
                       prof_[0] / n_forward, prof_[1] / n_forward, prof_[2] / n_forward, prof_[3] / n_forward);
#endif
        } else {
            // ===== environment threads =====
            asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
            unsigned char *myA = sA + group * 2 * tc::kABytes;   // hi tile, then lo tile
            const uint32_t b_rows = tc::smem_u32(&mb_rows[group]), b_full_lo = tc::smem_u32(&mb_full_lo[group]);
            const float *tail = sw + kTcTailOff;
            EnvState s;
            {
                const double2 p = pos[e], v = vel[e];
                const int4 q = ints[e];
                s.px = p.x; s.py = p.y; s.vx = v.x; s.vy = v.y;
                s.k = q.x; s.t = q.y; s.next_gate = q.z; s.passed = q.w;
            }
            float obs[kObsDim];
            {
                const float2 *src = reinterpret_cast<const float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
                for (int i = 0; i < kObsDim / 2; ++i) { const float2 v = src[i]; obs[2 * i] = v.x; obs[2 * i + 1] = v.y; }
            }
            float tc_ = cur_term[e], uc = cur_trunc[e];
            const uint32_t gid = (uint32_t)(env_offset + e);

            // operand row of pass f: obs | 1 | 0...  split into TF32 hi and lo
            auto write_row = [&](int f) {
                if (f > 0) tc::mbar_wait(b_full_c, (uint32_t)((f - 1) & 1));   // the previous pass's MMAs have read the tile
#pragma unroll
                for (int c = 0; c < tc::kKChunks; ++c) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = 4 * c + j;
                        const float x = k < kObsDim ? obs[k < kObsDim ? k : 0] : (k == kObsDim ? 1.0f : 0.0f);
                        hi[j] = tc::to_tf32(x);
                        lo[j] = tc::to_tf32(x - hi[j]);
                    }
                    const int off = tc::operand_offset(row, 4 * c);
                    *reinterpret_cast<float4 *>(myA + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(myA + tc::kABytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                tc::fence_async_smem();
                tc::mbar_arrive(b_rows);
            };

            if (group == 1 && stagger_cycles > 0) {                  // tuning hook (see the header)
                const long long t0 = clock64();
                while (clock64() - t0 < (long long)stagger_cycles) { }
            }
#ifdef CARENV_TC2_PROF
            long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last_ = clock64();
#endif
            for (int t = 0; t < n_steps; ++t) {
                const size_t idx = (size_t)t * (size_t)n_envs + (size_t)e;
                write_row(t);
                TC2_T(0);
                // Philox and the Buffer rows that do not depend on the action, while the tensor core and the policy thread work
                const unsigned long long gs = step0 + (unsigned long long)t;
                const uint32_t bits = philox_uniform_bits((uint32_t)seed, (uint32_t)(seed >> 32), gid, (uint32_t)gs,
                                                          (uint32_t)(gs >> 32), 0x43415245u);
                const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);
                if (active) {
                    if (obs_mode == kObsPose) {
                        store_pose(reinterpret_cast<PoseRec *>(obs_buf) + idx, s, obs[2], obs[3]);
                    } else {
                        float2 *dst = reinterpret_cast<float2 *>(obs_buf + idx * kObsDim);
#pragma unroll
                        for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
                    }
                    term_buf[idx] = tc_;
                    trunc_buf[idx] = uc;
                    if (u_dbg) u_dbg[idx] = u;
                }
                TC2_T(1);
                float2 L[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) L[q] = make_float2(0.0f, 0.0f);
                tc::mbar_wait(b_full_lo, (uint32_t)(t & 1));
                tc::tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < kHalf; c += 16) {                   // hidden units 0..127
                    float v[16];
                    tc::tmem_ld16(my_tmem + c, v);
                    if (c + 16 == kHalf) { tc::tc_fence_before(); tc::mbar_arrive(b_cons_a); }   // this half is read out
                    second_layer_chunk16(w2s + (uint32_t)c * 40u, v, L);
                }
                TC2_T(5);
                tc::mbar_wait(b_logit, (uint32_t)(t & 1));           // acquire: the policy thread's partial logits
                TC2_T(2);
                PolicyOut po;
                {
                    const float4 p0 = reinterpret_cast<const float4 *>(my_logit)[0], p1 = reinterpret_cast<const float4 *>(my_logit)[1];
                    const float2 p2 = reinterpret_cast<const float2 *>(my_logit)[4];
                    po.logit[0] = (L[0].x + p0.x) + tail[0]; po.logit[1] = (L[0].y + p0.y) + tail[1];
                    po.logit[2] = (L[1].x + p0.z) + tail[2]; po.logit[3] = (L[1].y + p0.w) + tail[3];
                    po.logit[4] = (L[2].x + p1.x) + tail[4]; po.logit[5] = (L[2].y + p1.y) + tail[5];
                    po.logit[6] = (L[3].x + p1.z) + tail[6]; po.logit[7] = (L[3].y + p1.w) + tail[7];
                    po.logit[8] = (L[4].x + p2.x) + tail[8]; po.logit[9] = (L[4].y + p2.y) + tail[9];
                    po.value = 0.0f;
                }
                float logp, us;
                const int a = sample_action(po, u, logp, us);
                if (active) {
                    act_buf[idx] = (float)a;
                    logp_buf[idx] = logp;
                }
                TC2_T(3);
                StepResult o;
                env_step<U>(s, a, reward_scale, P, T, o, active ? stats : nullptr);
                if (active) rew_buf[idx] = o.reward;
#pragma unroll
                for (int i = 0; i < kObsDim; ++i) obs[i] = o.obs[i];
                tc_ = o.terminated ? 1.0f : 0.0f;
                uc = o.truncated ? 1.0f : 0.0f;
                TC2_T(4);
            }
#ifdef CARENV_TC2_PROF
            if (blockIdx.x == 0 && lt == 0)
                printf("tc2 env thread cycles/step: row %lld philox+stores %lld wait+actor_lo %lld wait_partial %lld sample %lld env_step %lld\n",
                       prof_[0] / n_steps, prof_[1] / n_steps, prof_[5] / n_steps, prof_[2] / n_steps, prof_[3] / n_steps, prof_[4] / n_steps);
#endif
            if (last_val) {                                          // one more pass: the policy thread writes the bootstrap value
                write_row(n_steps);
                tc::mbar_wait(b_full_lo, (uint32_t)(n_steps & 1));
                tc::tc_fence_before();
                tc::mbar_arrive(b_cons_a);
            }
            if (active) {
                pos[e] = make_double2(s.px, s.py);
                vel[e] = make_double2(s.vx, s.vy);
                ints[e] = make_int4(s.k, s.t, s.next_gate, s.passed);
                float2 *dst = reinterpret_cast<float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
                for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
                cur_term[e] = tc_;
                cur_trunc[e] = uc;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

// ---- fused rollout, tensor cores, every second-layer weight load shared by TWO environments --------------------
// ncu of k_policy_rollout_tc2 (profiles/r2_policy_tc2_ncu.txt): the 256 -> 9 layer is bound by the shared-memory
// RETURN path — a broadcast LDS.128 delivers 16 bytes to 32 lanes in two wavefronts, i.e. the weights of ONE packed
// FFMA2 per cycle and SM, half of what the FMA pipes take — and all warps run that phase together because their
// MMAs complete together (the pipe is ~100 % busy for 11k of a step's 29k cycles, 41 % on average).  Here every
// weight load feeds two environments: the environment thread of lane r in group 0 runs hidden units 0..127 for
// BOTH environments on TMEM lane r (group 0's and group 1's columns), its twin in group 1 runs units 128..255 for
// both, and the two exchange the partial logits of each other's environment through shared memory.  The critic is
// done by 128 critic threads, lane r for both environments of the lane as well, off the critical path.  One CTA =
// 256 environment threads + 128 critic threads + the issuer warp (13 warps, 128 registers each at launch; the critic
// warps hand 64 of theirs to the environment warps).  Logits = (units 0..127) + (units 128..255) + bias, the order
// of k_policy_rollout_tc2: bit-identical to it; values bit-identical to every other kernel.
__device__ __forceinline__ void second_layer_chunk16_x2(uint32_t a, const float (&va)[16], const float (&vb)[16],
                                                        float2 (&La)[5], float2 (&Lb)[5]) {
    float4 wa[5], wb[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) wa[r] = tc::lds128(a + 16u * r);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
#pragma unroll
        for (int r = 0; r < 5; ++r) wb[r] = tc::lds128(a + 80u * (i + 1) + 16u * r);
        second_layer_pair(fmaxf(va[2 * i], 0.0f), fmaxf(va[2 * i + 1], 0.0f), wa, La);
        second_layer_pair(fmaxf(vb[2 * i], 0.0f), fmaxf(vb[2 * i + 1], 0.0f), wa, Lb);
        if (i + 2 < 8) {
#pragma unroll
            for (int r = 0; r < 5; ++r) wa[r] = tc::lds128(a + 80u * (i + 2) + 16u * r);
        }
        second_layer_pair(fmaxf(va[2 * i + 2], 0.0f), fmaxf(va[2 * i + 3], 0.0f), wb, La);
        second_layer_pair(fmaxf(vb[2 * i + 2], 0.0f), fmaxf(vb[2 * i + 3], 0.0f), wb, Lb);
    }
}

template <int U, bool TAB>
__global__ void __launch_bounds__(256 + 128 + 32, 1)
k_policy_rollout_tc3(const __grid_constant__ TrackParams P, const Tables G, const float *__restrict__ weights,
                     int n_envs, int n_steps, int env_offset, unsigned long long seed, unsigned long long step0,
                     double2 *__restrict__ pos, double2 *__restrict__ vel, int4 *__restrict__ ints,
                     float *__restrict__ cur_obs, float *__restrict__ cur_term, float *__restrict__ cur_trunc,
                     double reward_scale, float *__restrict__ obs_buf, float *__restrict__ act_buf,
                     float *__restrict__ rew_buf, float *__restrict__ val_buf, float *__restrict__ term_buf,
                     float *__restrict__ trunc_buf, float *__restrict__ logp_buf, float *__restrict__ last_val,
                     float *__restrict__ u_dbg, unsigned long long *stats, int table_bytes, int obs_mode,
                     const float4 *__restrict__ den4, int n_pairs, int row_f4) {
    constexpr int kGroups = 2, kCols = 256, kHalf = 128, kEnvThreads = kGroups * 128, kCriticThreads = 128;
    extern __shared__ __align__(16) unsigned char smem[];
    // [tables | pad to a 128-byte boundary | weights | A tiles (hi, lo per group) | exchanged partial logits [group][128][12]
    //  | TAB: the denominator table of k_rollout_tab, ONE copy whose rows are skewed by 16 bytes (row_f4 is odd in
    //    16-byte units): threads of a quarter warp with different headings mostly hit different bank groups (about
    //    2.6 instead of 1 wavefront per quarter — there is no room for the eight conflict-free copies) — this replaces
    //    the 144 x (DMUL + DFMA + F2F) per env-step that the arithmetic kernels spend on the same values]
    const int w_off = (int)((tc::smem_u32(smem) + (uint32_t)table_bytes + 127u) / 128u * 128u - tc::smem_u32(smem));
    float *sw = reinterpret_cast<float *>(smem + w_off);
    unsigned char *sA = smem + w_off + ((kTcWeightFloats * 4 + 127) / 128 * 128);
    float *s_xchg = reinterpret_cast<float *>(sA + kGroups * 2 * tc::kABytes);
    float4 *s_tab = reinterpret_cast<float4 *>(s_xchg + kGroups * 128 * 12);
    if (TAB)
        for (int i = threadIdx.x; i < kHeadings * n_pairs; i += blockDim.x) {
            const int k = i / n_pairs, jp = i - k * n_pairs;
            s_tab[k * row_f4 + jp] = den4[i];
        }
    const TabView tv{s_tab, row_f4};
    __shared__ __align__(8) unsigned long long mb_rows, mb_full_lo, mb_full_hi, mb_cons_a, mb_full_c, mb_cons_c, mb_xchg;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < kTcWeightFloats / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(sw)[i] = reinterpret_cast<const float4 *>(weights)[i];
    if (tid == 0) {
        tc::mbar_init(tc::smem_u32(&mb_rows), kEnvThreads);
        tc::mbar_init(tc::smem_u32(&mb_full_lo), 1);
        tc::mbar_init(tc::smem_u32(&mb_full_hi), 1);
        tc::mbar_init(tc::smem_u32(&mb_cons_a), kEnvThreads);
        tc::mbar_init(tc::smem_u32(&mb_full_c), 1);
        tc::mbar_init(tc::smem_u32(&mb_cons_c), kCriticThreads);
        tc::mbar_init(tc::smem_u32(&mb_xchg), kEnvThreads);
    }
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    const Tables T = stage_tables(G, P.n_gates, P.n_seg, smem);      // ends with __syncthreads()
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const int n_forward = n_steps + (last_val ? 1 : 0);
    const uint32_t b_rows = tc::smem_u32(&mb_rows), b_full_lo = tc::smem_u32(&mb_full_lo), b_full_hi = tc::smem_u32(&mb_full_hi);
    const uint32_t b_cons_a = tc::smem_u32(&mb_cons_a), b_full_c = tc::smem_u32(&mb_full_c), b_cons_c = tc::smem_u32(&mb_cons_c);
    const uint32_t b_xchg = tc::smem_u32(&mb_xchg);

    if (warp == (kEnvThreads + kCriticThreads) / 32) {
        // ===== MMA issuer (one lane); both groups advance together (their threads share the weight loads) =====
        if ((tid & 31) == 0) {
            const uint32_t sB = tc::smem_u32(sw);
            const uint32_t idesc_half = tc::make_idesc_tf32(128, kHalf), idesc_full = tc::make_idesc_tf32(128, kCols);
            auto issue = [&](uint32_t d, uint32_t ah, uint32_t bh, uint32_t idesc) {     // D = Ah*Bh + Al*Bh + Ah*Bl
                const uint32_t al = ah + tc::kABytes, bl = bh + (uint32_t)(kTcBFloats * 4);
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) {
                    const uint32_t a0 = (pr == 1) ? al : ah, b0 = (pr == 2) ? bl : bh;
#pragma unroll
                    for (int ks = 0; ks < tc::kK / 8; ++ks)
                        tc::mma_tf32(d, tc::make_smem_desc(a0 + ks * 2 * tc::kLBO), tc::make_smem_desc(b0 + ks * 2 * tc::kLBO),
                                     idesc, (pr | ks) != 0);
                }
            };
            const uint32_t a0 = tc::smem_u32(sA), a1 = a0 + 2 * tc::kABytes;
            const uint32_t b_hi_half = sB + (uint32_t)((kHalf / 8) * tc::kSBO), b_critic = sB + (uint32_t)(2 * kTcBFloats * 4);
            for (int f = 0; f < n_forward; ++f) {
                const uint32_t par = (uint32_t)(f & 1);
                tc::mbar_wait(b_rows, par);                          // this pass's operand rows are written
                if (f > 0) tc::mbar_wait(b_cons_c, par ^ 1u);        // the previous pass's critic columns are read out
                tc::tc_fence_after();
                issue(tbase, a0, sB, idesc_half);                    // actor units 0..127, both groups
                issue(tbase + kCols, a1, sB, idesc_half);
                tc::mma_commit(b_full_lo);
                issue(tbase + kHalf, a0, b_hi_half, idesc_half);     // actor units 128..255, both groups
                issue(tbase + kCols + kHalf, a1, b_hi_half, idesc_half);
                tc::mma_commit(b_full_hi);
                tc::mbar_wait(b_cons_a, par);                        // the actor columns are read out
                tc::tc_fence_after();
                issue(tbase, a0, b_critic, idesc_full);              // critic, 256 units, both groups
                issue(tbase + kCols, a1, b_critic, idesc_full);
                tc::mma_commit(b_full_c);
            }
        }
    } else if (warp >= kEnvThreads / 32) {
        // ===== critic threads: lane r's two environments, every weight load used twice =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
        const int row = tid - kEnvThreads;
        const uint32_t lane_tmem = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const int ea_raw = blockIdx.x * kEnvThreads + row, eb_raw = ea_raw + 128;
        const float4 *wc = reinterpret_cast<const float4 *>(sw + kTcW2cOff);
        const float bias = sw[kTcTailOff + 10];
        for (int f = 0; f < n_forward; ++f) {
            tc::mbar_wait(b_full_c, (uint32_t)(f & 1));
            tc::tc_fence_after();
            float2 Va = make_float2(0.0f, 0.0f), Vb = make_float2(0.0f, 0.0f);
#pragma unroll 1
            for (int c = 0; c < kCols; c += 16) {
                float va[16], vb[16];
                tc::tmem_ld16(lane_tmem + c, va);
                tc::tmem_ld16(lane_tmem + kCols + c, vb);
                if (c + 16 == kCols) { tc::tc_fence_before(); tc::mbar_arrive(b_cons_c); }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 w = wc[c / 4 + i];
                    Va = __ffma2_rn(make_float2(fmaxf(va[4 * i], 0.0f), fmaxf(va[4 * i + 1], 0.0f)), make_float2(w.x, w.y), Va);
                    Vb = __ffma2_rn(make_float2(fmaxf(vb[4 * i], 0.0f), fmaxf(vb[4 * i + 1], 0.0f)), make_float2(w.x, w.y), Vb);
                    Va = __ffma2_rn(make_float2(fmaxf(va[4 * i + 2], 0.0f), fmaxf(va[4 * i + 3], 0.0f)), make_float2(w.z, w.w), Va);
                    Vb = __ffma2_rn(make_float2(fmaxf(vb[4 * i + 2], 0.0f), fmaxf(vb[4 * i + 3], 0.0f)), make_float2(w.z, w.w), Vb);
                }
            }
            const float value_a = (Va.x + Va.y) + bias, value_b = (Vb.x + Vb.y) + bias;
            if (f < n_steps) {
                if (ea_raw < n_envs) val_buf[(size_t)f * (size_t)n_envs + (size_t)ea_raw] = value_a;
                if (eb_raw < n_envs) val_buf[(size_t)f * (size_t)n_envs + (size_t)eb_raw] = value_b;
            } else {                                                 // bootstrap values of the final observations
                if (ea_raw < n_envs) last_val[ea_raw] = value_a;
                if (eb_raw < n_envs) last_val[eb_raw] = value_b;
            }
        }
    } else {
        // ===== environment threads =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 160;\n");
        const int group = tid >> 7, row = tid & 127;
        const int e_raw = blockIdx.x * kEnvThreads + tid;
        const bool active = e_raw < n_envs;
        const int e = active ? e_raw : n_envs - 1;           // idle threads shadow the last env (no stores)
        const uint32_t lane_tmem = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t w2s = tc::smem_u32(sw + kTcW2Off) + (uint32_t)(group * kHalf) * 40u;   // this thread's 128 hidden units
        const uint32_t b_full_mine = group == 0 ? b_full_lo : b_full_hi;
        float *xchg_out = s_xchg + (size_t)((1 - group) * 128 + row) * 12;    // partial logits of the OTHER group's lane-r env
        const float *xchg_in = s_xchg + (size_t)(group * 128 + row) * 12;
        unsigned char *myA = sA + group * 2 * tc::kABytes;   // hi tile, then lo tile
        const float *tail = sw + kTcTailOff;
        EnvState s;
        {
            const double2 p = pos[e], v = vel[e];
            const int4 q = ints[e];
            s.px = p.x; s.py = p.y; s.vx = v.x; s.vy = v.y;
            s.k = q.x; s.t = q.y; s.next_gate = q.z; s.passed = q.w;
        }
        float obs[kObsDim];
        {
            const float2 *src = reinterpret_cast<const float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
            for (int i = 0; i < kObsDim / 2; ++i) { const float2 v = src[i]; obs[2 * i] = v.x; obs[2 * i + 1] = v.y; }
        }
        float tc_ = cur_term[e], uc = cur_trunc[e];
        const uint32_t gid = (uint32_t)(env_offset + e);

        auto write_row = [&](int f) {
            if (f > 0) tc::mbar_wait(b_full_c, (uint32_t)((f - 1) & 1));   // the previous pass's MMAs have read the tile
#pragma unroll
            for (int c = 0; c < tc::kKChunks; ++c) {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = 4 * c + j;
                    const float x = k < kObsDim ? obs[k < kObsDim ? k : 0] : (k == kObsDim ? 1.0f : 0.0f);
                    hi[j] = tc::to_tf32(x);
                    lo[j] = tc::to_tf32(x - hi[j]);
                }
                const int off = tc::operand_offset(row, 4 * c);
                *reinterpret_cast<float4 *>(myA + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(myA + tc::kABytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            tc::fence_async_smem();
            tc::mbar_arrive(b_rows);
        };

        for (int t = 0; t < n_steps; ++t) {
            const size_t idx = (size_t)t * (size_t)n_envs + (size_t)e;
            const uint32_t par = (uint32_t)(t & 1);
            write_row(t);
            const unsigned long long gs = step0 + (unsigned long long)t;
            const uint32_t bits = philox_uniform_bits((uint32_t)seed, (uint32_t)(seed >> 32), gid, (uint32_t)gs,
                                                      (uint32_t)(gs >> 32), 0x43415245u);
            const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);
            if (active) {
                if (obs_mode == kObsPose) {
                    store_pose(reinterpret_cast<PoseRec *>(obs_buf) + idx, s, obs[2], obs[3]);
                } else {
                    float2 *dst = reinterpret_cast<float2 *>(obs_buf + idx * kObsDim);
#pragma unroll
                    for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
                }
                term_buf[idx] = tc_;
                trunc_buf[idx] = uc;
                if (u_dbg) u_dbg[idx] = u;
            }
            // my 128 hidden units for the two environments of TMEM lane r (a = group 0's, b = group 1's)
            float2 La[5], Lb[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) { La[q] = make_float2(0.0f, 0.0f); Lb[q] = make_float2(0.0f, 0.0f); }
            tc::mbar_wait(b_full_mine, par);
            tc::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < kHalf; c += 16) {
                float va[16], vb[16];
                tc::tmem_ld16(lane_tmem + group * kHalf + c, va);
                tc::tmem_ld16(lane_tmem + kCols + group * kHalf + c, vb);
                if (c + 16 == kHalf) { tc::tc_fence_before(); tc::mbar_arrive(b_cons_a); }   // my columns are read out
                second_layer_chunk16_x2(w2s + (uint32_t)c * 40u, va, vb, La, Lb);
            }
            {   // hand the other environment's partial logits to its thread, take mine
                const float2 *o = group == 0 ? Lb : La;
                reinterpret_cast<float4 *>(xchg_out)[0] = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
                reinterpret_cast<float4 *>(xchg_out)[1] = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
                reinterpret_cast<float2 *>(xchg_out)[4] = o[4];
            }
            tc::mbar_arrive(b_xchg);
            tc::mbar_wait(b_xchg, par);
            PolicyOut po;
            {
                const float2 *m = group == 0 ? La : Lb;
                const float4 p0 = reinterpret_cast<const float4 *>(xchg_in)[0], p1 = reinterpret_cast<const float4 *>(xchg_in)[1];
                const float2 p2 = reinterpret_cast<const float2 *>(xchg_in)[4];
                const float r[10] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y};
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    // (units 0..127) + (units 128..255): the same operands in the same order on both threads
                    const float lo0 = group == 0 ? m[q].x : r[2 * q], hi0 = group == 0 ? r[2 * q] : m[q].x;
                    const float lo1 = group == 0 ? m[q].y : r[2 * q + 1], hi1 = group == 0 ? r[2 * q + 1] : m[q].y;
                    po.logit[2 * q] = (lo0 + hi0) + tail[2 * q];
                    po.logit[2 * q + 1] = (lo1 + hi1) + tail[2 * q + 1];
                }
                po.value = 0.0f;
            }
            float logp, us;
            const int a = sample_action(po, u, logp, us);
            if (active) {
                act_buf[idx] = (float)a;
                logp_buf[idx] = logp;
            }
            StepResult o;
            env_step<U, TAB>(s, a, reward_scale, P, T, o, active ? stats : nullptr, nullptr, TAB ? &tv : nullptr);
            if (active) rew_buf[idx] = o.reward;
#pragma unroll
            for (int i = 0; i < kObsDim; ++i) obs[i] = o.obs[i];
            tc_ = o.terminated ? 1.0f : 0.0f;
            uc = o.truncated ? 1.0f : 0.0f;
        }
        if (last_val) {                                          // one more pass: the critic threads write the bootstrap values
            write_row(n_steps);
            tc::mbar_wait(b_full_mine, (uint32_t)(n_steps & 1));
            tc::tc_fence_before();
            tc::mbar_arrive(b_cons_a);
        }
        if (active) {
            pos[e] = make_double2(s.px, s.py);
            vel[e] = make_double2(s.vx, s.vy);
            ints[e] = make_int4(s.k, s.t, s.next_gate, s.passed);
            float2 *dst = reinterpret_cast<float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
            for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
            cur_term[e] = tc_;
            cur_trunc[e] = uc;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

// ---- fused rollout for small batches: one WARP per environment --------------------------------------------------
// The reference's own training shape is 24 environments (README: "about 35 minutes" for 200 epochs).  There a
// thread-per-environment kernel is one long dependent chain on a nearly empty GPU (12 us per step in
// k_policy_rollout_tc3).  Here a warp owns one environment, as in k_rollout_warp: lane j carries wall segment j
// through the env step (per-ray extrema by integer REDUX) and, for the policy, hidden units j, j + 32, ... of both
// nets (16 units per lane: first layer from shared-memory rows [unit][20] — conflict-free for consecutive lanes —,
// ReLU, its share of the 256 -> 9 / 256 -> 1 sums), the ten partial sums are folded over the warp by an xor
// butterfly (every lane ends with the same bits), every lane samples the same action.  Float32 CUDA cores only:
// at 24 environments the tensor core has nothing to amortise its latency over.  The network parameters are read in
// nn.Linear layout (no packing launch).
constexpr int kWpRow = 20;                                   // first-layer row: 18 weights, bias, pad
constexpr int kWpW2 = 12;                                    // actor second-layer row: W2[0..8][j], pad
constexpr int kWpFloats = 2 * kHidden * kWpRow + kHidden * kWpW2 + kHidden + 12;   // 13,580 floats = 54,320 B

__global__ void __launch_bounds__(128)
k_policy_rollout_warp(const __grid_constant__ TrackParams P, const Tables G, const ppo::Params W, int n_envs, int n_steps,
                      int env_offset, unsigned long long seed, unsigned long long step0, double2 *__restrict__ pos,
                      double2 *__restrict__ vel, int4 *__restrict__ ints, float *__restrict__ cur_obs,
                      float *__restrict__ cur_term, float *__restrict__ cur_trunc, double reward_scale,
                      float *__restrict__ obs_buf, float *__restrict__ act_buf, float *__restrict__ rew_buf,
                      float *__restrict__ val_buf, float *__restrict__ term_buf, float *__restrict__ trunc_buf,
                      float *__restrict__ logp_buf, float *__restrict__ last_val, float *__restrict__ u_dbg,
                      unsigned long long *stats, int table_bytes, int obs_mode) {
    extern __shared__ __align__(16) unsigned char smem[];
    float *sw1 = reinterpret_cast<float *>(smem + table_bytes);   // [512][20]: rows 0..255 actor, 256..511 critic
    float *sw2 = sw1 + 2 * kHidden * kWpRow;                      // [256][12]
    float *swc = sw2 + kHidden * kWpW2;                           // [256]
    float *stail = swc + kHidden;                                 // b2a[0..8], 0, b2c, 0
    for (int i = threadIdx.x; i < 2 * kHidden * kWpRow; i += blockDim.x) {
        const int r = i / kWpRow, c = i - r * kWpRow, j = r & (kHidden - 1);
        const float *w1 = r < kHidden ? W.w1a : W.w1c, *b1 = r < kHidden ? W.b1a : W.b1c;
        sw1[i] = c < kObsDim ? w1[j * kObsDim + c] : (c == kObsDim ? b1[j] : 0.0f);
    }
    for (int i = threadIdx.x; i < kHidden * kWpW2; i += blockDim.x) {
        const int j = i / kWpW2, q = i - j * kWpW2;
        sw2[i] = q < kActions ? W.w2a[q * kHidden + j] : 0.0f;
    }
    for (int i = threadIdx.x; i < kHidden; i += blockDim.x) swc[i] = W.w2c[i];
    if (threadIdx.x < 12) stail[threadIdx.x] = threadIdx.x < kActions ? W.b2a[threadIdx.x] : (threadIdx.x == 10 ? W.b2c[0] : 0.0f);
    const Tables T = stage_tables(G, P.n_gates, P.n_seg, smem);      // ends with __syncthreads()
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= n_envs) return;

    WarpSeg ws;
    ws.active = lane < P.n_seg;
    ws.f = G.segf[ws.active ? lane : 0];
    ws.g = G.segd[ws.active ? lane : 0];
    EnvState s;
    {
        const double2 p = pos[e], v = vel[e];
        const int4 q = ints[e];
        s.px = p.x; s.py = p.y; s.vx = v.x; s.vy = v.y;
        s.k = q.x; s.t = q.y; s.next_gate = q.z; s.passed = q.w;
    }
    float obs[kObsDim];
    {
        const float2 *src = reinterpret_cast<const float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
        for (int i = 0; i < kObsDim / 2; ++i) { const float2 v = src[i]; obs[2 * i] = v.x; obs[2 * i + 1] = v.y; }
    }
    float tc_ = cur_term[e], uc = cur_trunc[e];
    const uint32_t gid = (uint32_t)(env_offset + e);
    unsigned long long *my_stats = lane == 0 ? stats : nullptr;

    // pre-activation of one hidden unit (row of sw1) for the observation in `obs`
    auto unit = [&](const float *row) {
        const float4 *r4 = reinterpret_cast<const float4 *>(row);
        const float4 w0 = r4[0], w1 = r4[1], w2 = r4[2], w3 = r4[3], w4 = r4[4];
        float pa = w4.z, pb = 0.0f, pc = 0.0f;                                   // bias; three chains
        pa = fmaf(w0.x, obs[0], pa); pb = fmaf(w0.y, obs[1], pb); pc = fmaf(w0.z, obs[2], pc);
        pa = fmaf(w0.w, obs[3], pa); pb = fmaf(w1.x, obs[4], pb); pc = fmaf(w1.y, obs[5], pc);
        pa = fmaf(w1.z, obs[6], pa); pb = fmaf(w1.w, obs[7], pb); pc = fmaf(w2.x, obs[8], pc);
        pa = fmaf(w2.y, obs[9], pa); pb = fmaf(w2.z, obs[10], pb); pc = fmaf(w2.w, obs[11], pc);
        pa = fmaf(w3.x, obs[12], pa); pb = fmaf(w3.y, obs[13], pb); pc = fmaf(w3.z, obs[14], pc);
        pa = fmaf(w3.w, obs[15], pa); pb = fmaf(w4.x, obs[16], pb); pc = fmaf(w4.y, obs[17], pc);
        return fmaxf((pa + pb) + pc, 0.0f);
    };
    // both nets for `obs`: logits (+ bias) and value in every lane
    auto forward = [&](PolicyOut &po, bool actor) {
        float part[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) part[q] = 0.0f;
#pragma unroll 2
        for (int i = 0; i < kHidden / 32; ++i) {
            const int j = lane + 32 * i;
            if (actor) {
                const float h = unit(sw1 + j * kWpRow);
                const float4 *v4 = reinterpret_cast<const float4 *>(sw2 + j * kWpW2);
                const float4 v0 = v4[0], v1 = v4[1], v2 = v4[2];
                part[0] = fmaf(v0.x, h, part[0]); part[1] = fmaf(v0.y, h, part[1]); part[2] = fmaf(v0.z, h, part[2]);
                part[3] = fmaf(v0.w, h, part[3]); part[4] = fmaf(v1.x, h, part[4]); part[5] = fmaf(v1.y, h, part[5]);
                part[6] = fmaf(v1.z, h, part[6]); part[7] = fmaf(v1.w, h, part[7]); part[8] = fmaf(v2.x, h, part[8]);
            }
            part[9] = fmaf(swc[j], unit(sw1 + (kHidden + j) * kWpRow), part[9]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 10; ++q) part[q] += __shfl_xor_sync(0xffffffffu, part[q], o);
        }
#pragma unroll
        for (int q = 0; q < kActions; ++q) po.logit[q] = part[q] + stail[q];
        po.logit[9] = 0.0f;
        po.value = part[9] + stail[10];
    };

    PolicyOut po;
    for (int t = 0; t < n_steps; ++t) {
        const size_t idx = (size_t)t * (size_t)n_envs + (size_t)e;
        forward(po, true);
        const unsigned long long gs = step0 + (unsigned long long)t;
        const uint32_t bits = philox_uniform_bits((uint32_t)seed, (uint32_t)(seed >> 32), gid, (uint32_t)gs,
                                                  (uint32_t)(gs >> 32), 0x43415245u);
        const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);
        float logp, us;
        const int a = sample_action(po, u, logp, us);
        if (lane == 0) {
            if (obs_mode == kObsPose) {
                store_pose(reinterpret_cast<PoseRec *>(obs_buf) + idx, s, obs[2], obs[3]);
            } else {
                float2 *dst = reinterpret_cast<float2 *>(obs_buf + idx * kObsDim);
#pragma unroll
                for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
            }
            act_buf[idx] = (float)a;
            val_buf[idx] = po.value;
            logp_buf[idx] = logp;
            term_buf[idx] = tc_;
            trunc_buf[idx] = uc;
            if (u_dbg) u_dbg[idx] = u;
        }
        StepResult o;
        env_step<kWarpPerEnv>(s, a, reward_scale, P, T, o, my_stats, &ws);
        if (lane == 0) rew_buf[idx] = o.reward;
#pragma unroll
        for (int i = 0; i < kObsDim; ++i) obs[i] = o.obs[i];
        tc_ = o.terminated ? 1.0f : 0.0f;
        uc = o.truncated ? 1.0f : 0.0f;
    }
    if (last_val) forward(po, false);                        // bootstrap value of the final observation (critic only)
    if (lane == 0) {
        pos[e] = make_double2(s.px, s.py);
        vel[e] = make_double2(s.vx, s.vy);
        ints[e] = make_int4(s.k, s.t, s.next_gate, s.passed);
        float2 *dst = reinterpret_cast<float2 *>(cur_obs + (size_t)e * kObsDim);
#pragma unroll
        for (int i = 0; i < kObsDim / 2; ++i) dst[i] = make_float2(obs[2 * i], obs[2 * i + 1]);
        cur_term[e] = tc_;
        cur_trunc[e] = uc;
        if (last_val) last_val[e] = po.value;
    }
}
