// TEST HARNESS — compiles the product's per-environment arithmetic (ppo_car_b200/csrc/carenv_core.cuh)
// for the HOST so that the kernel's float32 fast path, guard bands and float64 fallbacks can be
// replayed on a CPU-only machine against the float64 oracle.  Not part of the product; the product
// library has no CPU path.
#include <stdint.h>

#include "../../ppo_car_b200/csrc/carenv_tables.h"

using namespace carenv;

extern "C" int emul_rollout(const double *walls, int n_walls, const double *gates, int n_gates, double sx, double sy,
                            double angle, int n_envs, int env_lo, int env_hi, int T, const uint8_t *actions,
                            double reward_scale, double *state_pv, int32_t *state_i, int do_reset, int unrolled, float *reset_obs,
                            float *obs, float *rew, uint8_t *term, uint8_t *trunc, int32_t *info,
                            unsigned long long *stats) {
    HostTrack H;
    if (build_host_track(walls, n_walls, gates, n_gates, sx, sy, angle, H) != 0) return -1;
    const Tables Tb = H.tables();
    // unrolled == 2: the table variant of the pair path (k_rollout_tab), reading the host's den4 table directly
    const TabView tv{reinterpret_cast<const float4 *>(H.den4.data()), H.n_pairs};
    const bool tab = unrolled == 2 && H.n_pairs > 0 && H.P.unroll >= 2;
    if (reset_obs) for (int i = 0; i < kObsDim; ++i) reset_obs[i] = H.P.reset_obs[i];
    for (int e = env_lo; e < env_hi; ++e) {
        EnvState s;
        if (do_reset) {
            s = EnvState{sx, sy, 0.0, 0.0, 0, 0, 0, 0};
        } else {
            s = EnvState{state_pv[4 * e], state_pv[4 * e + 1], state_pv[4 * e + 2], state_pv[4 * e + 3],
                         state_i[4 * e], state_i[4 * e + 1], state_i[4 * e + 2], state_i[4 * e + 3]};
        }
        for (int t = 0; t < T; ++t) {
            const size_t k = (size_t)t * n_envs + e;
            StepResult o;
            const int U = H.P.n_seg > kMaxSeg ? 0 : (unrolled ? H.P.unroll : 1);
            if (tab && U == 6) env_step<6, true>(s, actions[k], reward_scale, H.P, Tb, o, stats, nullptr, &tv);
            else if (tab && U == 4) env_step<4, true>(s, actions[k], reward_scale, H.P, Tb, o, stats, nullptr, &tv);
            else if (tab && U == 2) env_step<2, true>(s, actions[k], reward_scale, H.P, Tb, o, stats, nullptr, &tv);
            else if (U == 0) env_step<0>(s, actions[k], reward_scale, H.P, Tb, o, stats);
            else if (U == 6) env_step<6>(s, actions[k], reward_scale, H.P, Tb, o, stats);
            else if (U == 4) env_step<4>(s, actions[k], reward_scale, H.P, Tb, o, stats);
            else if (U == 2) env_step<2>(s, actions[k], reward_scale, H.P, Tb, o, stats);
            else env_step<1>(s, actions[k], reward_scale, H.P, Tb, o, stats);
            if (obs) for (int i = 0; i < kObsDim; ++i) obs[k * kObsDim + i] = o.obs[i];
            if (rew) rew[k] = o.reward;
            if (term) term[k] = (uint8_t)o.terminated;
            if (trunc) trunc[k] = (uint8_t)o.truncated;
            if (info) {
                info[4 * k] = o.gates_passed; info[4 * k + 1] = o.time_passed; info[4 * k + 2] = o.next_gate;
                info[4 * k + 3] = o.gate_hit | (o.lap << 1);
            }
        }
        state_pv[4 * e] = s.px; state_pv[4 * e + 1] = s.py; state_pv[4 * e + 2] = s.vx; state_pv[4 * e + 3] = s.vy;
        state_i[4 * e] = s.k; state_i[4 * e + 1] = s.t; state_i[4 * e + 2] = s.next_gate; state_i[4 * e + 3] = s.passed;
    }
    return 0;
}

// The 72-entry float64 heading table (cos, sin of initial_angle + 5k degrees) exactly as the product builds it.
extern "C" int emul_heading_table(double angle, double *cos_sin_out /* [72][2] */) {
    const double walls[4] = {0.0, 0.0, 1.0, 0.0}, gates[4] = {0.0, 1.0, 1.0, 1.0};
    HostTrack H;
    if (build_host_track(walls, 1, gates, 1, 0.5, 0.5, angle, H) != 0) return -1;
    for (int k = 0; k < kHeadings; ++k) { cos_sin_out[2 * k] = H.trig64[k].x; cos_sin_out[2 * k + 1] = H.trig64[k].y; }
    return 0;
}

// The per-warp work list of the time-sliced table kernel, exactly as the kernel computes it.
extern "C" void emul_slice_items(int n_jobs, int n_steps, int slot, int n_slots, int *out4) {
    const SliceItems r = slice_items(n_jobs, n_steps, slot, n_slots);
    out4[0] = r.first_job; out4[1] = r.head_len; out4[2] = r.last_full; out4[3] = r.tail_len;
}
