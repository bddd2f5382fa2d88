// carenv_tables.h — host-side preparation of the per-track constants (runs once per handle).
//
// Input is the track exactly as the reference builds it in CarEnv.reset
// (lib/car_env.py:651-676): wall segments in pixel coordinates, outer polyline first then
// inner, gates from consecutive point pairs, the start pose in pixels / degrees.
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

#include "carenv_core.cuh"

namespace carenv {

struct HostTrack {
    TrackParams P;
    std::vector<F2> trig32, trig32s;
    std::vector<float> den4;     // [72][n_pairs][4]: seg_den of both segments of a pair (tracks the pair kernels take)
    int n_pairs = 0;
    std::vector<D2> trig64, acc64;
    std::vector<GateRec> gates;
    std::vector<double> walls64;
    std::vector<SegF> segf;      // every segment (TrackParams only holds them up to kMaxSeg)
    std::vector<SegD> segd;
    Tables tables() const {
        return Tables{trig32.data(), trig32s.data(), trig64.data(), acc64.data(), gates.data(), walls64.data(),
                      segf.data(), segd.data()};
    }
};

// returns 0 on success, <0 on invalid input
inline int build_host_track(const double *walls, int n_walls, const double *gates, int n_gates,
                            double sx, double sy, double angle_deg, HostTrack &H) {
    if (n_walls < 1 || n_walls > kMaxBigSeg || n_gates < 1 || !walls || !gates) return -1;
    memset(&H.P, 0, sizeof(H.P));
    TrackParams &P = H.P;
    P.n_seg = n_walls; P.n_gates = n_gates;
    P.start_x = sx; P.start_y = sy;

    const double DEG = M_PI / 180.0;   // np.radians
    H.trig32.resize(kHeadings); H.trig32s.resize(kHeadings); H.trig64.resize(kHeadings); H.acc64.resize(kHeadings);
    for (int k = 0; k < kHeadings; ++k) {
        const double ang = angle_deg + 5.0 * k;
        const double c = cos(ang * DEG), s = sin(ang * DEG);
        H.trig64[k] = D2{c, s};
        H.trig32[k] = F2{(float)c, (float)s};
        H.trig32s[k] = F2{(float)c * kQScale, (float)s * kQScale};      // exact: a power of two
        H.acc64[k] = D2{c * 0.8, s * 0.8};
    }

    H.walls64.assign(walls, walls + 4 * (size_t)n_walls);
    H.segf.resize(n_walls);
    H.segd.resize(n_walls);
    double min_sin = 1.0;
    for (int j = 0; j < n_walls; ++j) {
        const double ax = walls[4 * j], ay = walls[4 * j + 1], bx = walls[4 * j + 2], by = walls[4 * j + 3];
        SegF &f = H.segf[j];
        memset(&f, 0, sizeof(f));
        f.ahx = (float)ax; f.ahy = (float)ay; f.nahy = -f.ahy;
        f.bhx = (float)bx; f.bhy = (float)by; f.nbhy = -f.bhy;
        const double ex = bx - ax, ey = by - ay;
        f.ex = (float)ex; f.ey = (float)ey; f.ney = -f.ey;
        f.chain_start = (j == 0 || walls[4 * j - 2] != ax || walls[4 * j - 1] != ay) ? 1 : 0;
        H.segd[j] = SegD{fma(ex, ay, -(ey * ax)) * kQScaleD, ex * kQScaleD, ey * kQScaleD};   // exact scalings
        if (n_walls <= kMaxSeg) { P.segf[j] = H.segf[j]; P.segd[j] = H.segd[j]; }
        const double len = hypot(ex, ey);
        if (len > 0)
            for (int k = 0; k < kHeadings; ++k) {
                const double sn = fabs(ex * H.trig64[k].y - ey * H.trig64[k].x) / len;
                if (sn > 1e-7 && sn < min_sin) min_sin = sn;   // exactly parallel pairs cannot be hit
            }
    }

    H.gates.resize(n_gates);
    for (int g = 0; g < n_gates; ++g) {
        GateRec &r = H.gates[g];
        r.x1 = gates[4 * g]; r.y1 = gates[4 * g + 1]; r.x2 = gates[4 * g + 2]; r.y2 = gates[4 * g + 3];
        const double ex = r.x2 - r.x1, ey = r.y2 - r.y1;
        r.ex = (float)ex; r.ey = (float)ey; r.len = (float)hypot(ex, ey); r.pad = 0.0f;
    }

    // guard bands: bounds on the float32 error of q, r and the gate margin
    P.eps_q = 1.5e-3f;      // |dq| <= 5.7e-4 px for coordinates within 1280 x 720 (see wall_point)
    P.eps_qs = P.eps_q * kQScale;
    (void)min_sin;
    const double rel_r = 5.0e-7;   // relative error of r = cross(e,d) / cross(e,A'): both from float64, rounded once, one
                                   // approximate reciprocal, one multiply
    P.coll_band = (float)fmax(2.0e-4, 4.0 * rel_r);
    P.tiny_d = 1.0e-2f;
    P.gate_band = 2.0e-3f;
    P.tiny_un = 1.0e-3f * kQScale;   // float64 error of cross(e, A - pos) is ~1e-10: relative error < 1e-7 above 1e-3
    // loop unrolling the track allows (see cast_walls)
    P.unroll = (n_walls <= kMaxSeg) ? 1 : 0;
    P.unroll4 = 1;
    const int candidates[3] = {6, 4, 2};                     // measured on big_track: U = 6 2.31 ms, 4 2.36, 12 2.39, 2 2.53
    for (int ci = 0; ci < 3 && n_walls <= kMaxSeg; ++ci) {
        const int U = candidates[ci];
        bool ok = (n_walls % U == 0);
        for (int j = 0; j < n_walls && ok; ++j)
            if (H.segf[j].chain_start && j % U != 0) ok = false;
        if (ok && P.unroll <= 1) P.unroll = U;
        if (ok && U <= 4 && P.unroll4 <= 1) P.unroll4 = U;
    }

    // denominators of the pair kernels, tabulated per heading (k_rollout_tab): exactly what wall_pair computes
    H.n_pairs = 0;
    H.den4.clear();
    if (P.unroll >= 2) {
        H.n_pairs = n_walls / 2;
        H.den4.resize((size_t)kHeadings * H.n_pairs * 4);
        for (int k = 0; k < kHeadings; ++k)
            for (int jp = 0; jp < H.n_pairs; ++jp) {
                float *d = &H.den4[((size_t)k * H.n_pairs + jp) * 4];
                const D2 t = H.trig64[k];
                seg_den(H.segd[2 * jp].ex, H.segd[2 * jp].ey, t.y, t.x, d[0], d[1]);
                seg_den(H.segd[2 * jp + 1].ex, H.segd[2 * jp + 1].ey, t.y, t.x, d[2], d[3]);
            }
    }

    const Tables T = H.tables();
    P.start_destroyed = reset_observation(P, T, P.reset_obs) ? 1 : 0;
    return 0;
}

}  // namespace carenv
