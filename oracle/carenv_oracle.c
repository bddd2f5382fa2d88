/* TEST INFRASTRUCTURE — plain-C float64 oracle for the batched CarEnv step.
 *
 * A CPU restatement of the reference algorithm (lib/car_env.py in
 * /root/reference), written from scratch for speed so that parity tests can
 * compare the CUDA path on thousands of environments x 1024 steps in seconds.
 * It is a CHECKER: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it.  The product library never links or calls it.
 *
 * Parity status: PINNED through tests/test_oracle_golden.py, which requires this
 * file to reproduce the golden trajectories recorded from the unmodified
 * reference (tests/golden/carenv_*.npz): all integer outputs bit-exact, float32
 * observations equal to <= 1 float32 ulp, float64 rewards equal.  The only
 * arithmetic not bit-identical to the reference is libm cos/sin versus numpy's
 * (<= 1 ulp of double) — everything else is the same IEEE double operations in
 * the same order (compile with -ffp-contract=off).
 *
 * Reference map (file:line):
 *   cast()            lib/car_env.py:155-184   Ray.cast
 *   ray_distance()    lib/car_env.py:186-213   Ray.get_distance
 *   touches()         lib/car_env.py:376-392   Car.check_collision
 *   gate scan         lib/car_env.py:394-408   Car.get_passed_gate
 *   thrust / turn     lib/car_env.py:416-442   Car.move_car
 *   integrate()       lib/car_env.py:444-469   Car.update
 *   observe()         lib/car_env.py:569-597   CarEnv._get_obs
 *   reset_env()       lib/car_env.py:605-691   CarEnv.reset
 *   step_env()        lib/car_env.py:693-760   CarEnv.step
 *   autoreset         gymnasium 0.29.1 AsyncVectorEnv worker (un-vendored), SURVEY 3.5
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_RAYS 12
#define OBS_DIM 18
#define MAX_GATES 4096

typedef struct {
    double px, py, vx, vy, ax, ay, rot;
    double rdx[N_RAYS], rdy[N_RAYS]; /* ray directions left by the last integrate() */
    double rox, roy;                 /* ray origin left by the last integrate()     */
    int32_t t, next_gate, passed, remaining, destroyed;
    uint8_t active[MAX_GATES];
} env_t;

typedef struct {
    const double *walls; int n_walls;   /* [n_walls][4] x1 y1 x2 y2, outer first then inner */
    const double *gates; int n_gates;   /* [n_gates][4] */
    double sx, sy, angle;               /* start pose: pixels, degrees */
    int scan_all_gates;                 /* 1 = reference's full ordered scan */
} track_t;

static const double DEG = M_PI / 180.0;

/* returns 1 and the hit point when the ray meets the segment */
static int cast(double ox, double oy, double dx, double dy, const double *s, double *hx, double *hy) {
    double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
    double x3 = ox, y3 = oy, x4 = ox + dx, y4 = oy + dy;
    double den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    if (den == 0) return 0;
    double t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / den;
    double u = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)) / den;
    if (0 < t && t < 1 && u > 0) {
        *hx = x1 + t * (x2 - x1);
        *hy = y1 + t * (y2 - y1);
        return 1;
    }
    return 0;
}

static double ray_distance(double ox, double oy, double dx, double dy, const double *segs, int n) {
    double best = 1000.0;
    for (int j = 0; j < n; ++j) {
        double hx, hy;
        if (cast(ox, oy, dx, dy, segs + 4 * j, &hx, &hy)) {
            double ex = ox - hx, ey = oy - hy;
            double d = sqrt(ex * ex + ey * ey);
            if (d < best) best = d;
        }
    }
    return best;
}

static int touches(const env_t *e, const double *segs, int n) {
    for (int k = 0; k < N_RAYS; k += N_RAYS / 4)
        if (ray_distance(e->rox, e->roy, e->rdx[k], e->rdy[k], segs, n) < 10.0) return 1;
    return 0;
}

static void integrate(env_t *e, const track_t *tr) {
    e->vx += e->ax; e->vy += e->ay;
    if (sqrt(e->ax * e->ax + e->ay * e->ay) == 0) { e->vx *= 1 - 0.2; e->vy *= 1 - 0.2; }
    e->vx = fmin(fmax(e->vx, -10.0), 10.0);
    e->vy = fmin(fmax(e->vy, -10.0), 10.0);
    e->px += e->vx; e->py += e->vy;
    e->ax = 0.0; e->ay = 0.0;
    e->rox = e->px; e->roy = e->py;
    for (int k = 0; k < N_RAYS; ++k) {
        double ang = e->rot + (double)(k * (360 / N_RAYS));
        e->rdx[k] = cos(ang * DEG);
        e->rdy[k] = sin(ang * DEG);
    }
    if (touches(e, tr->walls, tr->n_walls)) e->destroyed = 1;
}

static void observe(const env_t *e, const track_t *tr, float *obs, double *dist_or_null) {
    obs[0] = (float)(e->px / 1280);
    obs[1] = (float)(e->py / 720);
    obs[2] = (float)(e->vx / 10.0);
    obs[3] = (float)(e->vy / 10.0);
    obs[4] = (float)cos(e->rot * DEG);
    obs[5] = (float)sin(e->rot * DEG);
    for (int k = 0; k < N_RAYS; ++k) {
        double d = ray_distance(e->rox, e->roy, e->rdx[k], e->rdy[k], tr->walls, tr->n_walls);
        obs[6 + k] = (float)(d / 1000.0);
        if (dist_or_null) dist_or_null[k] = d;
    }
}

static void reset_env(env_t *e, const track_t *tr) {
    e->t = 0;
    e->px = tr->sx; e->py = tr->sy; e->rot = tr->angle;
    e->vx = e->vy = e->ax = e->ay = 0.0;
    e->passed = 0; e->next_gate = 0; e->remaining = tr->n_gates; e->destroyed = 0;
    memset(e->active, 1, (size_t)tr->n_gates);
    integrate(e, tr);
}

static double step_env(env_t *e, const track_t *tr, int action, int *terminated, int *truncated) {
    double reward = 0.0;
    int fwd = (action == 0 || action == 4 || action == 5);
    int bwd = (action == 1 || action == 6 || action == 7);
    int left = (action == 2 || action == 4 || action == 6);
    int right = (action == 3 || action == 5 || action == 7);
    if (fwd || bwd) {
        double c = cos(e->rot * DEG), s = sin(e->rot * DEG);
        if (fwd) { e->ax = c * 0.8; e->ay = s * 0.8; reward += 0.01; }
        else     { e->ax = -c * 0.8; e->ay = -s * 0.8; }
    }
    if (left) e->rot -= 5.0; else if (right) e->rot += 5.0;

    int hit = -1;
    if (tr->scan_all_gates) {
        for (int g = 0; g < tr->n_gates; ++g)
            if (e->active[g] && touches(e, tr->gates + 4 * g, 1)) { hit = g; break; }
    } else if (touches(e, tr->gates + 4 * e->next_gate, 1)) {
        hit = e->next_gate;   /* equivalent: gates below next_gate are exactly the inactive ones */
    }
    if (hit >= 0 && hit == e->next_gate) {
        reward += 1.0;
        e->remaining -= 1;
        e->passed += 1;
        if (e->remaining == 0) {
            reward += 10.0;
            memset(e->active, 1, (size_t)tr->n_gates);
            e->remaining = tr->n_gates;
            e->next_gate = 0;
        } else {
            e->active[hit] = 0;
            e->next_gate += 1;
        }
    }
    integrate(e, tr);
    e->t += 1;
    *terminated = 0; *truncated = 0;
    if (e->destroyed) { *terminated = 1; reward -= 3.0; }
    else if (e->t >= 1000) { *truncated = 1; }
    return reward;
}

/* ---- exported C API (ctypes) ------------------------------------------------ */

size_t oracle_env_bytes(void) { return sizeof(env_t); }

/* All entry points work on the env sub-range [env_lo, env_hi) of arrays laid out for
 * n_envs envs, so the Python wrapper can shard a call over host threads (ctypes drops
 * the GIL); there is no OpenMP runtime in the image.
 * Reset every env of the range; writes the reset observation [N,18]. */
int oracle_reset(void *state, int env_lo, int env_hi, const double *walls, int n_walls, const double *gates, int n_gates,
                 double sx, double sy, double angle, float *obs, double *dist_or_null) {
    if (n_gates > MAX_GATES) return -1;
    track_t tr = {walls, n_walls, gates, n_gates, sx, sy, angle, 1};
    env_t *E = (env_t *)state;
    for (int i = env_lo; i < env_hi; ++i) {
        reset_env(&E[i], &tr);
        observe(&E[i], &tr, obs + (size_t)i * OBS_DIM, dist_or_null ? dist_or_null + (size_t)i * N_RAYS : NULL);
    }
    return 0;
}

/* T steps of N envs with same-step autoreset.  actions [T,N] uint8.  Any output may be NULL.
 * obs      [T,N,18] observation returned to the caller (reset obs on done steps)
 * fobs     [T,N,18] observation of the finished step before the autoreset
 * rew      [T,N] float64 (unscaled), term/trunc [T,N] uint8
 * gates_passed/time_passed/next_gate [T,N] int32: info of the finished step, pre-reset
 * pose     [T,N,5] float64 px py vx vy rot of the finished step, pre-reset             */
int oracle_rollout(void *state, int n_envs, int env_lo, int env_hi, int T, const uint8_t *actions,
                   const double *walls, int n_walls, const double *gates, int n_gates,
                   double sx, double sy, double angle, int scan_all_gates,
                   float *obs, float *fobs, double *rew, uint8_t *term, uint8_t *trunc,
                   int32_t *gates_passed, int32_t *time_passed, int32_t *next_gate, double *pose) {
    if (n_gates > MAX_GATES) return -1;
    track_t tr = {walls, n_walls, gates, n_gates, sx, sy, angle, scan_all_gates};
    env_t *E = (env_t *)state;
    for (int i = env_lo; i < env_hi; ++i) {
        env_t *e = &E[i];
        float o[OBS_DIM];
        for (int t = 0; t < T; ++t) {
            size_t k = (size_t)t * n_envs + i;
            int te, tu;
            double r = step_env(e, &tr, actions[k], &te, &tu);
            observe(e, &tr, o, NULL);
            if (fobs) memcpy(fobs + k * OBS_DIM, o, sizeof o);
            if (rew) rew[k] = r;
            if (term) term[k] = (uint8_t)te;
            if (trunc) trunc[k] = (uint8_t)tu;
            if (gates_passed) gates_passed[k] = e->passed;
            if (time_passed) time_passed[k] = e->t;
            if (next_gate) next_gate[k] = e->next_gate;
            if (pose) { double *p = pose + k * 5; p[0] = e->px; p[1] = e->py; p[2] = e->vx; p[3] = e->vy; p[4] = e->rot; }
            if (te || tu) { reset_env(e, &tr); observe(e, &tr, o, NULL); }
            if (obs) memcpy(obs + k * OBS_DIM, o, sizeof o);
        }
    }
    return 0;
}
