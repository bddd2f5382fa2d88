"""Per-launch time vs steps per launch (fit a + b*K) at several env counts."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_car_b200
track = ppo_car_b200.builtin_track("big_track")
def t_roll(env, a, o, reps):
    kw = dict(obs_out=o["obs"], reward_out=o["reward"], term_out=o["terminated"], trunc_out=o["truncated"])
    for _ in range(3): env.rollout(a, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): env.rollout(a, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (75776, 131072, 151552, 1048576, 1060864):
    env = ppo_car_b200.VecCarEnv(n, track); env.reset()
    for K in (1, 2, 4, 8, 16, 32, 64):
        a = torch.randint(0, 9, (K, n), device="cuda", dtype=torch.uint8)
        o = env.rollout(a)
        ms = t_roll(env, a, o, 20)
        print(json.dumps({"n": n, "K": K, "us": round(ms * 1e3, 1), "us_per_step": round(ms * 1e3 / K, 2)}), flush=True)
