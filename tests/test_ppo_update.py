"""Fused PPO minibatch update (csrc/ppo_update.cuh) against PyTorch autograd / clip_grad_norm_ / Adam on the same
minibatch — the update math of train.py:223-261 for the reference network."""
import math

import pytest
import torch
import torch.nn as nn

import ppo_car_b200
from ppo_car_b200.ppo_update import FusedPPOUpdate
from ppo_car_b200.train_ppo import ActorCritic, parse_args, train

pytestmark = pytest.mark.gpu


def _torch_loss(net, obs, act, old_logp, adv, ret, clip=0.2, vf=0.5, ent_c=0.001):
    _, new_logp, ent, new_val = net.act(obs, act)
    ratio = torch.exp(new_logp - old_logp)
    a = (adv - adv.mean()) / torch.clamp(adv.std(), min=1e-5)
    pol = torch.max(-a * ratio, -a * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
    vl = 0.5 * ((new_val.view(-1) - ret) ** 2).mean()
    e = ent.mean()
    return pol + vf * vl - ent_c * e, pol, vl, e


@pytest.mark.parametrize("batch,gathered", [(512, False), (200, False), (512, True), (1000, False)])
def test_gradients_match_autograd(batch, gathered):
    dev = torch.device("cuda")
    torch.manual_seed(batch)
    net = ActorCritic(18, 9).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.05)
        net.actor[2].weight.mul_(20.0)                       # a policy far from uniform: ratios leave the clip range
    M = 5000
    obs = torch.rand((M, 18), device=dev)
    act = torch.randint(0, 9, (M,), device=dev).float()
    old_logp = torch.log_softmax(net.actor(obs), -1).gather(-1, act.long().unsqueeze(-1)).squeeze(-1).detach()
    old_logp = old_logp + torch.randn(M, device=dev) * 0.3   # stale log-probs: both clip branches and the tie region
    adv, ret = torch.randn(M, device=dev) * 2 + 0.5, torch.randn(M, device=dev)
    idx = torch.randint(0, M, (batch,), device=dev)
    upd = FusedPPOUpdate(net.actor, net.critic, batch, lr=3e-4)
    if gathered:
        upd.grad(obs[idx].contiguous(), idx, act, old_logp, adv, ret, obs_is_gathered=True)
    else:
        upd.grad(obs, idx, act, old_logp, adv, ret)
    loss, pol, vl, e = _torch_loss(net, obs[idx], act[idx], old_logp[idx], adv[idx], ret[idx])
    net.zero_grad()
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in upd.params])
    got = upd.grads
    scale = ref.abs().max()
    assert torch.allclose(got, ref, rtol=2e-4, atol=float(scale) * 2e-6), float((got - ref).abs().max() / scale)
    inside = ((torch.exp(net.act(obs[idx], act[idx])[1] - old_logp[idx]) - 1).abs() <= 0.2).float().mean()
    assert 0.1 < float(inside) < 0.9                         # the test exercises clipped and unclipped samples
    # Adam step + statistics against torch (fresh optimiser: first step)
    params_before = [p.detach().clone() for p in upd.params]
    nn.utils.clip_grad_norm_(upd.params, 1.0)
    opt = torch.optim.Adam(upd.params, lr=3e-4, eps=1e-5)
    opt.step()
    want = [p.detach().clone() for p in upd.params]
    with torch.no_grad():
        for p, b in zip(upd.params, params_before):
            p.copy_(b)
    upd.apply()
    for p, w, b in zip(upd.params, want, params_before):
        # step = lr * g / (|g| + eps): for gradients near eps = 1e-5 a 1e-3 relative difference of g moves the step by
        # up to lr * 1e-3 / 4, hence the absolute term (the step itself is ~3e-4)
        assert torch.allclose(p - b, w - b, rtol=1e-3, atol=3e-7), float(((p - b) - (w - b)).abs().max())
    s = upd.sums.tolist()
    assert math.isclose(s[0], float(pol.detach()), rel_tol=1e-4, abs_tol=1e-6) and math.isclose(s[1], float(vl.detach()), rel_tol=1e-4)
    assert math.isclose(s[2], float(e.detach()), rel_tol=1e-4) and math.isclose(s[3], float(loss.detach()), rel_tol=1e-4, abs_tol=1e-6)
    assert int(upd.step_count) == 1


def test_many_adam_steps_track_torch_adam():
    """40 consecutive updates on fresh minibatches: parameters stay within 1e-4 of torch's clip + Adam."""
    dev = torch.device("cuda")
    torch.manual_seed(1)
    net_a, net_b = ActorCritic(18, 9).to(dev), ActorCritic(18, 9).to(dev)
    net_b.load_state_dict(net_a.state_dict())
    M, B = 4096, 512
    obs = torch.rand((M, 18), device=dev)
    act = torch.randint(0, 9, (M,), device=dev).float()
    old_logp = torch.full((M,), -math.log(9.0), device=dev)
    adv, ret = torch.randn(M, device=dev), torch.randn(M, device=dev) * 0.3
    upd = FusedPPOUpdate(net_a.actor, net_a.critic, B, lr=3e-4)
    opt = torch.optim.Adam(net_b.parameters(), lr=3e-4, eps=1e-5)
    for _ in range(40):
        idx = torch.randint(0, M, (B,), device=dev)
        upd.grad(obs, idx, act, old_logp, adv, ret)
        upd.apply()
        loss = _torch_loss(net_b, obs[idx], act[idx], old_logp[idx], adv[idx], ret[idx])[0]
        opt.zero_grad()
        loss.backward()
        nn.utils.clip_grad_norm_(list(net_b.parameters()), 1.0)
        opt.step()
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        assert torch.allclose(pa, pb, rtol=0, atol=1e-4), float((pa - pb).abs().max())
    assert int(upd.step_count) == 40


@pytest.mark.parametrize("extra", [["--fused-rollout"], ["--fused-rollout", "--per-minibatch-update", "--graph-update"],
                                   ["--fused-rollout", "--compact-obs", "--graph-update"], ["--cuda-graph"]])
def test_training_with_fused_update_improves_reward(extra):
    args = parse_args(["--track", "big_track", "--n-envs", "64", "--n-epochs", "12", "--n-steps", "256",
                       "--fused-update"] + extra)
    hist = train(args)
    assert all(math.isfinite(h["total_loss"]) for h in hist)
    assert hist[-1]["avg_reward"] > hist[0]["avg_reward"] + 0.03
    assert abs(hist[-1]["lr"] - 3e-4 * 0.99 ** 12) < 1e-9


def _epoch_problem(dev, M=6000, seed=3):
    torch.manual_seed(seed)
    net = ActorCritic(18, 9).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    obs = torch.rand((M, 18), device=dev)
    act = torch.randint(0, 9, (M,), device=dev).float()
    old_logp = torch.log_softmax(net.actor(obs), -1).gather(-1, act.long().unsqueeze(-1)).squeeze(-1).detach()
    old_logp = old_logp + torch.randn(M, device=dev) * 0.2
    adv, ret = torch.randn(M, device=dev) * 2 + 0.5, torch.randn(M, device=dev)
    return net, obs, act, old_logp, adv, ret


@pytest.mark.parametrize("batch,n_updates,n_ctas", [(512, 12, 0), (200, 5, 0), (1000, 4, 0), (512, 6, 128), (37, 3, 49)])
def test_epoch_kernel_matches_the_per_minibatch_kernels(batch, n_updates, n_ctas):
    """carenv_ppo_epoch (one persistent cooperative launch for all minibatches) against the three-launch update on
    the same minibatches: same gradients up to summation order, so the parameters, Adam moments, step counter and
    loss statistics agree closely after several dependent updates."""
    dev = torch.device("cuda")
    net_a, obs, act, old_logp, adv, ret = _epoch_problem(dev)
    net_b = ActorCritic(18, 9).to(dev)
    net_b.load_state_dict(net_a.state_dict())
    idx = torch.randint(0, obs.shape[0], (n_updates, batch), device=dev)
    ua = FusedPPOUpdate(net_a.actor, net_a.critic, batch, lr=3e-4)
    ub = FusedPPOUpdate(net_b.actor, net_b.critic, batch, lr=3e-4)
    for u in range(n_updates):
        ua.grad(obs, idx[u].contiguous(), act, old_logp, adv, ret)
        ua.apply()
    ub.run_epoch(obs, idx, act, old_logp, adv, ret, n_ctas=n_ctas)
    ub.check_epoch()
    assert int(ub.step_count) == n_updates
    for pa, pb in zip(ua.params, ub.params):
        assert torch.allclose(pa, pb, rtol=0, atol=2e-6), float((pa - pb).abs().max())
    assert torch.allclose(ua.exp_avg, ub.exp_avg, rtol=1e-3, atol=1e-7)
    assert torch.allclose(ua.exp_avg_sq, ub.exp_avg_sq, rtol=2e-3, atol=1e-10)
    assert torch.allclose(ua.sums, ub.sums, rtol=1e-4, atol=1e-5), (ua.sums.tolist(), ub.sums.tolist())
    # deterministic: a second run from the same start gives the same bits
    net_c = ActorCritic(18, 9).to(dev)
    net_c.load_state_dict(_epoch_problem(dev)[0].state_dict())
    uc = FusedPPOUpdate(net_c.actor, net_c.critic, batch, lr=3e-4)
    uc.run_epoch(obs, idx, act, old_logp, adv, ret, n_ctas=n_ctas)
    for pb, pc in zip(ub.params, uc.params):
        assert torch.equal(pb, pc)


def test_epoch_kernel_argument_checks():
    dev = torch.device("cuda")
    net, obs, act, old_logp, adv, ret = _epoch_problem(dev, M=1000)
    upd = FusedPPOUpdate(net.actor, net.critic, 64, lr=3e-4)
    idx = torch.randint(0, 1000, (3, 64), device=dev)
    with pytest.raises(ValueError):
        upd.run_epoch(obs, idx[:, :32].contiguous(), act, old_logp, adv, ret)
    with pytest.raises(ppo_car_b200.CarEnvError):
        upd.run_epoch(obs, idx, act, old_logp, adv, ret, world=2)          # no connect()
    with pytest.raises(ppo_car_b200.CarEnvError):
        upd.run_epoch(obs, idx, act, old_logp, adv, ret, n_ctas=500)


def _peer_worker(rank, world, port, q):
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        B, n_updates = 512, 10
        net_a, *_ = _epoch_problem(dev, seed=3)                      # same initial weights on every rank
        net_b = ActorCritic(18, 9).to(dev)
        net_b.load_state_dict(net_a.state_dict())
        _, obs, act, old_logp, adv, ret = _epoch_problem(dev, seed=100 + rank)   # a different shard per rank
        torch.manual_seed(7 + rank)
        idx = torch.randint(0, obs.shape[0], (n_updates, B), device=dev)
        ua = FusedPPOUpdate(net_a.actor, net_a.critic, B, lr=3e-4)
        ub = FusedPPOUpdate(net_b.actor, net_b.critic, B, lr=3e-4)
        for u in range(n_updates):                                   # reference: NCCL all-reduce between two launches
            ua.grad(obs, idx[u].contiguous(), act, old_logp, adv, ret)
            ua.apply(world)
        ub.connect()
        ub.run_epoch(obs, idx[:4].contiguous(), act, old_logp, adv, ret, world=world)     # two launches: the sequence
        ub.run_epoch(obs, idx[4:].contiguous(), act, old_logp, adv, ret, world=world)     # numbers carry over
        ub.check_epoch()
        err = max(float((pa - pb).abs().max()) for pa, pb in zip(ua.params, ub.params))
        flat = torch.cat([p.detach().reshape(-1) for p in ub.params])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        ub.close()
        q.put((rank, err, same, int(ub.step_count)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one node")
def test_epoch_kernel_all_reduces_over_peer_memory():
    """Two ranks: the in-kernel NVLink gradient exchange against the NCCL all-reduce path, and bit-identical
    parameters on both ranks."""
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, 29731, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
    for rank, err, same, steps in res:
        assert err < 2e-6 and same and steps == 10, (rank, err, same, steps)
