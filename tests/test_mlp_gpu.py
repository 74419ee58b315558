"""tcgen05 fused MLP forward (fp16 operands with fp32 accumulation in the default four-slot kernel, TF32 in the two-slot one) vs the
plain fp32 torch forward of the same nn.Sequential.  Both keep a 10-bit mantissa: tolerance 5e-3 absolute on O(1) outputs (SURVEY.md
§8d cfg 5 allows a tensor-core contraction for G4; the PPO update runs the same fp16 operand precision, so rollout and update agree)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _net(i, hs, o):
    layers, d = [], i
    for h in hs:
        layers += [nn.Linear(d, h), nn.ELU()]
        d = h
    return nn.Sequential(*layers, nn.Linear(d, o)).cuda()


@pytest.mark.parametrize("num_obs,hidden,out,batch", [(48, (128, 64, 32), 12, 4096), (48, (128, 64, 32), 1, 1000), (48, (128, 64, 32), 12, 7),
                                                      (48, (128, 64, 32), 12, 24576), (30, (64, 48), 5, 300)])
def test_fused_mlp_matches_torch(num_obs, hidden, out, batch):
    from legged_gym_dev_b200.mlp import FusedMLP
    torch.manual_seed(0)
    net = _net(num_obs, hidden, out)
    x = torch.randn(batch, num_obs, device="cuda")
    with torch.no_grad():
        want = net(x)
    got = FusedMLP(net)(x)
    torch.cuda.synchronize()
    err = (got - want).abs().max().item()
    assert got.shape == want.shape
    assert err < 5e-3, f"max abs error {err}"
    assert (got - want).abs().mean().item() < 1e-3
    # repack after a weight change
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(0.5)
    f = FusedMLP(net)
    with torch.no_grad():
        want2 = net(x)
    assert (f(x) - want2).abs().max().item() < 5e-3


@pytest.mark.parametrize("batch_a,batch_b", [(4096, 4096), (7, 1000), (24576, 4096), (131072, 131072)])
def test_pair_launch_equals_two_single_launches(batch_a, batch_b):
    """b200gym_mlp_forward_pair (actor + critic of PPO.act in one launch) is bit-identical to two b200gym_mlp_forward calls:
    each CTA runs the single-net kernel body on its own range of the grid."""
    from legged_gym_dev_b200.mlp import FusedMLP
    torch.manual_seed(3)
    actor, critic = FusedMLP(_net(48, (128, 64, 32), 12)), FusedMLP(_net(48, (128, 64, 32), 1))
    assert actor.pairable and critic.pairable
    xa, xb = torch.randn(batch_a, 48, device="cuda"), torch.randn(batch_b, 48, device="cuda")
    oa, ob = FusedMLP.forward_pair(actor, xa, critic, xb)
    torch.cuda.synchronize()
    assert torch.equal(oa, actor(xa)) and torch.equal(ob, critic(xb))
    # the same tensor as input of both nets (critic obs = obs on flat terrain)
    oa2, ob2 = FusedMLP.forward_pair(actor, xa, critic, xa)
    assert torch.equal(oa2, oa) and torch.equal(ob2, critic(xa))


def test_pair_launch_rejects_nets_outside_the_fp16_kernel():
    from legged_gym_dev_b200.mlp import FusedMLP
    odd = FusedMLP(_net(30, (64, 48), 5))
    ok = FusedMLP(_net(48, (128, 64, 32), 12))
    assert not odd.pairable
    with pytest.raises(RuntimeError):
        FusedMLP.forward_pair(ok, torch.randn(64, 48, device="cuda"), odd, torch.randn(64, 30, device="cuda"))


def test_fused_mlp_rejects_large_nets():
    from legged_gym_dev_b200.mlp import FusedMLP
    with pytest.raises(ValueError):
        FusedMLP(_net(235, (512, 256, 128), 12))


def test_actor_critic_uses_fused_forward_without_grad():
    from legged_gym_dev_b200.ppo import ActorCritic, PPO
    torch.manual_seed(1)
    ac = ActorCritic(48, 48, 12, actor_hidden_dims=[128, 64, 32], critic_hidden_dims=[128, 64, 32])
    alg = PPO(ac, device="cuda")
    assert ac._fused_actor is not None
    x = torch.randn(512, 48, device="cuda")
    with torch.no_grad():
        fused = ac.act_inference(x)
        v_fused = ac.evaluate(x)
    ref, v_ref = ac.actor(x), ac.critic(x)            # grad mode: autograd path
    assert ref.requires_grad and not fused.requires_grad
    assert (fused - ref).abs().max().item() < 5e-3 and (v_fused - v_ref).abs().max().item() < 5e-3
    with torch.no_grad():
        ac.flat_param.mul_(0.9)
    ac.repack_fused()
    with torch.no_grad():
        assert (ac.act_inference(x) - ac.actor(x)).abs().max().item() < 5e-3


@pytest.mark.parametrize("batch", [4096, 100])
def test_rough_nets_forward_on_grouped_gemm(batch):
    """235 -> 512 -> 256 -> 128 -> 12 / 1 (legged_robot_config.py:242-243) does not fit the weights-resident kernel: the forward runs
    layer by layer on the grouped tcgen05 GEMM (fp16 operands, fp32 accumulation), never on torch."""
    from legged_gym_dev_b200.ppo import ActorCritic, PPO
    torch.manual_seed(2)
    ac = ActorCritic(235, 235, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128])
    PPO(ac, device="cuda")
    assert ac._fused_actor is None and ac._trainer is not None
    x = torch.randn(batch, 235, device="cuda").clamp(-5, 5)
    with torch.no_grad():
        mu, v = ac.act_inference(x), ac.evaluate(x)
        mu_ref, v_ref = ac.actor(x), ac.critic(x)
    assert mu.shape == mu_ref.shape and v.shape == v_ref.shape
    assert (mu - mu_ref).abs().max().item() < 1e-2 and (mu - mu_ref).abs().mean().item() < 2e-3
    assert (v - v_ref).abs().max().item() < 1e-2
    both = ac._trainer.forward_both(x, x)
    assert torch.equal(both[0], mu) and torch.equal(both[1], v)


def test_unsupported_activation_raises():
    from legged_gym_dev_b200.ppo import ActorCritic, PPO
    ac = ActorCritic(48, 48, 12, actor_hidden_dims=[32], critic_hidden_dims=[32], activation="tanh")
    with pytest.raises(ValueError):
        PPO(ac, device="cuda")
