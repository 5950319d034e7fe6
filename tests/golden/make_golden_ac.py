"""Golden fixtures for the continuous off-policy learners: runs the UNMODIFIED reference DDPG / TD3 / SAC learn() on an
injected minibatch (memory.sample patched) with every normal draw injected (torch.randn_like for TD3's target
smoothing, Normal.rsample's _standard_normal for SAC) and records the result dict and the post-step parameters of
every network (targets included).  Run in the build container: `python tests/golden/make_golden_ac.py`."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_inputs as G  # noqa: E402
from refimport import import_reference  # noqa: E402


def _load(net, case, name):
    net.load_state_dict({k: torch.from_numpy(v) for k, v in G.ac_params(case, name).items()})


def gen(agent_mod, name, case):
    inp = G.ac_case_inputs(case)
    ag = case["agent"]
    optim = {"actor": "adam", "critic": "adam", "alpha": "adam", "actor_lr": case["actor_lr"], "critic_lr": case["critic_lr"],
             "alpha_lr": case["alpha_lr"]}
    kw = dict(state_size=case["D"], action_size=case["A"], hidden_size=case["H"], optim_config=optim, gamma=case["gamma"],
              buffer_size=64, batch_size=case["B"], device="cpu", run_step=1000, lr_decay=False, tau=case["tau"])
    if ag == "sac":
        kw["use_dynamic_alpha"] = case["dynamic_alpha"]
    agent = agent_mod.Agent(ag, **kw)
    _load(agent.actor, case, "actor")
    if ag == "ddpg":
        _load(agent.critic, case, "critic1"); _load(agent.target_critic, case, "target_critic1")
        _load(agent.target_actor, case, "target_actor")
    else:
        _load(agent.critic1, case, "critic1"); _load(agent.critic2, case, "critic2")
        _load(agent.target_critic1, case, "target_critic1"); _load(agent.target_critic2, case, "target_critic2")
        if ag == "td3":
            _load(agent.target_actor, case, "target_actor")
            agent.num_learn = case["num_learn"]
    transitions = {k: inp[k] for k in ("state", "action", "reward", "next_state", "done")}
    agent.memory.sample = lambda bs: {k: v.copy() for k, v in transitions.items()}
    import torch.distributions.normal as tdn
    real_randn_like, real_std_normal = torch.randn_like, tdn._standard_normal
    out = {}
    try:
        for i, nz in enumerate(inp["noise"]):
            torch.randn_like = lambda t, nz=nz: torch.from_numpy(nz["target"]).clone()
            queue = [torch.from_numpy(nz["next"]), torch.from_numpy(nz["actor"])]
            tdn._standard_normal = lambda shape, dtype, device, queue=queue: queue.pop(0).clone()
            result = agent.learn()
            for k, v in result.items():
                out[f"result{i}.{k}"] = np.float64(v)
    finally:
        torch.randn_like, tdn._standard_normal = real_randn_like, real_std_normal
    nets = {"actor": agent.actor}
    if ag == "ddpg":
        nets.update(critic1=agent.critic, target_critic1=agent.target_critic, target_actor=agent.target_actor)
    else:
        nets.update(critic1=agent.critic1, critic2=agent.critic2, target_critic1=agent.target_critic1,
                    target_critic2=agent.target_critic2)
        if ag == "td3":
            nets["target_actor"] = agent.target_actor
    for n, net in nets.items():
        for k, v in net.state_dict().items():
            out[f"param.{n}.{k}"] = G.subsample(v.numpy())
    if ag == "sac":
        out["log_alpha"] = np.float64(agent.log_alpha.detach().item())
        out["alpha"] = np.float64(agent.alpha.detach().item())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: round(float(v), 6) for k, v in out.items() if k.startswith("result") or k in ("log_alpha", "alpha")})


def main():
    agent_mod, _, _ = import_reference()
    for name, case in G.AC_CASES.items():
        gen(agent_mod, name, case)


if __name__ == "__main__":
    main()
