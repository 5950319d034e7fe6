"""Mints tests/golden/*.npz from the UNMODIFIED reference classes (run in the build container:
`python tests/golden/make_golden.py`).  Inputs are regenerated from gen_inputs.py by the tests, so
the fixtures hold only what the reference produced.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_inputs as G  # noqa: E402
from refimport import import_reference  # noqa: E402


def snapshot_locals(code_obj, names, store):
    """sys.setprofile hook capturing selected locals when `code_obj` returns."""
    def prof(frame, event, arg):
        if event == "return" and frame.f_code is code_obj:
            for n in names:
                if n in frame.f_locals:
                    v = frame.f_locals[n]
                    store[n] = v.detach().clone().numpy() if torch.is_tensor(v) else np.array(v)
    return prof


def gen_ppo(agent_mod, name, case):
    torch.manual_seed(0)
    inp = G.ppo_case_inputs(case)
    shapes = G.ppo_shapes(case)
    params = G.make_params(shapes, case["seed"])
    net = "continuous_policy_value" if case["continuous"] else "discrete_policy_value"
    agent = agent_mod.Agent(
        "ppo", state_size=case["D"], action_size=case["A"], hidden_size=case["H"], network=net,
        optim_config={"name": "adam", "lr": case["lr"]}, gamma=case["gamma"], use_standardization=case["standardize"],
        run_step=1000, lr_decay=False, device="cpu", batch_size=case["batch_size"], n_step=case["T"],
        n_epoch=case["n_epoch"], _lambda=case["lam"], epsilon_clip=case["eps_clip"], vf_coef=case["vf_coef"],
        ent_coef=case["ent_coef"], clip_grad_norm=case["clip_grad_norm"])
    agent.network.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    NT = case["N"] * case["T"]
    transitions = []
    for i in range(NT):
        transitions.append({"state": inp["state"][i:i + 1], "action": inp["action"][i:i + 1],
                            "reward": inp["reward"][i:i + 1], "next_state": inp["next_state"][i:i + 1],
                            "done": inp["done"][i:i + 1]})
    agent.memory.first_store = False
    agent.memory.store(transitions)
    perms = iter(inp["perms"])

    def fake_shuffle(arr):
        arr[:] = next(perms)

    real_shuffle = np.random.shuffle
    np.random.shuffle = fake_shuffle
    store = {}
    sys.setprofile(snapshot_locals(type(agent).learn.__code__, ["value", "next_value", "adv", "ret", "log_prob_old"], store))
    try:
        result = agent.learn()
    finally:
        sys.setprofile(None)
        np.random.shuffle = real_shuffle
    out = {f"result.{k}": np.float64(v) for k, v in result.items()}
    for k, v in store.items():
        out[f"pre.{k}"] = G.subsample(v.astype(np.float32))
    for k, v in agent.network.state_dict().items():
        out[f"param.{k}"] = G.subsample(v.numpy())
        out[f"pnorm.{k}"] = np.float64(np.linalg.norm(v.numpy().astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: round(float(v), 6) for k, v in result.items()})


def main():
    agent_mod, buffer_mod, network_mod = import_reference()
    for name, case in G.PPO_CASES.items():
        gen_ppo(agent_mod, name, case)
    try:
        import make_golden_more
        make_golden_more.main(agent_mod, buffer_mod, network_mod)
    except ImportError:
        pass


if __name__ == "__main__":
    main()
