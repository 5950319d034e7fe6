"""Golden fixtures for the value-based learners: runs the UNMODIFIED reference agent's learn() on an
injected minibatch (memory.sample patched), with NoisyNet draws injected by patching torch.randn, and
records loss / stats / priorities written to the tree / post-step parameters."""
import os

import numpy as np
import torch

import gen_inputs as G

HERE = os.path.dirname(os.path.abspath(__file__))


def gen_q(agent_mod, name, case):
    inp = G.q_case_inputs(case)
    params = G.q_params(case)
    tparams = G.q_params(case, seed_offset=1000)
    optim = case.get("optim", {"name": "adam", "lr": case["lr"]})
    kw = dict(state_size=case["D"], action_size=case["A"], hidden_size=case["H"], optim_config=dict(optim),
              gamma=case["gamma"], buffer_size=case["buffer_size"], batch_size=case["B"], device="cpu", run_step=1000,
              lr_decay=False)
    ag = case["agent"]
    if case.get("head"):
        kw["head"] = case["head"]
    if ag in ("multistep", "rainbow", "ape_x"):
        kw["n_step"] = case["n_step"]
    if ag in ("per", "rainbow", "ape_x"):
        kw["alpha"] = case["alpha"]
    if ag in ("c51", "rainbow"):
        kw.update(v_min=case["v_min"], v_max=case["v_max"], num_support=case["K"])
    if ag == "ape_x":
        kw.update(network="dueling", clip_grad_norm=case["clip"], num_workers=2)
    if ag == "dueling":
        kw.update(network="dueling")
    agent = agent_mod.Agent(ag, **kw)
    agent.network.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    agent.target_network.load_state_dict({k: torch.from_numpy(v) for k, v in tparams.items()})
    transitions = {k: inp[k] for k in ("state", "action", "reward", "next_state", "done")}
    written = []
    if ag in ("per", "rainbow", "ape_x"):
        agent.memory.sample = lambda beta, bs: ({k: v.copy() for k, v in transitions.items()}, inp["weights"].copy(),
                                                inp["indices"].copy(), 0.5, 0.25)
        agent.memory.update_priority = lambda p, i: written.append((int(i), float(p)))
    else:
        agent.memory.sample = lambda bs: {k: v.copy() for k, v in transitions.items()}
    real_randn = torch.randn
    if inp["noise"] is not None:
        passes = inp["noise"] if ag == "rainbow" else [inp["noise"][0], inp["noise"][2]]
        queue = [torch.from_numpy(e) for layers in passes for pair in layers for e in pair]

        def fake_randn(*size, **kwargs):
            t = queue.pop(0)
            assert t.numel() == int(np.prod(size)), (t.shape, size)
            return t.clone()

        torch.randn = fake_randn
    try:
        result = agent.learn()
    finally:
        torch.randn = real_randn
    out = {f"result.{k}": np.float64(v) for k, v in result.items() if isinstance(v, (int, float, np.floating))}
    for k, v in agent.network.state_dict().items():
        out[f"param.{k}"] = G.subsample(v.numpy())
    if written:
        out["prio.idx"] = np.array([w[0] for w in written], dtype=np.int64)
        out["prio.p"] = np.array([w[1] for w in written], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: round(float(v), 6) for k, v in result.items() if isinstance(v, (int, float, np.floating))})


def main(agent_mod, buffer_mod, network_mod):
    for name, case in G.Q_CASES.items():
        gen_q(agent_mod, name, case)
