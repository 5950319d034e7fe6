import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jorldy_b200.core import Agent, Env
from jorldy_b200.core.collect import RolloutCollector
def run(N, T, B, cont=False, fused=True, epochs=3, D=4, A=2):
    if cont:
        # synthetic obs-11 / act-3 rollout (Hopper dims): random tensors in a DeviceRollout
        from jorldy_b200.core.buffer import DeviceRollout
        agent = Agent("ppo", state_size=D, action_size=A, hidden_size=512, network="continuous_policy_value", batch_size=B,
                      n_step=T, n_epoch=epochs, optim_config={"name": "adam", "lr": 3e-4}, device="cuda", run_step=10**9, use_fused=fused)
        ro = DeviceRollout(N, T, D, A, "continuous", device="cuda")
        ro.state.normal_(); ro.action.uniform_(-0.9, 0.9); ro.reward.normal_(); ro.done.bernoulli_(0.001); ro.last_next_state.normal_()
        fn = lambda: (setattr(ro, "t", T), agent.learn_rollout(ro))
    else:
        env = Env("cartpole", num_envs=N, seed=0)
        agent = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=B, n_step=T, n_epoch=epochs,
                      optim_config={"name": "adam", "lr": 2.5e-4}, device="cuda", run_step=10**9, use_fused=fused)
        col = RolloutCollector(env, agent); col.collect()
        fn = lambda: agent.learn_rollout(col.rollout)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    steps = epochs * (N * T // B)
    print(f"N={N} T={T} B={B} cont={cont} fused={fused}: learn {ms:.1f} ms, {ms/steps*1000:.1f} us/minibatch-step, {N*T*epochs/ms/1e3:.2f} M transitions/s", flush=True)
run(4096, 128, 256, fused=True)
run(4096, 128, 256, fused=False)
run(4096, 128, 2048, fused=True)
run(1024, 256, 2048, cont=True, fused=True, epochs=2, D=11, A=3)
run(1024, 256, 2048, cont=True, fused=False, epochs=2, D=11, A=3)
