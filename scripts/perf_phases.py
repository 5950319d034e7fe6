import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from jorldy_b200.core import Agent, Env
    from jorldy_b200.core.collect import RolloutCollector
    N, T, B = 4096, int(os.environ.get("T", 32)), 256
    env = Env("cartpole", num_envs=N, seed=0)
    agent = Agent("ppo", state_size=4, action_size=2, hidden_size=512, batch_size=B, n_step=T, n_epoch=1,
                  optim_config={"name": "adam", "lr": 2.5e-4}, device="cuda", run_step=10**9, use_fused=True)
    col = RolloutCollector(env, agent, use_cuda_graph=False); col.collect()
    agent.learn_rollout(col.rollout); col.rollout.t = T
    st = agent._st; fr = agent._fused[B]
    n = N * T // B
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agent._cursor.zero_(); e0.record(); fr.run(st, n); e1.record(); torch.cuda.synchronize()
    print(f"skip={os.environ.get('JB_FUSED_SKIP','0'):>3}: {e0.elapsed_time(e1)/n*1000:.1f} us/step over {n} steps", flush=True)
else:
    for skip in [0, 31, 30, 29, 27, 23, 15, 1, 2, 4, 8, 16]:
        env = dict(os.environ, JB_FUSED_SKIP=str(skip))
        subprocess.run([sys.executable, __file__, "child"], env=env)
