"""Multi-GPU learner plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch) for the only real exchange step of the PPO path — the gradient all-reduce between
backward and the clip+Adam launch.  Collection, GAE and advantage standardisation
are per-env-row and never cross ranks.  The reference has no counterpart (single learner; ray
actors only collect: manager/distributed_manager.py:7-65).

Known, documented deviation at world_size > 1: `critic_loss = max(mean1, mean2)` (ppo.py:151-154)
is evaluated per rank on its local minibatch shard (a 2-float all-reduce before backward would
make it global; at B=256/rank the two means are equal in all but the clipped-value regime).
"""
import torch
import torch.distributed as dist


def attach(agent, world_size, average_with="avg"):
    """Makes `agent` a data-parallel learner: identical initial weights on every rank (broadcast from
    rank 0) and an averaged flat gradient before every optimiser step."""
    if world_size <= 1:
        return agent
    dist.broadcast(agent.network.flat, src=0)
    if hasattr(agent, "target_network"):
        dist.broadcast(agent.target_network.flat, src=0)
    agent.world_size = world_size

    def allreduce(flat_grad):
        if average_with == "avg":                       # NCCL: averaging fused into the collective
            dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)
        else:                                           # gloo (CPU tests) has no AVG
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
            flat_grad.div_(world_size)

    agent.allreduce = allreduce
    mem = getattr(agent, "memory", None)
    if mem is not None and hasattr(mem, "sample_device") and hasattr(mem, "tree_size"):
        mem.shard_world = world_size            # PER tree becomes one shard of a world_size-way replay
    return agent
