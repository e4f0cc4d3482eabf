"""torchrun smoke of the data-parallel train() loop: three PPO iterations of CartPole over WORLD_SIZE GPUs, parameters must end
bit-identical on every rank (the per-update health check of the peer exchange runs too)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
from aur_ppo_b200.ppo import ppo  # noqa: E402
from bench import params  # noqa: E402

p = params("ppo", dist.get_world_size(), 3)
p.update(num_envs=2048 * dist.get_world_size(), total_timesteps=2048 * dist.get_world_size() * 128 * 3)
agent = ppo(p)
rets, lens, xs = agent.train()
flat = agent.flat.clone()
gathered = [torch.empty_like(flat) for _ in range(dist.get_world_size())]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
if dist.get_rank() == 0:
    print(f"train() over {dist.get_world_size()} ranks: {len(rets)} logged episodes, mean return {sum(rets) / max(len(rets), 1):.1f}, "
          f"parameters identical on every rank: {same}, stats {agent.last_stats}")
assert same
dist.destroy_process_group()
