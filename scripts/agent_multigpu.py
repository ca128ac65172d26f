"""Env-sharded fine-tuning run under torchrun (one rank per GPU): two iterations of the agent, then consistency checks.
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/agent_multigpu.py [workload]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dppo_b200.workloads import get_workload, make_agent_cfg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
name = sys.argv[1] if len(sys.argv) > 1 else "furniture"
w = get_workload(name)
from dppo_b200.agent.finetune.train_ppo_diffusion_agent import TrainPPODiffusionAgent

cfg = make_agent_cfg(w, f"cuda:{local}", tempfile.mkdtemp(), n_envs=50 if name != "hopper" else 40, n_steps=6, batch_size=100,
                     update_epochs=2, n_train_itr=3)
ag = TrainPPODiffusionAgent(cfg)
ag.n_critic_warmup_itr = 1
res = ag.run()
# every rank must hold identical parameters and running reward statistics after the all-reduced updates
flat = torch.cat([p.detach().reshape(-1) for p in list(ag.model.actor_ft.parameters()) + list(ag.model.critic.parameters())])
ref = flat.clone()
dist.broadcast(ref, 0)
same = bool(torch.equal(flat, ref))
stats = ag.running_reward_scaler.stats.clone()
ref_s = stats.clone()
dist.broadcast(ref_s, 0)
ok = torch.tensor([int(same and torch.equal(stats, ref_s))], device=f"cuda:{local}")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    last = res[-1]
    print(f"{name}: world {world}, envs/rank {ag.n_envs}, itrs {len(res)}, pg_loss {last['pg_loss']:.4e} v_loss {last['v_loss']:.4e} "
          f"kl {last['approx_kl']:.3e} minibatches {last['minibatches']}  params+stats identical on all ranks: {bool(ok.item())}"
          f"  param checksum {float(flat.double().sum()):.12e} / {float(flat.double().pow(2).sum()):.12e}")
    assert ok.item() == 1
from dppo_b200 import distributed as D

D.shutdown()  # registered gradient buffers first, then the process group
