"""torchrun --nproc-per-node 2: FusedAdam.all_reduce_and_step (chunked all-reduce, Adam per arrived chunk) must equal
all-reduce of the whole buffer followed by step()."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import fused_adam
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
shapes = [(100003, 3), (100003, 15, 3), (100003, 1), (64, 6), (7,)]
lrs = [1.6e-4, 1.25e-4, 0.05, 1e-3, 1e-2]
g = torch.Generator().manual_seed(0)
base = [torch.randn(s, generator=g).cuda() for s in shapes]
mk = lambda: [b.clone().requires_grad_(True) for b in base]
pa, pb = mk(), mk()
oa = fused_adam.FusedAdam([{"params": [p], "lr": lr, "name": "g%d" % i} for i, (p, lr) in enumerate(zip(pa, lrs))], lr=0.0, eps=1e-15)
ob = fused_adam.FusedAdam([{"params": [p], "lr": lr, "name": "g%d" % i} for i, (p, lr) in enumerate(zip(pb, lrs))], lr=0.0, eps=1e-15)
gr = torch.Generator().manual_seed(100 + rank)
for step in range(3):
    oa.zero_grad(); ob.zero_grad()
    for p, q in zip(pa, pb):
        x = torch.randn(p.shape, generator=gr).cuda() * 0.1
        p.grad.copy_(x); q.grad.copy_(x)
    oa.all_reduce_and_step(chunks=5, skip=("g3",) if step == 1 else ())
    ob.grads.all_reduce(); ob.step(skip=("g3",) if step == 1 else ())
torch.cuda.synchronize()
ok = all(torch.equal(p.detach(), q.detach()) for p, q in zip(pa, pb)) and torch.equal(oa.exp_avg, ob.exp_avg) and oa._steps == ob._steps
chk = torch.tensor([float(pa[0].sum())], device="cuda")
lst = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(lst, chk)
same = all(float(x) == float(lst[0]) for x in lst)
print("rank %d PIPELINE_%s replicas_%s" % (rank, "OK" if ok else "MISMATCH", "identical" if same else "DIFFER"), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok and same else 1)
