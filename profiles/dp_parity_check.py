"""N-rank data-parallel LSA gradients == one-rank gradients on the concatenated batch, bit for bit (run under torchrun)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
import nerfq_b200
from nerfq_b200 import codec, model as nmodel, lsa, distributed as D, render as R
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from bench import synth_batch
n = 1024
def make():
    torch.manual_seed(0)
    w = nmodel.LSA(nmodel.NeRFWrapper()).add_lsa_params().to(dev)
    codec.quantize_model(w, -20)
    return w
batches = [synth_batch(n, 2 + 10 * r) for r in range(world)]
# data parallel
D.enable_data_parallel(True)
w = make()
step = lsa.LSAStep(w, n, lr=1e-3, perturb=0.0, white_bkgd=True)
o, d, t = batches[rank]
loss = step(torch.stack([o, d]).to(dev), t.to(dev))
g_dp = step.grad.clone(); p_dp = step.flat.clone()
lsum = loss.clone(); dist.all_reduce(lsum)
# single rank, whole batch
D.enable_data_parallel(False)
w1 = make()
step1 = lsa.LSAStep(w1, n * world, lr=1e-3, perturb=0.0, white_bkgd=True)
o = torch.cat([b[0] for b in batches]); d = torch.cat([b[1] for b in batches]); t = torch.cat([b[2] for b in batches])
loss1 = step1(torch.stack([o, d]).to(dev), t.to(dev))
torch.cuda.synchronize()
eq_g = torch.equal(g_dp, step1.grad); eq_p = torch.equal(p_dp, step1.flat)
gl = [torch.empty_like(g_dp) for _ in range(world)]; dist.all_gather(gl, g_dp)
same = all(torch.equal(gl[0], x) for x in gl)
print(f"rank {rank}: grads bitwise equal to single-rank {eq_g}, scales after Adam {eq_p}, equal across ranks {same}, max|g| {float(g_dp.abs().max()):.3e}, "
      f"max diff {float((g_dp - step1.grad).abs().max()):.3e}, loss dp-sum {float(lsum):.7f} single {float(loss1):.7f}", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0 if (eq_g and eq_p and same) else 1)
