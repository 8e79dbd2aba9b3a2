"""all_reduce(sum, fp64) of the E-step statistics' size (5 n + 2 doubles) over NCCL: microseconds per call on the device,
max over ranks.  Run under torchrun; NCCL_* environment variables select algorithm / protocol."""
import os, sys, json
import torch, torch.distributed as dist
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5 * 199999 + 2
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(n, dtype=torch.float64, device="cuda")
for _ in range(20):
    dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": dist.get_world_size(), "doubles": n, "us_per_allreduce": float(t.item()),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}), flush=True)
dist.destroy_process_group()
