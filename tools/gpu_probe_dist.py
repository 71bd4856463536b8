"""Phase-by-phase probe of the multi-rank captured training iteration (run under torchrun with a short `timeout`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
T0 = time.time()
def log(*a):
    torch.cuda.synchronize()
    print(f"[r{rank} +{time.time() - T0:6.2f}s]", *a, file=sys.stderr, flush=True)

dist.init_process_group("nccl", device_id=dev)
log("pg up")
from uaps_b200.train import UAPSConfig, UAPSTrainer
from uaps_b200.unet import UNet_UAPS
graph = os.environ.get("PROBE_GRAPH", "1") == "1"
B, H = int(os.environ.get("PROBE_B", "4")), int(os.environ.get("PROBE_H", "64"))
torch.manual_seed(100 + rank)
model = UNet_UAPS(3, 4).to(dev)
tr = UAPSTrainer(model, UAPSConfig(cuda_graph=graph, graph_warmup=2), group=dist.group.WORLD)
log("trainer up: device-state", tr.state is not None, "xchg", tr.xchg is not None)
g = torch.Generator().manual_seed(rank)
xl, xu = torch.randn(B, 3, H, H, generator=g).to(dev), torch.randn(B, 3, H, H, generator=g).to(dev)
yl = ((xl[:, 0] > 0).long() + 2 * (xl[:, 1] > 0).long())
for i in range(6):
    out = tr.step(xl, yl, xu)
    log(f"step {i}: loss {float(out['loss']):.5f} graphs {len(tr._graphs)} skipped {tr.skipped_steps()}")
chk = tr.optimizer.flat_p.double().sum()
all_ = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(all_, chk)
log("param checksums", [float(c) for c in all_])
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(10):
    tr.step(xl, yl, xu)
t1.record(); torch.cuda.synchronize()
log("ms/iter", t0.elapsed_time(t1) / 10)
from uaps_b200 import comm
dist.barrier()
comm.close_all()
dist.destroy_process_group()
log("done")
