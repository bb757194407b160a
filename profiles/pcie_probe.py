"""What the host<->device links of the box allow for bench.py's `e2e` figure: the HostPipeline copy pattern (115.2 MB of pinned
PCM up, 115.2 MB of PCM down per step, three streams) WITHOUT any kernel, next to the one-directional copy rates.

    python profiles/pcie_probe.py                                                      # one GPU, prints one JSON line
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29544 profiles/pcie_probe.py                                     # every GPU at once (what bench.py --gpus 8 does)

With N ranks every rank drives its own GPU with the same pattern at the same time (barrier before, max over ranks after):
`aggregate_GBps_per_direction` against N x the single-GPU figure says whether the HOST side (PCIe root complexes / memory /
the VM) or the per-GPU link is the limit of the N-GPU end-to-end number.
"""
import json
import os
import time

import torch

N = 57_600_000                      # int16 samples of one hour at 16 kHz
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(300, exit=True)
    dist.init_process_group("nccl", device_id=dev)
depth = 3
h_in = [torch.empty(N, dtype=torch.int16).pin_memory() for _ in range(depth)]
h_out = [torch.empty(N, dtype=torch.int16).pin_memory() for _ in range(depth)]
d_in = [torch.empty(N, dtype=torch.int16, device=dev) for _ in range(depth)]
d_out = [torch.zeros(N, dtype=torch.int16, device=dev) for _ in range(depth)]
streams = [torch.cuda.Stream(dev) for _ in range(depth)]


def run(steps, up, down):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        s = i % depth
        with torch.cuda.stream(streams[s]):
            if up:
                d_in[s].copy_(h_in[s], non_blocking=True)
            if down:
                h_out[s].copy_(d_out[s], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    return dt


for _ in range(2):
    run(6, True, True)
res = {"n_gpus": world}
for name, up, down in (("h2d_only", True, False), ("d2h_only", False, True), ("both_directions", True, True)):
    dt = run(30, up, down)
    res[name] = {"ms_per_step": round(dt * 1e3, 3), "GBps_per_direction_per_gpu": round(2 * N / dt / 1e9, 1),
                 "aggregate_GBps_per_direction": round(world * 2 * N / dt / 1e9, 1)}
res["e2e_ceiling_audio_s_per_s"] = round(world * 3600.0 / (res["both_directions"]["ms_per_step"] * 1e-3))
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
