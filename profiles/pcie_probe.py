"""What the PCIe link of the box allows for bench.py's `e2e` figure: the HostPipeline copy pattern (115.2 MB of pinned PCM up,
115.2 MB of PCM down per step, three streams) WITHOUT any kernel, next to the one-directional copy rates.

    python profiles/pcie_probe.py        # prints one JSON line
"""
import json
import time

import torch

N = 57_600_000                      # int16 samples of one hour at 16 kHz
dev = torch.device("cuda:0")
depth = 3
h_in = [torch.empty(N, dtype=torch.int16).pin_memory() for _ in range(depth)]
h_out = [torch.empty(N, dtype=torch.int16).pin_memory() for _ in range(depth)]
d_in = [torch.empty(N, dtype=torch.int16, device=dev) for _ in range(depth)]
d_out = [torch.zeros(N, dtype=torch.int16, device=dev) for _ in range(depth)]
streams = [torch.cuda.Stream(dev) for _ in range(depth)]


def run(steps, up, down):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        s = i % depth
        with torch.cuda.stream(streams[s]):
            if up:
                d_in[s].copy_(h_in[s], non_blocking=True)
            if down:
                h_out[s].copy_(d_out[s], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


for _ in range(2):
    run(6, True, True)
res = {}
for name, up, down in (("h2d_only", True, False), ("d2h_only", False, True), ("both_directions", True, True)):
    dt = run(30, up, down)
    res[name] = {"ms_per_step": round(dt * 1e3, 3), "GBps_per_direction": round(2 * N / dt / 1e9, 1)}
res["e2e_ceiling_audio_s_per_s"] = round(3600.0 / (res["both_directions"]["ms_per_step"] * 1e-3))
print(json.dumps(res))
