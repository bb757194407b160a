import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, json
from modular_audio_pipeline_b200 import FrontEnd
from modular_audio_pipeline_b200.transcriber import GpuLogMelFrontEnd
fe = FrontEnd("cuda:0"); lm = GpuLogMelFrontEnd(fe.device); lm._fe = fe
mel = torch.randn(128, 363000, device=fe.device)
seeks = list(range(0, 363000, 3000))
res = {}
for dt in (torch.float16, torch.float32):
    for _ in range(3): out = lm.windows(mel, seeks, None, dtype=dt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=fe.device)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = lm.windows(mel, seeks, None, dtype=dt); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    by = mel.numel() * 4 + out.numel() * out.element_size()
    res[str(dt)] = {"ms": round(ms, 4), "GBps": round(by / ms / 1e6, 1), "windows": len(seeks)}
print(json.dumps(res))
