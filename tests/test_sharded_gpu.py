"""Multi-rank runs of the time-sharded CUDA path against the unsharded CUDA path and the oracle.

With as many GPUs as ranks every rank owns a GPU and the collectives are NCCL over NVLink (a whole step is also replayed as
ONE CUDA graph, collectives included); on a box with fewer GPUs the ranks share cuda:0 and the (tiny) collectives are staged
through host memory over gloo -- the kernels and the shard / halo / pack / unpack logic are the same.  ``MAF_EXPECT_GPUS=N``
(set when the suite is run under ``gpurun --gpus N``) turns a missing GPU into a FAILURE instead of the gloo route."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modular_audio_pipeline_b200.synth import speech_noise
from oracle import reference_path as O

pytestmark = pytest.mark.gpu
SR = 16000


EXPECT = int(os.environ.get("MAF_EXPECT_GPUS", "0"))


def _worker(rank, world, port, pcm, stationary, n_mels, padding, results):
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("MAF_RANK_TIMEOUT", "240")), exit=True)   # a stuck collective must not hold the box
    from modular_audio_pipeline_b200 import FrontEnd
    from modular_audio_pipeline_b200.sharded import Comm, CudaShardBackend, ShardGeometry, run_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    multi = torch.cuda.device_count() >= world
    assert multi or EXPECT < world, f"{world} ranks need {world} GPUs (MAF_EXPECT_GPUS={EXPECT}), found {torch.cuda.device_count()}"
    dev = torch.device("cuda", rank if multi else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if multi else "gloo", rank=rank, world_size=world, **({"device_id": dev} if multi else {}))
    try:
        fe = FrontEnd(dev)
        geo = ShardGeometry(len(pcm), rank, world)
        redzone = world == 2 and n_mels == 128               # one case doubles as the bounds check of the sharded stages
        if redzone:
            os.environ["MAF_DEBUG_REDZONE"] = "4096"
        backend = CudaShardBackend(fe, geo, torch.from_numpy(pcm[geo.local_slice()].copy()), stationary=stationary, n_mels=n_mels,
                                   mel_padding=padding)
        os.environ.pop("MAF_DEBUG_REDZONE", None)
        if redzone:                                          # workspace = 0xFF everywhere; the gaps must survive the step
            backend.plan._ws = torch.full((backend.plan.workspace_bytes,), 255, dtype=torch.uint8, device=dev)
            backend._mel_max = backend.plan.buffer("mel_max", torch.int32)
        comm = Comm(None, via_cpu=not multi)
        out, mel, lufs = run_sharded(backend, geo, stationary=stationary, comm=comm)
        torch.cuda.synchronize()
        if redzone:
            ws, lay = backend.plan._ws, backend.plan.workspace_layout()
            starts = sorted(o for o, _ in lay) + [backend.plan.workspace_bytes]
            for e in sorted(o + sz for o, sz in lay):
                nxt = min(x for x in starts if x >= e)
                assert nxt == e or bool((ws[e:nxt] == 255).all()), f"rank {rank}: workspace red zone [{e}, {nxt}) written"
        res = [out.cpu().numpy().copy(), mel.cpu().numpy().copy(), float(lufs.cpu().item()), "nccl" if multi else "gloo", comm.calls, None]
        if multi:                                          # the same step as one CUDA graph (kernels + NCCL collectives)
            replay = backend.capture(Comm(None))
            o2, m2, l2 = replay()
            torch.cuda.synchronize()
            res[5] = bool(np.array_equal(o2.cpu().numpy(), res[0]) and np.array_equal(m2.cpu().numpy(), res[1]) and float(l2.cpu().item()) == res[2])
            backend.release_graph()                        # a live graph with NCCL nodes would block destroy_process_group
        results[rank] = tuple(res)
    finally:
        faulthandler.cancel_dump_traceback_later()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world,stationary,n_mels,padding,n_samples", [
    (2, True, 80, 0, 1500000), (2, False, 128, 480000, 1500000), (2, False, 80, 480000, 12345678),
    (4, True, 80, 480000, 12345678), (4, False, 128, 0, 2403001)])
def test_ranks_equal_one_gpu(frontend, world, stationary, n_mels, padding, n_samples):
    # 3 chunks (ranks own 2 + 1), a 12.9 min recording (21 chunks, ragged last chunk), and 5 chunks over 4 ranks with a 3001-sample last shard
    pcm = speech_noise(n_samples / SR, 8, meeting=not stationary).numpy()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), pcm, stationary, n_mels, padding, results), nprocs=world, join=True)
    out = np.concatenate([results[r][0] for r in range(world)])
    mel = np.concatenate([results[r][1] for r in range(world)], axis=1)
    backend_kind = results[0][3]
    assert backend_kind == ("nccl" if torch.cuda.device_count() >= world else "gloo")
    assert results[0][4] == (5 if stationary else 3)       # collectives per step: one per barrier of the algorithm
    if backend_kind == "nccl":
        assert all(results[r][5] for r in range(world)), "CUDA-graph replay of the sharded step differs from the eager step"
    # single-GPU, unsharded run of the same recording: must agree bit for bit (same kernels, same chunk grids)
    plan = frontend.plan([len(pcm)], stationary=stationary, n_mels=n_mels, mel_padding=padding)
    res = plan.process(plan.pack([torch.from_numpy(pcm)]))
    assert np.array_equal(out, res.pcm[0].cpu().numpy())
    assert all(results[r][2] == results[0][2] for r in range(world)) and abs(results[0][2] - res.lufs.item()) < 1e-9
    assert mel.shape == tuple(res.log_mel[0].shape) and np.abs(mel - res.log_mel[0].cpu().numpy()).max() <= 1e-6
    # and the oracle, on the GPU's own samples
    want = O.log_mel_spectrogram(out.astype(np.float32) / 32768.0, n_mels, padding)
    assert np.abs(mel - want).max() <= 1e-4
