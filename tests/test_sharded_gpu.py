"""Two-rank run of the time-sharded CUDA path against the unsharded CUDA path and the oracle.

With >= 2 GPUs every rank owns a GPU and the collectives are NCCL; on a single-GPU box both ranks share cuda:0 and
the (tiny) collectives are staged through host memory over gloo -- the kernels and the shard / halo logic are the same."""
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


def _worker(rank, world, port, pcm, stationary, n_mels, padding, results):
    from modular_audio_pipeline_b200 import FrontEnd
    from modular_audio_pipeline_b200.sharded import Comm, CudaShardBackend, ShardGeometry, run_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    multi = torch.cuda.device_count() >= world
    dev = torch.device("cuda", rank if multi else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if multi else "gloo", rank=rank, world_size=world, **({"device_id": dev} if multi else {}))
    try:
        fe = FrontEnd(dev)
        geo = ShardGeometry(len(pcm), rank, world)
        backend = CudaShardBackend(fe, geo, torch.from_numpy(pcm[geo.local_slice()].copy()), stationary=stationary, n_mels=n_mels,
                                   mel_padding=padding)
        out, mel, lufs = run_sharded(backend, geo, stationary=stationary, comm=Comm(None, via_cpu=not multi))
        torch.cuda.synchronize()
        results[rank] = (out.cpu().numpy(), mel.cpu().numpy(), float(lufs.cpu().item()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("stationary,n_mels,padding,n_samples", [(True, 80, 0, 1500000), (False, 128, 480000, 1500000),
                                                                  (False, 80, 480000, 12345678)])
def test_two_ranks_equal_one_gpu(frontend, stationary, n_mels, padding, n_samples):
    # 3 chunks (ranks own 2 + 1), and a 12.9 min recording: 21 chunks, ranks own 11 + 10, ragged last chunk
    pcm = speech_noise(n_samples / SR, 8, meeting=not stationary).numpy()
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), pcm, stationary, n_mels, padding, results), nprocs=world, join=True)
    out = np.concatenate([results[r][0] for r in range(world)])
    mel = np.concatenate([results[r][1] for r in range(world)], axis=1)
    # single-GPU, unsharded run of the same recording: must agree bit for bit (same kernels, same chunk grids)
    plan = frontend.plan([len(pcm)], stationary=stationary, n_mels=n_mels, mel_padding=padding)
    res = plan.process(plan.pack([torch.from_numpy(pcm)]))
    assert np.array_equal(out, res.pcm[0].cpu().numpy())
    assert results[0][2] == results[1][2] and abs(results[0][2] - res.lufs.item()) < 1e-9
    assert mel.shape == tuple(res.log_mel[0].shape) and np.abs(mel - res.log_mel[0].cpu().numpy()).max() <= 1e-6
    # and the oracle, on the GPU's own samples
    want = O.log_mel_spectrogram(out.astype(np.float32) / 32768.0, n_mels, padding)
    assert np.abs(mel - want).max() <= 1e-4
