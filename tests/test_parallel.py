"""CPU (gloo, world_size 2) tests of the data-parallel plumbing: clip sharding, hypothesis-major row selection, and
that a sharded step + all-reduce of the shared-parameter gradients reproduces the single-process step.  The step is
the hot path's math restated in torch (oracle/torch_ref.py) because the CUDA kernels cannot run here; the GPU
counterpart is bench.py under torchrun."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acfm_video_3d_reconstruction_b200 import parallel
from oracle import torch_ref
from tests import util


def test_shard_clips_partitions_exactly():
    for clips in (1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_clips(clips, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == clips
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    idx = parallel.shard_frames(5, 4, 1, 2)               # rank 1 of 2 owns clips 3,4 -> frames 12..19
    assert idx.tolist() == list(range(12, 20))
    rows = parallel.multiplex_rows(idx, 20, 3)
    assert rows.tolist() == [g * 20 + i for g in range(3) for i in range(12, 20)]


def _step(mean_v, lbs_raw, delta, cams, frame_rows, G, n_total):
    """Toy hot-path step in torch: skin -> multiplex projection -> smooth per-render loss -> hypothesis weighting.
    Normalised by the GLOBAL frame count so that shard sums equal the full-batch value."""
    lbs = torch.softmax(lbs_raw, dim=0)
    pred_v = mean_v[None] + torch.einsum("vk,bkc->bvc", lbs, delta)
    nb = delta.shape[0]
    p = torch_ref.orthographic_proj_withz(pred_v.repeat(G, 1, 1), cams, 0.0)
    per = (p[..., :2] ** 2).mean((1, 2)).view(G, nb) + 0.1 * p[..., 2].mean(1).view(G, nb)
    probs = torch.softmax(-per, dim=0).detach()
    return (per * probs).sum() / n_total


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    v, _ = util.template("bird")
    clips, T, G, Kh = 5, 2, 3, 8
    V = v.shape[0]
    mean_v = torch.from_numpy(v).double().requires_grad_(True)
    lbs_raw = torch.randn(V, Kh, dtype=torch.float64).requires_grad_(True)
    delta = 0.05 * torch.randn(clips * T, Kh, 3, dtype=torch.float64)
    cams = torch.from_numpy(util.synth_cams(G * clips * T, seed=3)).double()
    idx = parallel.shard_frames(clips, T, rank, world)
    rows = parallel.multiplex_rows(idx, clips * T, G)
    d_loc = delta[idx].clone().requires_grad_(True)
    c_loc = cams[rows].clone().requires_grad_(True)
    loss = _step(mean_v, lbs_raw, d_loc, c_loc, rows, G, clips * T)
    loss.backward()
    calls = parallel.allreduce_shared_grads([mean_v, lbs_raw], bucket_bytes=1 << 14)   # forces several buckets
    tot = torch.tensor([float(loss)], dtype=torch.float64)
    dist.all_reduce(tot)
    q.put((rank, calls, float(tot), mean_v.grad.numpy(), lbs_raw.grad.numpy(), idx.numpy(), rows.numpy(), d_loc.grad.numpy(),
           c_loc.grad.numpy(), parallel.max_over_ranks(float(rank), "cpu")))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_step_matches_single_process():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=180) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth on the full batch
    torch.manual_seed(0)
    v, _ = util.template("bird")
    clips, T, G, Kh = 5, 2, 3, 8
    V = v.shape[0]
    mean_v = torch.from_numpy(v).double().requires_grad_(True)
    lbs_raw = torch.randn(V, Kh, dtype=torch.float64).requires_grad_(True)
    delta = (0.05 * torch.randn(clips * T, Kh, 3, dtype=torch.float64)).requires_grad_(True)
    cams = torch.from_numpy(util.synth_cams(G * clips * T, seed=3)).double().requires_grad_(True)
    loss = _step(mean_v, lbs_raw, delta, cams, None, G, clips * T)
    loss.backward()
    assert outs[0][1] >= 2 and outs[0][9] == 1.0                        # several buckets; max over ranks
    for o in outs:
        assert abs(o[2] - float(loss)) < 1e-12
        assert np.allclose(o[3], mean_v.grad.numpy(), rtol=1e-10, atol=1e-14)    # shared grads: identical on all ranks
        assert np.allclose(o[4], lbs_raw.grad.numpy(), rtol=1e-10, atol=1e-14)
        assert np.allclose(o[7], delta.grad.numpy()[o[5]], rtol=1e-10, atol=1e-14)   # per-frame grads: no exchange needed
        assert np.allclose(o[8], cams.grad.numpy()[o[6]], rtol=1e-10, atol=1e-14)
    assert sorted(np.concatenate([o[5] for o in outs]).tolist()) == list(range(clips * T))
