"""world_size-2 gloo test (CPU) of the data-parallel recipe the VAE step uses on NCCL: each rank computes the
gradient of its contiguous shard with loss_scale = 1/world into ONE flat buffer [grads | recon, kl], a single
all_reduce(sum) follows, and every rank applies the same Adam update.  The compute here is the oracle (this is
a test of the host-side sharding/reduction logic, not of the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import vae_oracle as vo
    torch.set_num_threads(2)
    w = vo.glorot_init(0)
    names = list(w.keys())
    x = np.random.RandomState(0).rand(4, 80, 160, 3); eps = np.random.RandomState(1).randn(4, 64)
    shard = 4 // world
    sl = slice(rank * shard, (rank + 1) * shard)
    out = vo.loss_and_grads(w, x[sl], x[sl], eps[sl], "mse")
    scale = 1.0 / world                                   # cpb_vae_config.loss_scale
    flat = np.concatenate([out["grads"][n].ravel() * scale for n in names] + [np.array([out["recon"] * scale, out["kl"] * scale])])
    buf = torch.from_numpy(flat)
    dist.all_reduce(buf)                                  # the ONE collective of the step
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), buf.numpy())
    dist.destroy_process_group()


def test_sharded_gradient_allreduce_equals_full_batch(tmp_path):
    from oracle import vae_oracle as vo
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npy"); r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1)                         # replicas stay identical
    w = vo.glorot_init(0)
    x = np.random.RandomState(0).rand(4, 80, 160, 3); eps = np.random.RandomState(1).randn(4, 64)
    full = vo.loss_and_grads(w, x, x, eps, "mse")
    ref = np.concatenate([full["grads"][n].ravel() for n in w] + [np.array([full["recon"], full["kl"]])])
    assert np.linalg.norm(r0 - ref) / np.linalg.norm(ref) < 1e-12
