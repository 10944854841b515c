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


def _worker_host_logic(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from carla_ppo_b200.vae.models import dp_noise_rows, dp_shared_permutation
    gen = torch.Generator(); gen.manual_seed(123)                        # the same seed on every rank
    rows = dp_noise_rows(torch, gen, 3, 64, rank, world, "cpu")
    np.random.seed(1000 + rank)                                          # ranks seed np.random independently ...
    idx = np.arange(50); np.random.shuffle(idx)
    shared = dp_shared_permutation(torch, dist, idx, "cpu")              # ... and adopt rank 0's permutation
    # replicas start from rank 0's state (VAE._broadcast_state) and the verify_range flag travels with the gradients
    params = torch.full((8,), float(rank + 1)); dist.broadcast(params, 0)
    buf = torch.zeros(8 + 3); buf[8 + 2] = 1.0 if rank == 1 else 0.0      # [grads | recon, kl, flag]: only rank 1 saw a bad value
    dist.all_reduce(buf)
    np.savez(os.path.join(out_dir, "host%d.npz" % rank), rows=rows.numpy(), perm=shared, params=params.numpy(), flag=buf[10].numpy())
    dist.destroy_process_group()


def test_data_parallel_host_logic_noise_permutation_broadcast_flag(tmp_path):
    """The host-side rules of the data-parallel VAE step: (1) the ranks' noise rows concatenate to the
    single-process draw of the global batch, (2) one shared epoch permutation, (3) replicas adopt rank 0's state,
    (4) a verify_range flag raised on ONE rank reaches every rank through the step's single all-reduce."""
    port = _free_port()
    mp.spawn(_worker_host_logic, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "host0.npz"); b = np.load(tmp_path / "host1.npz")
    gen = torch.Generator(); gen.manual_seed(123)
    full = torch.randn(6, 64, generator=gen).numpy()
    assert np.array_equal(np.concatenate([a["rows"], b["rows"]]), full)
    assert np.array_equal(a["perm"], b["perm"]) and sorted(a["perm"]) == list(range(50))
    np.random.seed(1000); ref = np.arange(50); np.random.shuffle(ref)
    assert np.array_equal(a["perm"], ref)
    assert np.array_equal(a["params"], np.ones(8)) and np.array_equal(b["params"], np.ones(8))
    assert a["flag"] > 0 and b["flag"] > 0
