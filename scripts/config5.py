"""frames -> batched ConvVAE.encode -> latents (+3 measurements) -> PPO.learn().

    python scripts/config5.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/config5.py

100 000 uint8 frames resident on the GPU(s) (synthetic: the reference's PNGs are not on the GPU box), encoded in
batches of 4096 (frames sharded over the ranks, latents all-gathered), then 48 rollouts of 2048 steps on rank 0:
states = latent (64) + steer, throttle, speed (vae_common.py:45-61), PPO.learn() = GAE + normalisation + theta_old
copy + 4 epochs x 8 minibatches of 256 (train.py:171-207).  Prints one JSON line with the end-to-end frames/s and
the encode / learn split (CUDA events; max over ranks for the encode)."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # one JSON line on stdout: whatever native libraries print meanwhile (NCCL's banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from carla_ppo_b200.vae.models import ConvVAE
    from carla_ppo_b200.ppo import PPO

    n_frames, batch, horizon = 100_000, 4096, 2048
    per_rank = -(-n_frames // world)
    tmp = tempfile.mkdtemp()
    golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    vae = ConvVAE((80, 160, 3), z_dim=64, loss_fn="mse", model_dir=os.path.join(tmp, "vae%d" % rank), seed=0, training=False)
    vae.init_session(init_logging=False)
    zv = np.load(os.path.join(golden, "vae_rgb_ckpt232.npz"))
    vae.set_weights({k: zv[k] for k in vae._names})                 # the reference's shipped rgb VAE (checkpoint-232)
    g = torch.Generator(device="cuda"); g.manual_seed(rank)
    frames = torch.randint(0, 256, (per_rank, 80, 160, 3), dtype=torch.uint8, device="cuda", generator=g)   # 3.84 GB / world
    latents = torch.empty(per_rank, 64, device="cuda")

    def encode_all():
        for i in range(0, per_rank, batch):
            latents[i:i + batch] = vae.encode_device(frames[i:i + batch], check=False)

    encode_all(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); encode_all()
    if world > 1:
        gathered = torch.empty(world * per_rank, 64, device="cuda")
        dist.all_gather_into_tensor(gathered, latents)
    else:
        gathered = latents
    e1.record(); torch.cuda.synchronize()
    enc_ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(enc_ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        class Box:
            low = np.array([-1.0, 0.0], np.float32); high = np.array([1.0, 1.0], np.float32); shape = (2,)
        ppo = PPO((67,), Box(), learning_rate=1e-4, value_scale=1.0, model_dir=os.path.join(tmp, "ppo"), seed=0)
        ppo.init_session(init_logging=False)
        zp = np.load(os.path.join(golden, "ppo_ckpt705.npz"))         # the reference's shipped agent (checkpoint-705)
        ppo.set_weights({k: zp["policy/" + k] for k in ppo._names}, {k: zp["policy_old/" + k] for k in ppo._names},
                        {k: zp["adam_m/" + k] for k in ppo._names}, {k: zp["adam_v/" + k] for k in ppo._names},
                        (float(zp["beta1_power"]), float(zp["beta2_power"])))
        gen = torch.Generator(device="cuda"); gen.manual_seed(1)
        n_roll = (gathered.shape[0] // horizon)
        n_roll = min(n_roll, 48)
        meas = torch.rand(n_roll * horizon, 3, device="cuda", generator=gen)
        states = torch.cat([gathered[:n_roll * horizon], meas], dim=1).reshape(n_roll, horizon, 67)
        actions = torch.randn(n_roll, horizon, 2, device="cuda", generator=gen).clamp_(min=torch.tensor([-1.0, 0.0], device="cuda"),
                                                                                    max=torch.tensor([1.0, 1.0], device="cuda"))
        rewards = torch.rand(n_roll, horizon, device="cuda", generator=gen, dtype=torch.float64)
        values = torch.randn(n_roll, horizon, device="cuda", generator=gen, dtype=torch.float64)
        dones = torch.zeros(horizon, device="cuda", dtype=torch.float64); dones[-1] = 1
        perms = torch.stack([torch.randperm(horizon, device="cuda", generator=gen) for _ in range(4)]).to(torch.int32)
        ppo.learn(states[0], actions[0], values[0], rewards[0], dones, 0.3, num_epochs=4, batch_size=256, perms=perms)
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for r in range(n_roll):
            ppo.learn(states[r], actions[r], values[r], rewards[r], dones, 0.3, num_epochs=4, batch_size=256, perms=perms)
        l1.record(); torch.cuda.synchronize()
        learn_ms = l0.elapsed_time(l1)
        total_ms = float(enc_ms.item()) + learn_ms
        n_used = n_roll * horizon
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps({"config": "SURVEY 8(d) config 5", "n_gpus": world, "frames_encoded": world * per_rank,
                          "encode_ms": float(enc_ms.item()), "encode_frames_per_s": world * per_rank / float(enc_ms.item()) * 1e3,
                          "rollouts": n_roll, "learn_ms_total": learn_ms, "ms_per_learn": learn_ms / n_roll,
                          "end_to_end_frames_per_s": n_used / total_ms * 1e3,
                          "allgather_bytes": int(world * per_rank * 64 * 4) if world > 1 else 0,
                          "data": "synthetic uint8 frames resident in HBM; shipped VAE checkpoint-232 and agent checkpoint-705 weights"}), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
