mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; head -c 300 gpurun_out/r2_bench_2gpu.json; echo
timeout 100 python -c "
import torch, tempfile
from carla_ppo_b200.vae.models import ConvVAE
v = ConvVAE((80,160,3), z_dim=64, loss_fn='mse', model_dir=tempfile.mkdtemp(), seed=0, device='cuda:1'); v.init_session(init_logging=False)
x = torch.rand(64,80,160,3, device='cuda:1'); e = torch.randn(64,64, device='cuda:1')
print('cuda:1 while cuda:0 current:', v.train_step_device(x,x,e).cpu().numpy(), 'workspace MB', sum(w.numel() for w in v._ws.values() if w is not None)/1e6)
" 2>&1 | tail -1
