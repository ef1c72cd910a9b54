"""N > 1 on real GPUs (needs >= 2 devices; the 1-GPU round-end run skips it — the builder runs it with `gpurun --gpus 2` and
commits the log under profiles/): the sharded sample() gathers, over NCCL, exactly what one rank computes for the whole batch."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from founddiff_b200 import distributed as fdist, weights
from founddiff_b200.diffusion import ResidualDiffusion, UnetRes
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sd = weights.random_state_dict(10)
model = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, input_condition=False, objective='pred_res', test_res_or_noise='res')
model.load_state_dict({"unet0." + k: v for k, v in sd.items()})
H, n = 128, 6
ldct = torch.rand(n, 1, H, H, generator=torch.Generator().manual_seed(5))
for dt, S in ((torch.float32, 2), (torch.bfloat16, 2), (torch.bfloat16, 1000)):
    model.compute_dtype = dt
    d = ResidualDiffusion(model, image_size=H, timesteps=1000, sampling_timesteps=S, objective='pred_res', loss_type='l2',
                          condition=True, sum_scale=0.01).to(dev)
    d.init()
    if S == 1000:
        d.num_timesteps = 6                      # a 6-step ancestral chain (same override as tests/golden/ancestral_32.npz)
    got = fdist.sample_sharded(d, ldct, noise_seed=99)
    assert got.shape == (n, 1, H, H)
    # the same call on ONE rank: rank 0 samples the whole batch with the same per-slice generators
    if rank == 0:
        sn = fdist.SliceNoise(99, range(n), (1, H, H))
        noise = {"init": sn.init()}
        if S == 1000:
            noise["steps"] = sn.steps(d.num_timesteps - 1)
        one = d.sample([ldct.to(dev)], batch_size=n, last=True, noise=noise)[-1]
        diff = float((one - got).abs().max())
        print(f"{dt} S={S}: 2-rank gather vs 1-rank, max abs diff {diff:.3e}", flush=True)
        assert torch.equal(one, got), (dt, S, diff)
    dist.barrier()
dist.destroy_process_group()
print("ok", rank, flush=True)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_gather_is_bit_identical_to_one_rank(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", port, str(script), ROOT], capture_output=True, text=True, timeout=1500)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-4000:]
