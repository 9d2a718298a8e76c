"""torchrun --nproc-per-node N scripts/multigpu_entry_check.py : `clip_guided_diffusion(rank=, world_size=)` on N GPUs over NCCL -- every rank
owns one image, the final frames of the whole batch are gathered (one all_gather_into_tensor) and saved by rank 0.  Also checks the
sharded run against the same images computed on ONE GPU (rank 0 re-runs the whole batch): identical seeds -> identical draws."""
import os
import sys
import tempfile

import torch as th
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_guided_diffusion_b200 import cgd, unet as pu, vit as pv, weights as pw  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
th.cuda.set_device(local)
dist.init_process_group("nccl", device_id=th.device("cuda", local))
ucfg, vcfg = pu.config_for(64, True), pv.VIT_CONFIGS["ViT-B/32"]
usd = pw.seeded_state_dict(pw.unet_param_shapes(ucfg), 1234)
vsd = pw.seeded_state_dict(pw.vit_param_shapes(vcfg), 1235)
tgt = th.randn(1, 512, generator=th.Generator().manual_seed(0))
out = tempfile.mkdtemp() if rank == 0 else None
box = [out]
dist.broadcast_object_list(box, src=0)
out = box[0]
os.makedirs(os.path.join(out, f"cwd{rank}"), exist_ok=True)
os.chdir(os.path.join(out, f"cwd{rank}"))
kw = dict(image_size=64, num_cutouts=4, prompts=["multi gpu"], timestep_respacing="ddim25", skip_timesteps=0, save_frequency=100, progress=False, seed=3,
          device=f"cuda:{local}", unet_state_dict=usd, clip_state_dict=vsd, target_embeds=tgt)
got = list(cgd.clip_guided_diffusion(batch_size=world, prefix_path=os.path.join(out, "sharded"), rank=rank, world_size=world, **kw))
dist.barrier()
if rank == 0:
    from PIL import Image
    import numpy as np
    # frames: step 0 of the own image (save_frequency 100) + the gathered final step (current_timestep == -1) for EVERY image
    assert sorted(b for b, _ in got) == [0] + list(range(world)), got
    ref = list(cgd.clip_guided_diffusion(batch_size=world, prefix_path=os.path.join(out, "single"), **kw))
    finals = sorted(p for _, p in ref if p.endswith("0024.png"))
    assert len(finals) == world
    worst = 0
    for p in finals:
        a = np.asarray(Image.open(p)).astype(int)
        b = np.asarray(Image.open(p.replace("single", "sharded"))).astype(int)
        worst = max(worst, int(np.abs(a - b).max()))
    print(f"multi-GPU entry check: world {world}, rank 0 saved {world} gathered final frames; max |sharded - single| = {worst} of 255")
    # fp16 kernels pick different tilings for batch 1 and batch N and 25 free-running steps of seeded-random weights amplify the
    # rounding differences (the exact shard == full-batch statement is tests/test_multirank_cpu.py, one step): loose bound only
    assert worst <= 48, worst
else:
    assert [b for b, _ in got] == [rank], got
dist.barrier()
dist.destroy_process_group()
