"""CPU, gloo, world_size 2: batch sharding of the guided step.  Each rank owns one image, draws the full-batch
noise / classes with the shared seed and keeps its rows; the all-gathered samples must equal the single-process batch-2 step.
Kernels are interpreted (tests/plan_interp.py); this checks the host-side sharding logic only."""
import os

import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.plan_interp import Interp
from tests.step_parity import build_tiny


def _one_step(ctx, seed=11, mode="ddim", scale_rows=False):
    eng, pdiff = ctx["eng"], ctx["pdiff"]
    it = Interp(eng.plan)
    th.manual_seed(seed)
    x = eng.draw_initial_noise()
    if scale_rows:  # make the images (and their gradient norms) clearly different: global row k is scaled by 1 + 2k
        for i in range(eng.B):
            x[i] *= 1.0 + 2.0 * (eng.rank * eng.B + i)
    y = eng.draw_classes()
    coords = [(3, 5, 24), (0, 2, 30), (7, 1, 20)]
    eng.stage_step(pdiff.scalar_table(14, 14, 0.0), coords, pdiff.model_timestep(14), y)
    eng.img(eng.unet.x_in).copy_(x)
    eng.draw_noise()
    eng._run_all(mode, it.run_range)
    return th.cat([eng.img(eng.sample), x, eng.img(eng.noise), y.view(-1, 1, 1, 1).expand(-1, 3, 32, 32).float(), eng.img(eng.g)], dim=1).clone()


def _worker(rank, world, port, ret, kw=None, mode="ddim"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    th.set_num_threads(2)
    ctx = build_tiny("cpu", B=1, cutn=3, image=32, rank=rank, world_size=world, **(kw or {}))
    s = _one_step(ctx, mode=mode, scale_rows=bool(kw))
    out = [th.empty_like(s) for _ in range(world)]
    dist.all_gather(out, s)
    if rank == 0:
        ret["gathered"] = th.cat(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_equals_full_batch():
    ref = _one_step(build_tiny("cpu", B=2, cutn=3, image=32))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, ret), nprocs=2, join=True)
    got = ret["gathered"]
    assert got.shape == ref.shape
    # inputs of every rank (x_T rows, per-step noise rows, classes) are bit-identical to the full-batch draws ...
    assert th.equal(got[:, 3:12], ref[:, 3:12])
    # ... and the step output agrees up to fp16 rounding noise (different batch shapes pick different CPU conv algorithms)
    rel = float((got[:, :3] - ref[:, :3]).norm() / ref[:, :3].norm())
    assert rel < 1e-2, rel


def test_two_rank_whole_batch_rms_clamp():
    """use_magnitude clamps by the RMS of the WHOLE batch (cgd/cgd.py:229-232): two ranks all-reduce their partial sums of squares
    between FINAL_GRAD and MAG_CLAMP and must reproduce the single-process batch-2 step (whose two images share one clamp)."""
    kw = dict(use_magnitude=True)
    ref = _one_step(build_tiny("cpu", B=2, cutn=3, image=32, **kw), mode="ancestral", scale_rows=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29519, ret, kw, "ancestral"), nprocs=2, join=True)
    got = ret["gathered"]
    assert th.equal(got[:, 3:12], ref[:, 3:12])
    rel = float((got[:, :3] - ref[:, :3]).norm() / ref[:, :3].norm())
    assert rel < 1e-2, rel
    # the clamp really is global: one factor for both images keeps the ratio of their gradient norms (a per-rank clamp would bring
    # both to the same RMS), and the whole-batch RMS sits at the 0.05 cap
    g_ref, g_got = ref[:, 12:15], got[:, 12:15]
    r_ref = g_ref.flatten(1).norm(dim=1)
    r_got = g_got.flatten(1).norm(dim=1)
    assert abs(float(r_ref[0] / r_ref[1]) - 1.0) > 0.02, "degenerate case: both images have the same gradient norm"
    assert abs(float(r_got[0] / r_got[1]) / float(r_ref[0] / r_ref[1]) - 1.0) < 1e-2
    assert abs(float(g_got.square().mean().sqrt()) - 0.05) < 1e-3


def test_two_rank_sat_loss_is_a_whole_batch_mean():
    """sat loss = mean over the WHOLE batch (cgd/cgd.py:214-218): a rank holding B/G images must still divide by B_global"""
    kw = dict(sat_scale=500.0)
    ref = _one_step(build_tiny("cpu", B=2, cutn=3, image=32, **kw), scale_rows=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29521, ret, kw, "ddim"), nprocs=2, join=True)
    got = ret["gathered"]
    assert th.equal(got[:, 3:12], ref[:, 3:12])
    rel_g = float((got[:, 12:15] - ref[:, 12:15]).norm() / ref[:, 12:15].norm())
    assert rel_g < 1e-2, rel_g
