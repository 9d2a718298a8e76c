"""CPU: the LPIPS-VGG16 op list (clip_guided_diffusion_b200/lpips.py), interpreted in PyTorch, against the oracle restatement of
``lpips.LPIPS(net='vgg')`` (oracle/lpips.py): loss value and the gradient the reference adds through
``lpips_vgg(x_in, init_tensor).sum() * init_scale`` (cgd/cgd.py:220-224)."""
import torch as th

from clip_guided_diffusion_b200.lpips import LpipsB200
from clip_guided_diffusion_b200.plan import Plan
from oracle import lpips as ol
from tests.plan_interp import Interp


def build(device, B=2, H=32, W=48, init_scale=1000.0):
    sd = ol.seeded_state_dict()
    plan = Plan()
    x_src = plan.new(B * 3 * H * W, "f", "x_in")
    g_dst = plan.new(B * 3 * H * W, "f", "g")
    lp = LpipsB200(sd, B, H, W, plan, x_src, g_dst, init_scale)
    plan.finalize(device)
    return sd, plan, lp, x_src, g_dst


def reference(sd, x, init, init_scale):
    net = ol.LPIPSVgg(sd)
    xg = x.clone().requires_grad_()
    val = net(xg, init)  # [B,1,1,1], init broadcast over the batch like the reference call
    (g,) = th.autograd.grad(val.sum() * init_scale, xg)
    return val.detach().flatten(), g


def test_parameter_inventory():
    sh = ol.param_shapes()
    assert sum(int(th.tensor(v).prod()) for k, v in sh.items() if k.startswith("net.")) == 14714688  # VGG16 features (published)
    assert sum(int(th.tensor(v).prod()) for k, v in sh.items() if k.startswith("lin")) == 64 + 128 + 256 + 512 + 512


def test_lpips_plan_matches_oracle():
    B, H, W, scale = 2, 32, 48, 1000.0
    sd, plan, lp, x_src, g_dst = build("cpu", B, H, W, scale)
    it = Interp(plan)
    th.manual_seed(0)
    init = th.rand(1, 3, H, W) * 2 - 1
    x = (th.rand(B, 3, H, W) * 2 - 1) * 0.9
    lp.set_init_image(init, runner=it.run_range)
    plan.view(x_src, (B, 3, H, W)).copy_(x)
    plan.view(g_dst, (B, 3, H, W)).fill_(0.25)  # the accumulator already holds the CLIP gradient in the engine
    it.run_range("lpips", "lpips_end")
    val, g = reference(sd, x, init, scale)
    got_val = lp.loss_value().clone()
    got_g = plan.view(g_dst, (B, 3, H, W)) - 0.25
    # fp16 activations / weights through 13 conv + ReLU layers vs the fp32 oracle
    assert th.allclose(got_val, val, rtol=3e-2, atol=1e-4), (got_val, val)
    cosg = float(th.nn.functional.cosine_similarity(got_g.flatten(), g.flatten(), dim=0))
    relg = float((got_g - g).norm() / g.norm())
    assert cosg > 0.998 and relg < 6e-2, (cosg, relg)
