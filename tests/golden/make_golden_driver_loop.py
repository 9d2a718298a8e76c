"""Which frames the reference's driver loop saves, by executing its own statements (cgd/cgd.py:241-271).

The loop lives in the body of `clip_guided_diffusion` (a module that does not import here), so its statements -- the sampler choice, the
`try:` block with the `current_timestep` bookkeeping and the save rule -- are cut out with `ast` and compiled unmodified into a generator
whose free variables are stand-ins: a diffusion whose `*_sample_loop_progressive` yields `num_timesteps - skip_timesteps` samples (what
guided-diffusion's loop does, [3P]) and a `script_util.log_image` that records (step, batch_idx).

    python tests/golden/make_golden_driver_loop.py        # writes tests/golden/driver_loop_golden.json (committed); needs /root/reference
"""
import ast
import json
import os
import types

import torch as th

tree = ast.parse(open("/root/reference/cgd/cgd.py").read())
outer = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "clip_guided_diffusion")
start = next(i for i, n in enumerate(outer.body) if isinstance(n, ast.If) and "timestep_respacing" in ast.unparse(n.test) and "ddim" in ast.unparse(n.test))
stop = next(i for i, n in enumerate(outer.body) if isinstance(n, ast.Try))
body = outer.body[start:stop + 1]
fn = ast.FunctionDef(name="_driver", args=ast.arguments(posonlyargs=[], args=[], kwonlyargs=[], kw_defaults=[], defaults=[]), body=body, decorator_list=[],
                     type_params=[])
mod = ast.fix_missing_locations(ast.Module([fn], []))
code = compile(mod, "cgd/cgd.py", "exec")


def run(respacing, T, skip, save_frequency, B):
    calls = []

    def loop(kind):
        def gen(model, shape, **kw):
            assert kw["skip_timesteps"] == skip and kw["cond_fn_with_grad"] is True and kw["clip_denoised"] is False
            calls.append(kind)
            for _ in range(T - skip):
                yield {"pred_xstart": th.zeros(shape), "sample": th.zeros(shape)}
        return gen

    ns = dict(timestep_respacing=respacing, gd_model=None, batch_size=B, image_size=8, height_offset=0, width_offset=0, model_kwargs={}, cond_fn=None,
              progress=False, skip_timesteps=skip, init_tensor=None, randomize_class=True, save_frequency=save_frequency, prefix_path="out", prompts=["p"],
              clip_model_name="ViT-B/32", diffusion=types.SimpleNamespace(num_timesteps=T, p_sample_loop_progressive=loop("ancestral"),
                                                                          ddim_sample_loop_progressive=loop("ddim")),
              script_util=types.SimpleNamespace(log_image=lambda image, prefix, prompts, step, batch_idx: f"{step}/{batch_idx}"))
    exec(code, ns)
    got = [(b, p) for b, p in ns["_driver"]()]
    return {"sampler": calls, "yields": [[b, int(p.split("/")[0])] for b, p in got]}


out = []
for respacing, T in (("25", 25), ("ddim25", 25)):
    for skip, sf, B in ((0, 1, 1), (0, 7, 2), (22, 2, 1), (20, 25, 1), (24, 25, 2), (0, 25, 1), (10, 5, 1), (23, 1, 2)):
        out.append({"respacing": respacing, "T": T, "skip": skip, "save_frequency": sf, "batch": B, **run(respacing, T, skip, sf, B)})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "driver_loop_golden.json"), "w"))
for o in out[:8]:
    print(o["respacing"], o["skip"], o["save_frequency"], o["batch"], o["sampler"], [s for _, s in o["yields"]][:12])
