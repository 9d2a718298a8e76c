"""Known answers of the reference's entry helpers, produced by EXECUTING the reference's own function bodies: `cgd/script_util.py`
does not import here (guided_diffusion is absent), so `parse_prompt`, `alphanumeric_filter`, `clean_and_combine_prompts` and `log_image`
(cgd/script_util.py:60-67, 81-101) are cut out of its source with `ast` and run unmodified in a namespace holding the modules they use.
Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_script_util.py        # writes tests/golden/script_util_golden.json (committed)
"""
import ast
import json
import os
import re
import tempfile

import numpy as np
import torch as th
import torchvision.transforms.functional as tvf
from PIL import Image

src = open("/root/reference/cgd/script_util.py").read()
want = ("parse_prompt", "alphanumeric_filter", "clean_and_combine_prompts", "log_image")
ns = {"os": os, "re": re, "th": th, "tvf": tvf}
for node in ast.parse(src).body:
    if isinstance(node, ast.FunctionDef) and node.name in want:
        exec(compile(ast.Module([node], []), "cgd/script_util.py", "exec"), ns)

prompts = ["Loose seal.:0.4", "Loose seal.:-0.4", "Loose seal.", "a:b:2", "https://a.b/c.png:2", "https://a.b/c.png", "http://x.y/z:0.5", "12:30 at night:3",
           "trailing colon:1", "weight only:1e-2"]
out = {"parse_prompt": [[p, *ns["parse_prompt"](p)] for p in prompts]}
texts = [["a cat: 0.5/x!", "b"], ["x" * 400], ["hello world", "ünïcode dog?", "tabs\tand  spaces"], ["under_score-dash.dot"], [""]]
out["clean_and_combine_prompts"] = [[t, b, ns["clean_and_combine_prompts"]("o", t, b)] for t in texts for b in (0, 7)]
g = th.Generator().manual_seed(3)
img = th.cat([th.rand(3, 6, 8, generator=g) * 2.6 - 1.3, th.linspace(-1, 1, 24).view(1, 1, 24).expand(3, 1, 24).reshape(3, 3, 8)], dim=1)  # some values outside [-1, 1]
cwd = os.getcwd()
with tempfile.TemporaryDirectory() as d:
    os.chdir(d)
    path = ns["log_image"](img, "out", ["a b", "c!"], 12, 3)
    out["log_image"] = {"image": img.numpy().astype(np.float64).tolist(), "path": path, "pixels": np.asarray(Image.open(path)).tolist(),
                        "current_png": os.path.exists("current.png")}
    os.chdir(cwd)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "script_util_golden.json"), "w"))
print(out["parse_prompt"][:4], out["clean_and_combine_prompts"][:2], out["log_image"]["path"])
