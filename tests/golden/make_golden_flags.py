"""Golden copy of the reference's checkpoint table and CLIP model list, produced by importing the REFERENCE's own data modules
(data/diffusion_model_flags.py: DIFFUSION_LOOKUP; cgd/clip_util.py:17-29 read as text -- the module itself does not import without
`clip`).  Only the fields the hot path depends on are kept (file names and shape-defining model flags; no URLs).  Run in the build
container only (needs /root/reference):

    python tests/golden/make_golden_flags.py        # writes tests/golden/model_flags_golden.json (committed)
"""
import ast
import json
import os
import sys

sys.path.insert(0, "/root/reference")
from data.diffusion_model_flags import DIFFUSION_LOOKUP  # noqa: E402

out = {"diffusion": {}, "clip": {}}
for cond, table in DIFFUSION_LOOKUP.items():
    for size, entry in table.items():
        out["diffusion"][f"{cond}/{size}"] = {"filename": entry["filename"], "model_flags": entry["model_flags"]}
src = open("/root/reference/cgd/clip_util.py").read()
tree = ast.parse(src)
for node in tree.body:
    if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "CLIP_MODEL_NAMES":
        out["clip"]["names"] = list(ast.literal_eval(node.value))
    if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "CLIP_MODEL_URLS":
        out["clip"]["files"] = {k: v.rsplit("/", 1)[1] for k, v in ast.literal_eval(node.value).items()}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_flags_golden.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(sorted(out["diffusion"]), out["clip"])
