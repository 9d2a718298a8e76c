"""CPU: the C-ABI never hangs, aborts or "succeeds" on malformed ops (errors are return codes -- negative = invalid
argument, positive = cudaError_t -- "never exit/abort, never throw across the ABI").

Every distinct (op code, flags) of the guided-step plans is cloned and one field at a time is replaced -- each of the 24 integers by
0, -1, small odd values, 2^31 - 1 and +-2^40, each pointer by NULL and by a misaligned address, the flags by out-of-range masks -- and
handed to `cgd_run_op` without a device: the call must return non-zero (a rejected argument, or the launch failing for lack of a GPU)
within the time limit.  The children run in their own process so that a hang or a crash is a test failure, not a stuck suite.  (First
run of this test found one: a 2^31-row image spun forever in the conv tile search -- now rejected as "dims out of range".)"""
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch as th

pytestmark = pytest.mark.skipif(th.cuda.is_available(), reason="needs a box WITHOUT a GPU: only the argument checks may run")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = textwrap.dedent('''
    import ctypes, json, sys
    sys.path.insert(0, sys.argv[1])
    from clip_guided_diffusion_b200 import _lib
    from tests.plan_interp import CODE
    from tests.step_parity import build_tiny
    lib = _lib.load()
    plan = build_tiny("cpu", **json.loads(sys.argv[2]))["eng"].plan
    arr = plan.lower(1 << 40)
    vals = [0, -1, 1, 3, 7, (1 << 31) - 1, 1 << 40, -(1 << 40)]
    seen, n, zero = set(), 0, []
    for k, op in enumerate(plan.ops):
        kind = (op.code, op.flags)
        if kind in seen:
            continue
        seen.add(kind)
        base = arr[k]
        muts = [("i", j, v) for j in range(24) for v in vals if base.i[j] != v]
        muts += [("p", j, v) for j in range(12) for v in (0, 3) if base.p[j]] + [("flags", 0, v) for v in (-1, 1 << 30, 255)]
        for w, j, v in muts:
            o = type(base)()
            ctypes.memmove(ctypes.byref(o), ctypes.byref(base), ctypes.sizeof(base))
            if w == "i":
                o.i[j] = v
            elif w == "p":
                o.p[j] = 0 if v == 0 else o.p[j] + v
            else:
                o.flags = v
            print("TRY", CODE[op.code], op.flags, w, j, v, flush=True)
            n += 1
            if lib.cgd_run_op(ctypes.byref(o), None) == 0:
                zero.append((CODE[op.code], op.flags, w, j, v))
    print("DONE", n, len(seen), json.dumps(zero), flush=True)
''')

CONFIGS = [dict(B=2, cutn=3, image=32), dict(B=1, cutn=4, image=32, use_magnitude=True, sat_scale=30.0, cutn_variants=(2, 4)),
           dict(B=2, cutn=2, image=32, init_scale=1000.0), dict(B=1, cutn=3, image=64, cutout_resize="lanczos3"),
           dict(B=2, cutn=3, image=32, tower="rn"), dict(B=1, cutn=2, image=32, use_augs=True)]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items() if k not in ("B", "cutn", "image")) or "plain")
def test_malformed_ops_return_error_codes(cfg):
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, ROOT, json.dumps(cfg)], capture_output=True, text=True, timeout=180)
    except subprocess.TimeoutExpired as e:
        last = [l for l in (e.stdout or b"").decode().splitlines() if l.startswith("TRY")][-1:]
        pytest.fail(f"cgd_run_op hung on {last}")
    lines = r.stdout.splitlines()
    tries = [l for l in lines if l.startswith("TRY")]
    assert r.returncode == 0 and lines and lines[-1].startswith("DONE"), f"crashed (rc {r.returncode}) on {tries[-1:]}: {r.stderr[-400:]}"
    _, n, kinds, zero = lines[-1].split(" ", 3)
    assert int(n) > 3000 and int(kinds) >= 20
    assert json.loads(zero) == [], f"rc 0 without a device: {json.loads(zero)[:5]}"


ENTRY_CHILD = r"""
import ctypes, sys, itertools
sys.path.insert(0, sys.argv[1])
from clip_guided_diffusion_b200 import _lib
_lib.load()
import os
lib = ctypes.CDLL(os.path.join(sys.argv[1], 'clip_guided_diffusion_b200', 'libcgd_b200.so'))  # no argtypes: raw C calls
lib.cgd_last_error.restype = ctypes.c_char_p
P = lambda v: ctypes.c_void_p(v)
good = 1 << 40
f3 = (ctypes.c_float * 3)(0.5, 0.5, 0.5)
f3p = ctypes.cast(f3, ctypes.c_void_p)
I = ctypes.c_int64
F = ctypes.c_float
def call(name, args):
    fn = getattr(lib, name); fn.restype = ctypes.c_int
    print("TRY", name, [getattr(a, 'value', a) for a in args], flush=True)
    rc = fn(*args)
    if rc == 0: print("ZERO", name, [getattr(a, 'value', a) for a in args], flush=True)
    return rc
dims = [0, -1, 1, 7, 224, (1 << 31) - 1, 1 << 40]
ptrs = [0, good, good + 3]
# cutouts fwd: x, coords, patches, B,H,W,cutn,cs,patch,kpad, mean, std, stream
base = [P(good), P(good), P(good), I(1), I(64), I(64), I(2), I(32), I(16), I(768), f3p, f3p, P(0)]
for j in range(13):
    for v in ([0] if j in (10, 11) else ptrs if j in (0, 1, 2) else dims if j < 12 else [0]):
        a = list(base); a[j] = (P(v) if j in (0, 1, 2, 10, 11, 12) else I(v)); call("cgd_cutouts_fwd", a)
base = [P(good), P(good), P(good), I(1), I(64), I(64), I(2), I(32), I(16), I(768), f3p, F(1.0), P(0)]
for j in range(12):
    for v in ([0] if j == 10 else ptrs if j in (0, 1, 2) else dims):
        a = list(base); a[j] = (P(v) if j in (0, 1, 2, 10) else I(v)); call("cgd_cutouts_bwd", a)
base = [P(good)] * 5 + [I(4), I(1), I(1), I(64), F(1000.0), F(1.0), P(0)]
for j in range(9):
    for v in (ptrs if j < 5 else dims):
        a = list(base); a[j] = (P(v) if j < 5 else I(v)); call("cgd_spherical_fwd_bwd", a)
base = [P(good)] * 7 + [I(1), I(32), I(32), I(8), F(1), F(1), F(0), F(1), P(0)]
for j in range(11):
    for v in (ptrs if j < 7 else dims):
        a = list(base); a[j] = (P(v) if j < 7 else I(v)); call("cgd_guidance_losses_fwd_bwd", a)
base = [P(good)] * 7 + [I(1024), P(0)]
for j in range(8):
    for v in (ptrs if j < 7 else dims):
        a = list(base); a[j] = (P(v) if j < 7 else I(v)); call("cgd_sample_update_ancestral", a)
base = [P(good)] * 6 + [I(1024), P(0)]
for j in range(7):
    for v in (ptrs if j < 6 else dims):
        a = list(base); a[j] = (P(v) if j < 6 else I(v)); call("cgd_sample_update_ddim", a)
# handles
h = ctypes.c_void_p()
for name in ("cgd_plan_create", "cgd_step_create"):
    for n in (0, -1, 5):
        call(name, [P(0), ctypes.c_int32(n), ctypes.byref(h)])
    call(name, [P(0), ctypes.c_int32(1), P(0)])
for name in ("cgd_unet_create", "cgd_vit_create"):
    for nf, nb in ((0, 0), (-1, 2), (2, -1), (1, 1)):
        call(name, [P(0), ctypes.c_int32(nf), ctypes.c_int32(nb), ctypes.byref(h)])
for name in ("cgd_plan_destroy", "cgd_unet_destroy", "cgd_vit_destroy", "cgd_step_destroy", "cgd_unet_fwd", "cgd_unet_bwd_input", "cgd_vit_fwd", "cgd_vit_bwd_input", "cgd_step"):
    args = [P(0)] if "destroy" in name else [P(0), P(0)]
    call(name, args)
call("cgd_plan_run", [P(0), ctypes.c_int32(0), ctypes.c_int32(1), P(0)])
call("cgd_plan_num_launches", [P(0), ctypes.c_int32(0), ctypes.c_int32(1)])
call("cgd_run_op", [P(0), P(0)])
for bn, sp in ((0, 0), (-1, 2), (256, 4), (1 << 30, 1 << 30)):
    print("cap", bn, sp, lib.cgd_conv_cluster_capacity(ctypes.c_int32(bn), ctypes.c_int32(sp)))
print("DONE")
"""


def test_stand_alone_entry_points_reject_malformed_arguments():
    """the same for the named entry points of include/cgd_b200.h (cutouts, spherical, guidance losses, sampler updates, handle
    create / run / destroy): NULL and misaligned device pointers, NULL host triples, zero / negative / 2^31 - 1 / 2^40 sizes, NULL
    handles.  Only `*_destroy(NULL)` may return 0 (like free(NULL)); `cgd_conv_cluster_capacity` answers -1 without a device."""
    r = subprocess.run([sys.executable, "-c", ENTRY_CHILD, ROOT], capture_output=True, text=True, timeout=180)
    lines = r.stdout.splitlines()
    assert r.returncode == 0 and lines and lines[-1] == "DONE", f"crashed (rc {r.returncode}) after {[l for l in lines if l.startswith('TRY')][-1:]}: {r.stderr[-400:]}"
    assert len([l for l in lines if l.startswith("TRY")]) >= 290
    zero = [l for l in lines if l.startswith("ZERO") and "_destroy" not in l]
    assert not zero, zero[:5]
    assert all(l.split()[-1] == "-1" for l in lines if l.startswith("cap "))
