"""``use_augs`` of the reference's MakeCutouts (cgd/modules.py:12-24, 62): per-cutout random parameters, host side.

Every crop ``input[:, :, oy:oy+S, ox:ox+S]`` (values in [0, 1]) goes through the torchvision pipeline

    RandomHorizontalFlip(0.5) -> +N(0, .01) -> RandomAffine(degrees=15, translate=(.1, .1)) -> +N(0, .01)
    -> RandomPerspective(distortion_scale=.4, p=.7) -> +N(0, .01) -> RandomGrayscale(.15) -> +N(0, .01)

before ``adaptive_avg_pool2d`` (cgd/modules.py:62-63).  The pipeline is linear in the image (plus additive noise), so forward and
input gradient are one gather / one scatter kernel (``cutouts_aug_fwd / _bwd``, csrc/augs.cu) driven by 20 numbers per cutout that
this module draws **in torchvision's own order from the same generators**: the decisions and geometric parameters from the CPU
default generator (``torch.rand(1)``, ``torch.empty(1).uniform_``, ``torch.randint`` inside the transforms), the four noise fields
with ``torch.randn`` on the image's device in the reference's shapes ``[B, 3, S, S]`` -- so a run with the same seed consumes both
generators exactly like the reference does.

Parameter layout per cutout (float32 x 20, ``AUG_NP``):
    0      flip (0 / 1)
    1..6   inverse affine matrix of torchvision (F._get_inverse_affine_matrix, centre (0, 0) in centred pixel coordinates)
    7      perspective applied (0 / 1)
    8..15  perspective coefficients a..h (F._get_perspective_coeffs: least squares in float64, stored as float32)
    16     grayscale (0 / 1)
    17..19 reserved
"""
from __future__ import annotations

import math

import torch as th

AUG_NP = 20
NOISE_STD = 0.01  # the four tvt.Lambda(lambda x: x + th.randn_like(x) * 0.01) stages
DEGREES, TRANSLATE, DISTORTION, P_PERSP, P_GRAY, P_FLIP = 15.0, (0.1, 0.1), 0.4, 0.7, 0.15, 0.5


def inverse_affine_matrix(angle: float, translate, scale: float = 1.0, shear=(0.0, 0.0), center=(0.0, 0.0)):
    """torchvision.transforms.functional._get_inverse_affine_matrix (inverted=True): M^-1 = C * RSS^-1 * C^-1 * T^-1"""
    rot, sx, sy = math.radians(angle), math.radians(shear[0]), math.radians(shear[1])
    cx, cy = center
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d / scale, -b / scale, 0.0, -c / scale, a / scale, 0.0]
    m[2] += m[0] * (-cx - tx) + m[1] * (-cy - ty)
    m[5] += m[3] * (-cx - tx) + m[4] * (-cy - ty)
    m[2] += cx
    m[5] += cy
    return m


def perspective_coeffs(startpoints, endpoints):
    """torchvision.transforms.functional._get_perspective_coeffs: output pixel (x, y) samples the input at
    ((ax + by + c) / (gx + hy + 1), (dx + ey + f) / (gx + hy + 1))"""
    A = th.zeros(8, 8, dtype=th.float64)
    for i, (p1, p2) in enumerate(zip(endpoints, startpoints)):
        A[2 * i, :] = th.tensor([p1[0], p1[1], 1, 0, 0, 0, -p2[0] * p1[0], -p2[0] * p1[1]], dtype=th.float64)
        A[2 * i + 1, :] = th.tensor([0, 0, 0, p1[0], p1[1], 1, -p2[1] * p1[0], -p2[1] * p1[1]], dtype=th.float64)
    bvec = th.tensor(startpoints, dtype=th.float64).view(8)
    return th.linalg.lstsq(A, bvec, driver="gels").solution.to(th.float32).tolist()


def _draw_affine(width: int, height: int):
    """RandomAffine.get_params(degrees=(-15, 15), translate=(.1, .1), scale=None, shear=None): three CPU-generator draws"""
    angle = float(th.empty(1).uniform_(-DEGREES, DEGREES).item())
    max_dx, max_dy = float(TRANSLATE[0] * width), float(TRANSLATE[1] * height)
    tx = int(round(th.empty(1).uniform_(-max_dx, max_dx).item()))
    ty = int(round(th.empty(1).uniform_(-max_dy, max_dy).item()))
    return angle, (tx, ty)


def _draw_perspective(width: int, height: int):
    """RandomPerspective.get_params(width, height, 0.4): eight CPU-generator randint draws in torchvision's order"""
    hh, hw = height // 2, width // 2
    dw, dh = int(DISTORTION * hw), int(DISTORTION * hh)

    def ri(lo, hi):
        return int(th.randint(lo, hi, size=(1,)).item())

    topleft = [ri(0, dw + 1), ri(0, dh + 1)]
    topright = [ri(width - dw - 1, width), ri(0, dh + 1)]
    botright = [ri(width - dw - 1, width), ri(height - dh - 1, height)]
    botleft = [ri(0, dw + 1), ri(height - dh - 1, height)]
    start = [[0, 0], [width - 1, 0], [width - 1, height - 1], [0, height - 1]]
    return start, [topleft, topright, botright, botleft]


def draw_aug_params(coords, B: int, H: int, W: int, noise_device="cpu", noise_out: th.Tensor = None, rows=None):
    """Parameters [cutn, AUG_NP] (float32, CPU) and the four noise fields of every cutout for one MakeCutouts.forward call, consuming
    the generators in the reference's order: for each cutout (cgd/modules.py:60-64) flip decision, noise, affine parameters, noise,
    perspective decision (+ its eight corner draws), noise, grayscale decision, noise.

    ``noise_out``: [cutn, 4, B, 3, Smax, Smax] fp32 buffer on ``noise_device`` receiving ``randn(B, 3, Sy, Sx) * 0.01`` in its top-left
    corner (None: no noise is drawn -- tests of the geometry alone).  ``rows = (lo, hi)``: the batch is sharded over ranks -- the noise is
    drawn for all ``B`` images like the single-process reference does and rows [lo, hi) are kept (``noise_out`` holds hi - lo images)."""
    params = th.zeros(len(coords), AUG_NP)
    for k, (ox, oy, S) in enumerate(coords):
        Sy, Sx = min(S, H - oy), min(S, W - ox)  # slices clip at the border like the reference's indexing (quirk B3)

        def noise(stage):
            if noise_out is not None:
                full = th.randn(B, 3, Sy, Sx, device=noise_device) * NOISE_STD
                noise_out[k, stage, :, :, :Sy, :Sx] = full if rows is None else full[rows[0]:rows[1]]

        p = params[k]
        p[0] = float(bool(th.rand(1) < P_FLIP))
        noise(0)
        angle, trans = _draw_affine(Sx, Sy)
        p[1:7] = th.tensor(inverse_affine_matrix(angle, [float(trans[0]), float(trans[1])]))
        noise(1)
        if th.rand(1) < P_PERSP:
            p[7] = 1.0
            p[8:16] = th.tensor(perspective_coeffs(*_draw_perspective(Sx, Sy)))
        noise(2)
        p[16] = float(bool(th.rand(1) < P_GRAY))
        noise(3)
    return params
