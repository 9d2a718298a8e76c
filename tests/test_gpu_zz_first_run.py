"""GPU: cases committed after the round's last device call (no GPU minutes left to run them), kept in the file pytest visits last so
that under `-x` they cannot hide the result of a test that has already passed on a B200.  Same procedure and tolerances as
tests/test_gpu_baseline_configs.py (teacher-forced guided step vs the fp32 oracle on the same GPU).

    vit_l14_336  the one CLIP_MODEL_URLS entry (cgd/clip_util.py:28) with 336 px cutouts: 24 x 24 patches of 14 + class token = 577 tokens"""
import pytest

from tests.test_gpu_baseline_configs import FIRST_RUN_PENDING, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", FIRST_RUN_PENDING)
def test_full_size_step_vs_fp32_oracle_first_run(name):
    run_case(name)
