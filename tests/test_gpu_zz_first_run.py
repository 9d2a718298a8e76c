"""GPU, opt-in (CGD_TEST_FIRST_RUN=1): cases committed after the round's last device call (no GPU minutes were left to run them).
They add no kernel -- only a configuration of validated ones -- but a case that has never executed on a B200 does not belong in the
default suite; the file is also the last one pytest visits, so that under `-x` it cannot hide a test that has already passed.  Same procedure and tolerances as
tests/test_gpu_baseline_configs.py (teacher-forced guided step vs the fp32 oracle on the same GPU).

    vit_l14_336  the one CLIP_MODEL_URLS entry (cgd/clip_util.py:28) with 336 px cutouts: 24 x 24 patches of 14 + class token = 577 tokens"""
import os

import pytest

from tests.test_gpu_baseline_configs import FIRST_RUN_PENDING, run_case

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("CGD_TEST_FIRST_RUN") != "1",
                                                  reason="no device run yet; CGD_TEST_FIRST_RUN=1 to run")]


@pytest.mark.parametrize("name", FIRST_RUN_PENDING)
def test_full_size_step_vs_fp32_oracle_first_run(name):
    run_case(name)
