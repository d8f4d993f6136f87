"""CPU/GPU fp32 restatement of the CCDM_unified denoiser hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ccdm_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may.  The product path never routes through it.

Parity status: the reference ships no tests, fixtures or golden vectors for
this path, so the oracle is pinned against outputs of the
reference's own modules imported from ``/root/reference`` in the build
container: ``tests/golden/make_golden.py`` produced ``tests/golden/*.pt`` and
``tests/test_oracle_golden.py`` replays them on every CPU test run.
"""
from .unet_ref import UnetSpec, unet_forward, unet_forward_cfg, make_state_dict  # noqa: F401
from .diffusion_ref import (  # noqa: F401
    Schedule,
    make_schedule,
    q_sample,
    model_predictions,
    ddim_sample,
    ddpm_sample,
    p_losses,
    y2h_sinusoidal,
    y2cov_sinusoidal,
)
