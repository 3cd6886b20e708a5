"""CPU oracle for the guided reverse-diffusion hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a from-scratch, functional (state_dict in, tensor out) fp32 restatement of
the reference algorithm for the path named in BASELINE.json `north_star`:

  * oracle/unet.py      — `unet_fast` / `unetca_fast` eps-prediction and the CFG
                          wrapper (reference: dynamic/diffusionmodules/openaimodel.py,
                          openaimodel_ca.py, dynamic/crossattetion_lr.py,
                          dynamic/diffusionmodules/util.py)
  * oracle/schedule.py  — DDPM / DDIM schedule tables (reference:
                          diffusion/sampler/ddpm_sampler.py:25-103,
                          ddim_plms_sampler.py:38-81, diffusionmodules/util.py:23-74)
  * oracle/sampler.py   — DDPM ("native"), DDIM and PLMS loops driven by a
                          host-supplied noise tape (reference: ddpm_sampler.py:154-238,
                          ddim_plms_sampler.py:302-525, diffusion/ddpm.py:108-122)
  * oracle/condition.py — condition lookup (reference: dynamic_input/condition.py:5-157)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it, and only as the checker or the timed CPU baseline.
The product (`sgdm_b200`) never imports it and fails loudly without its CUDA library.

PARITY PIN: the reference ships no golden vectors for this path, so
the oracle is pinned against outputs of the UNMODIFIED reference executed in the
authoring container: tests/golden/make_golden.py imports /root/reference, runs it on
seeded inputs and commits the results under tests/golden/*.npz;
tests/test_oracle_golden.py checks this restatement against them (schedules and
indices bit-exact, floating point to 1e-5 relative).

All arithmetic is torch fp32 on CPU (the reference runs `precision: 32`); numpy
float64 where the reference uses numpy float64 (schedule construction).
"""
