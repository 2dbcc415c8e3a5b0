"""CPU oracle for the PSO training-step hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
reported CPU baseline -- never as the thing measured or shipped.  The product
package ``pairwise_sample_optimization_b200`` must not import this package.

What is in here (each function cites the reference file:line it follows;
paths are relative to ``/root/reference``):

* ``reference_loader``  -- executes the reference's own code *verbatim from
  /root/reference*: the two step files under a tiny ``diffusers`` import stub,
  and -- cut out of the trainer files at run time, since the trainers cannot be
  imported -- the micro-step loss lines of both online trainers (turbo :810-850,
  dmd2 :812-854), their ``sample_compare`` / ``compare`` functions, the
  DreamBooth trainer's loss lines (:1846-1935) and the denoising loops of the
  two sampler pipelines (sdxl_turbo_with_logprob.py :112-151,
  sdxl_dmd_with_logprob.py :109-164).  Only usable in the build
  container (the GPU box has no /root/reference); it is what pins the
  restatement (``oracle/make_golden.py`` -> ``tests/golden``).
* ``schedules``  -- the diffusers==0.27.0 scheduler constants the path reads
  (third-party, absent from /root/reference; restated from the published
  formula, anchored on sigma(999)=14.6146).
* ``steps``      -- torch-fp32 restatement of the two step functions, plus an
  fp64 closed form.
* ``losses``     -- restated inline online-PSO loss, ``sample_compare`` /
  ``compare`` and the DreamBooth-PSO loss.
* ``lora``       -- restated peft==0.11.1 ``lora.Linear`` forward and the
  diffusers==0.27.0 ``AttnProcessor2_0`` data flow (third-party, restated).
* ``samplers``   -- the two few-step sampler loops.

Parity pinning status: the reference ships NO tests, golden vectors or
fixtures, so the pin is "outputs of the reference itself
run here": ``tests/golden/*.npz`` were produced by ``oracle/make_golden.py``
executing the reference's own step functions AND the trainers' own loss lines
(see ``reference_loader``), with the restatements in ``steps`` / ``losses``
asserted bit-identical (loss and autograd gradients), and
``tests/test_oracle_golden.py`` re-checks both against the fixtures.  What stays
"parity unpinned" is the arithmetic of the un-vendored third-party packages
(diffusers 0.27.0 schedule constants, peft 0.11.1 ``lora.Linear``, diffusers
``AttnProcessor2_0`` -- ``schedules`` / ``lora``): restated from the published
behaviour and anchored on the reference's call sites and on sigma(999)=14.6146.
"""
