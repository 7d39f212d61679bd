"""CPU oracle for the INDM hot path: a restatement of the reference's algorithm, used ONLY as a checker.

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import anything from this package.  The product (`indm_b200/`) never does, and
fails loudly if its CUDA library is missing — there is no CPU fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the restatement is
pinned against outputs of the reference itself executed in the build container
(`tests/golden/make_golden.py` imports `/root/reference` through `oracle/ref_loader.py` and writes the committed
fixtures `tests/golden/*.npz`); `tests/test_oracle_golden.py` checks every oracle function against them.

Modules
  ops.py      upfirdn2d / fused_leaky_relu         (op/upfirdn2d.py, op/fused_act.py)
  sde.py      VP / VE SDE math                     (sde_lib.py)
  ncsnpp.py   NCSN++ / DDPM++ forward, score_fn    (models/ncsnpp.py, layerspp.py, layers.py, utils.py)
  sampler.py  predictor / corrector / pc_sampler   (sampling.py)
  flow.py     wolf flow: posterior encoder, prior flow, iResBlocks (forward / inverse / log-det series / differentiable
              training forward)                    (flow_models/wolf/**: wolf.py, resflow_.py, iresblock.py, priors/flow.py ...)
  ref_loader.py + ref_stubs/   import shim for the unmodified reference (build container only)
"""
