#!/usr/bin/env python
"""BASELINE configs[3] / configs[4] on one GPU (3x64x64, NCSN++ / DDPM++ num_res_blocks = 8, wolf flow with flow.squeeze), timed
with CUDA events; one JSON line per leg.  Not the benchmark of record (bench.py is): a side measurement for DESIGN.md §7.
  leg 1  ve/CELEBA/indm: PC sampling, reverse-diffusion predictor + Langevin corrector (2 NFE per step), `--pc-steps` steps of
         the 1000-step schedule per call (every step replays the same CUDA graph, so ms/step is the 1000-step figure / 1000)
  leg 2  vp/CELEBA/indm_nll: `flow_step_fn_nll` training step, flow + score trained jointly (falls back to training.freeze_flow
         and says so if the joint step raises)
    python tools/celeba_bench.py [--batch 64] [--pc-steps 50] [--train-steps 4] [--train-warmup 8]
"""
import argparse
import json
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs, sde_lib, sampling, losses, _lib as L  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from indm_b200.models.ema import ExponentialMovingAverage  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402

GFLOP_FWD = 142.9          # score-network forward per 64x64 image, nres = 8


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.launches
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, L.launches - l0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--pc-steps', type=int, default=50)
    ap.add_argument('--train-steps', type=int, default=4)
    ap.add_argument('--train-warmup', type=int, default=8)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    B = a.batch
    torch.manual_seed(0)

    # ---------------------------------------------------------------- leg 1: VE CelebA PC sampling with the Langevin corrector
    try:
        cfg = configs.get_config('ve/CELEBA/indm')
        cfg.model.num_res_blocks = 8
        cfg.device = dev
        cfg.sampling.num_scales = a.pc_steps
        model = mutils.create_model(cfg)
        flow = fm.create_flow_model(cfg)
        flow.eval()
        sde = sde_lib.get_sde(cfg)
        fn = sampling.get_sampling_fn(cfg, sde, (B, 3, 64, 64), lambda v: v, cfg.sampling.truncation_time)
        out = {}

        def call():
            out['r'] = fn(model, flow, seed=1)

        ms, launches = timed(call, 2, 2)
        before, after, nfe = out['r']
        nfe_call = 2 * a.pc_steps
        print(json.dumps({"leg": "ve/CELEBA/indm PC sampling (reverse_diffusion + langevin), nres=8, 64x64, flow.squeeze wolf inverse per call",
                          "batch": B, "pc_steps_per_call": a.pc_steps, "nfe_per_call": nfe_call, "ms_per_call": ms,
                          "ms_per_pc_step": ms / a.pc_steps, "images_per_sec_at_1000_steps": B / (ms * 1e-3 * 1000 / a.pc_steps),
                          "score_forward_tflops_algorithmic": GFLOP_FWD * B * nfe_call / (ms * 1e-3) / 1e3,
                          "gpu_launches_per_call": launches // 2, "finite": bool(torch.isfinite(after).all()),
                          "note": "flow inverse amortised over pc_steps here, over 1000 steps in a full run"}), flush=True)
        del model, flow, fn, out
        torch.cuda.empty_cache()
    except Exception:
        traceback.print_exc()
        print(json.dumps({"leg": "ve/CELEBA/indm PC sampling", "error": traceback.format_exc()[-400:]}), flush=True)

    # ---------------------------------------------------------------- leg 2: VP CelebA training step
    for freeze in (False, True):
        try:
            cfg = configs.get_config('vp/CELEBA/indm_nll')
            cfg.model.num_res_blocks = 8
            cfg.device = dev
            cfg.training.freeze_flow = freeze
            model = mutils.create_model(cfg)
            flow = fm.create_flow_model(cfg)
            sde = sde_lib.get_sde(cfg)
            opt = losses.get_optimizer(cfg, model.parameters())
            state = dict(optimizer=opt, model=model, ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
            if freeze:
                flow.eval()
                flow_state = dict(model=flow, step=0)
            else:
                fopt = losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr)
                flow_state = dict(optimizer=fopt, model=flow, ema=ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate), step=0)
            step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
            batch_host = (torch.rand(B, 3, 64, 64) * 2 - 1).pin_memory()
            out = {}

            def one():
                out['r'] = step_fn(state, flow_state, batch_host.to(dev, non_blocking=True))

            ms, launches = timed(one, a.train_steps, a.train_warmup)
            print(json.dumps({"leg": "vp/CELEBA/indm_nll flow_step_fn_nll training step, nres=8, 64x64" + (" (training.freeze_flow=True: score network only)" if freeze else " (JOINT flow + score)"),
                              "batch": B, "ms_per_step": ms, "samples_per_sec": B / (ms * 1e-3), "gpu_launches_per_step": launches // a.train_steps,
                              "score_fwd_bwd_tflops_algorithmic": 3 * GFLOP_FWD * B / (ms * 1e-3) / 1e3,
                              "loss_mean": float(out['r'][0].mean()), "finite": bool(torch.isfinite(out['r'][0]).all()),
                              "mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
            break
        except Exception:
            traceback.print_exc()
            print(json.dumps({"leg": "vp/CELEBA/indm_nll training", "freeze_flow": freeze, "error": traceback.format_exc()[-400:]}), flush=True)
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
