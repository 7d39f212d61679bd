#!/usr/bin/env python
"""Benchmark of record for the INDM hot path on B200.

Workload (BASELINE.json configs[1]): configs/vp/CIFAR10/indm_fid.py — DDPM++ (NCSN++ nres=4, nf=128) VP latent score
network + wolf flow — with the overrides sampling.method=pc, predictor=reverse_diffusion, corrector=none (SURVEY.md §0
fact 3), 1000 predictor steps, batch 1024 batch-sharded over 8 GPUs = 128 images per GPU (weak scaling: every rank
samples its own 128 images, no data-path collective).  Synthetic data, random-init weights (seeded).

One "step" = one complete `pc_sampler` call on the rank's batch: 1000 x (score-network forward + fused predictor update)
followed by the flow inverse.  `value` = images/s with the prior sample already resident in HBM; `e2e` = the same metric
through the public API `sampling.get_pc_sampler(...)(model, flow_model)` with the prior drawn on the host (pinned) and
copied in, and the samples copied back, every step.

Extra legs on the same JSON line (each says its own config; `--skip-extras` leaves them out): `ve_pc` = the north-star target
workload (ve/CIFAR10/indm: reverse diffusion + Langevin corrector, 2 NFE per step, FIR resampling, 1000 steps, 128 images per
GPU), `train` = the INDM-VP joint training step, `nll` = PF-ODE likelihood (sharded by image under torchrun), `celeba` = the
64x64 nres=8 configs (bounded slices), `ref_ops` = the reference's two native ops through the C-ABI.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                   CPU arm: the oracle port of the reference's PyTorch CPU path
    python bench.py --num-scales 4 --skip-train --skip-cpu   profiling slice of the same command for `ncu` (a line printed by such
                                                           a run says so in config.workload and is never a bench value)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pc_sampling_images_per_sec"
UNIT = "images/s"
PER_GPU_BATCH = 128
NUM_SCALES = 1000
WORKLOAD = ("vp/CIFAR10/indm_fid (DDPM++ nres=4 nf=128) + sampling.method=pc predictor=reverse_diffusion corrector=none, 1000 steps, "
            "128 images per GPU (1024 / 8 GPUs); step = one full pc_sampler call")
GFLOP_PER_IMAGE_FWD = 21.69      # score-net forward, CIFAR nres=4 (2*MAC, forward hooks on the reference)


def igemm_traffic():
    """DRAM bytes per igemm launch (dram__bytes_read.sum + dram__bytes_write.sum averaged over the launches of one forward) from the
    committed `ncu --set full` capture, if one was summarised into profiles/igemm_traffic.json; else None."""
    p = os.path.join(ROOT, "profiles", "igemm_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


def workload_config(device, num_scales=NUM_SCALES):
    from indm_b200 import configs
    cfg = configs.get_config("vp/CIFAR10/indm_fid")
    cfg.sampling.method = "pc"
    cfg.sampling.predictor = "reverse_diffusion"
    cfg.sampling.corrector = "none"
    cfg.sampling.num_scales = num_scales                  # loop length only; the SDE keeps N = model.num_scales = 1000
    cfg.device = device
    return cfg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(threads, samples=3):
    """cpu_baseline of the GPU arm's line: the `--impl reference` arm itself (same bounded samples), run as a child process with the
    GPU hidden (the reference wraps its model in nn.DataParallel, which grabs every visible GPU): `samples` timed samples after one
    warm-up sample + one flow inverse, ~15-30 s of host work."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(samples), "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not line:
        raise RuntimeError("cpu arm failed: " + (r.stderr or r.stdout)[-400:])
    d = json.loads(line[-1])
    return d["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["sample"]


def cpu_train_sample(threads, batch=8):
    """Bounded sample of the second metric on the host: ONE joint training step (flow_step_fn_nll, reference losses.py:258-320) of the
    oracle port at batch `batch`: wolf flow forward in training mode with the Neumann log-det series (autograd VJPs, last one with
    create_graph), score-network forward on the latent, DSM loss with importance-sampled t, prior log-p, and ONE backward through
    both networks.  The optimiser update (AdamW + EMA: ~0.1 % of the step on CPU) is left out; dropout is not part of the oracle."""
    import numpy as np
    import torch
    from indm_b200 import configs
    from oracle import ncsnpp as oncsnpp, sde as osde, flow as oflow
    torch.set_num_threads(threads)
    cfg = configs.get_config("vp/CIFAR10/indm_nll")
    cfg.device = torch.device("cpu")
    Ps = {k: v.requires_grad_(True) for k, v in oncsnpp.to_torch(oncsnpp.synth_params(cfg, 0)).items()}
    Pf = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in oflow.to_torch(oflow.synth_params(cfg, 1)).items()}
    sde = osde.get_sde(cfg)
    g = torch.Generator().manual_seed(0)
    rng = np.random.RandomState(0)
    x = torch.rand(batch, 3, 32, 32, generator=g) * 2 - 1
    layout = oflow.block_layout(cfg)
    shp = oflow.flow_input_shape(cfg)
    ns = rng.poisson(2.0, size=len(layout))
    varepss = []
    for (s_, b_, c_, first_) in layout:
        f = 2 ** s_
        varepss.append(torch.randn(batch, c_, shp[1] // f, shp[2] // f, generator=g))
    eps_post = torch.randn(batch, 64, generator=g)
    u = torch.rand(batch, generator=g)
    noise = torch.randn(batch, 3, 32, 32, generator=g)
    noise_T = torch.randn(batch, 3, 32, 32, generator=g)
    t0 = time.perf_counter()
    z, logdet_minus_kl, _, _ = oflow.wolf_train_forward(cfg, Pf, x, eps_post, ns, varepss)
    t, Z = sde.importance_time(u, cfg.training.truncation_time)
    mean, std = sde.marginal_prob(z, t)
    perturbed = mean + std[:, None, None, None] * noise
    score = oncsnpp.score_fn(cfg, sde, Ps, perturbed, t)
    losses_score = 0.5 * Z * ((score * std[:, None, None, None] + noise) ** 2).reshape(batch, -1).sum(1)
    meanT, stdT = sde.marginal_prob(z, torch.ones(batch))
    logp = sde.prior_logp(meanT + stdT[:, None, None, None] * noise_T)
    loss = torch.mean(losses_score - logdet_minus_kl - logp)
    loss.backward()
    dt = time.perf_counter() - t0
    assert sum(1 for v in Ps.values() if v.grad is None) <= 1, "score-network parameters without gradient"     # `sigmas` is a buffer
    n_flow = sum(1 for v in Pf.values() if v.requires_grad and v.grad is not None)
    no_grad = [k for k, v in Pf.items() if v.requires_grad and v.grad is None and k.rsplit('.', 1)[-1] in ('weight', 'bias')]
    assert n_flow > 300 and not no_grad, f"flow parameters without gradient: {no_grad[:4]}"      # the rest are buffers (running stats, scale, lamb ...)
    n_vjp = int(sum(int(n) + 3 for n in ns))
    return batch / dt, (f"1 joint step (flow fwd with {n_vjp} autograd VJPs + score fwd + one backward through both) of the oracle port at "
                        f"batch {batch}, FP32, {dt:.1f} s of CPU work")


REF_VENDORED = os.path.join(ROOT, "oracle", "_ref")
REF_SAMPLE_STEPS = 1           # PC steps per bounded sample at the full per-GPU batch (128): a few seconds of host work


def _reference_sampler(threads):
    """The reference's OWN pc_sampler on the host (oracle/_ref: the unmodified reference modules placed by oracle/make_ref.py),
    on this arm's config: returns (sample_fn() -> seconds for REF_SAMPLE_STEPS PC steps at batch 128, flow_fn() -> seconds for one
    flow inverse at batch 128, kind).  Falls back to the oracle port when oracle/_ref is absent."""
    import tempfile
    import torch
    torch.set_num_threads(threads)
    from oracle import ncsnpp as oncsnpp, flow as oflow
    if os.path.isdir(os.path.join(REF_VENDORED, "models")):
        os.environ["INDM_REFERENCE_ROOT"] = REF_VENDORED
        from oracle import ref_loader as rl
        mutils, sde_lib, sampling, fm, _ = rl.load("models.utils", "sde_lib", "sampling", "flow_models.flow_model", "models.ncsnpp")
        cfg = rl.get_config("configs/vp/CIFAR10/indm_fid.py")
        cfg.sampling.method, cfg.sampling.predictor, cfg.sampling.corrector = "pc", "reverse_diffusion", "none"
        cfg.sampling.num_scales = REF_SAMPLE_STEPS
        pcfg = workload_config(torch.device("cpu"))
        model = mutils.create_model(cfg)
        model.module.load_state_dict({k: torch.from_numpy(v) for k, v in oncsnpp.synth_params(pcfg, 0).items()})
        model.eval()
        with rl.reference_cwd():
            flow = fm.create_flow_model(cfg)
        flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in oflow.synth_params(pcfg, 1).items()})
        flow.eval()
        sde = sde_lib.get_sde(cfg)
        shape = (PER_GPU_BATCH, 3, 32, 32)
        cfg_id = rl.get_config("configs/vp/CIFAR10/indm_fid.py")
        cfg_id.sampling.method, cfg_id.sampling.predictor, cfg_id.sampling.corrector = "pc", "reverse_diffusion", "none"
        cfg_id.sampling.num_scales = REF_SAMPLE_STEPS
        cfg_id.flow.model = "identity"
        fn = sampling.get_sampling_fn(cfg_id, sde, shape, lambda v: (v + 1.) / 2., cfg.sampling.truncation_time)
        tmp = tempfile.mkdtemp()

        def sample_fn():
            t0 = time.perf_counter()
            fn(model, None, sample_dir=tmp, r=0)            # sampling.get_pc_sampler(...).pc_sampler: the reference's own loop
            return time.perf_counter() - t0

        def flow_fn():
            z = torch.randn(shape)
            t0 = time.perf_counter()
            with torch.no_grad():
                fm.flow_forward(cfg, flow, z, log_det=None, reverse=True)
            return time.perf_counter() - t0
        return sample_fn, flow_fn, "reference"
    # oracle port (the restatement pinned against the reference's outputs, tests/test_oracle_*.py)
    from oracle import sde as osde, sampler as osampler
    cfg = workload_config(torch.device("cpu"))
    params = oncsnpp.to_torch(oncsnpp.synth_params(cfg, 0))
    fparams = oflow.to_torch(oflow.synth_params(cfg, 1))
    sde = osde.get_sde(cfg)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(PER_GPU_BATCH, 3, 32, 32, generator=g)
    noises = [torch.randn(PER_GPU_BATCH, 3, 32, 32, generator=g) for _ in range(REF_SAMPLE_STEPS)]

    def sample_fn():
        t0 = time.perf_counter()
        with torch.no_grad():
            osampler.pc_sampler(sde, lambda a_, b_: oncsnpp.score_fn(cfg, sde, params, a_, b_), x, noises, REF_SAMPLE_STEPS, 1e-5, cfg.sampling.snr)
        return time.perf_counter() - t0

    def flow_fn():
        t0 = time.perf_counter()
        with torch.no_grad():
            oflow.wolf_reverse(cfg, fparams, x, torch.randn(PER_GPU_BATCH, 64, generator=g))
        return time.perf_counter() - t0
    return sample_fn, flow_fn, "port"


def run_reference(args):
    """CPU arm: the reference's own implementation of the path on the box's host cores, on this arm's config.  One step = one
    bounded sample: REF_SAMPLE_STEPS PC steps of `pc_sampler` at the full per-GPU batch (128 images); the flow inverse (one pass per
    1000 PC steps) is timed once during warm-up; images/s = 128 / (1000 * seconds per PC step + seconds per flow inverse)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = ""      # before torch initialises CUDA: the reference's nn.DataParallel would grab the GPUs
    threads = os.cpu_count() or 1
    sample_fn, flow_fn, kind = _reference_sampler(threads)
    t_flow = flow_fn()
    per_step = []
    for i in range(args.warmup + args.steps):
        dt = sample_fn() / REF_SAMPLE_STEPS
        if i >= args.warmup:
            per_step.append(dt)
    t_step = sum(per_step) / len(per_step)
    t_call = NUM_SCALES * t_step + t_flow
    value = PER_GPU_BATCH / t_call
    sample = (f"{REF_SAMPLE_STEPS} PC steps (1 NFE each) of the {'reference' if kind == 'reference' else 'oracle port of the reference'}'s pc_sampler at "
              f"batch {PER_GPU_BATCH}, FP32, {t_step * REF_SAMPLE_STEPS:.1f} s of CPU work per sample ({t_step:.2f} s per PC step) x {args.steps} timed samples; "
              f"flow inverse at batch {PER_GPU_BATCH} timed once: {t_flow:.1f} s; figure = 128 / (1000 x step + flow)")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": t_call * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "per_gpu_batch": PER_GPU_BATCH, "num_scales": NUM_SCALES,
                      "cpu_arm": "each step is a bounded sample of that workload on the host cores (see cpu_baseline.sample); ms_per_step is the "
                                 "extrapolated time of one full 1000-step pc_sampler call at batch 128, not the duration of a sample"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------- GPU arm
def _closure(op, cls_name=None, key=None):
    for c in (op.__closure__ or ()):
        v = c.cell_contents
        if cls_name is not None and v.__class__.__name__ == cls_name:
            return v
    if key is not None:
        names = op.__code__.co_freevars
        if key in names:
            return op.__closure__[names.index(key)].cell_contents
    return None


def _gn_apply_bytes(op):
    """Algorithmic HBM bytes of one `indm_gn_apply` launch: input read once (fp32 residual stream or bf16) + operand-dtype output
    (+ the raw operand copy when a skip conv follows) written once.  None for any other launch."""
    name = _closure(op, key="name")
    if name not in ("indm_gn_apply", "indm_gn_apply_dropout", "indm_gn_apply_pp"):
        return None
    a = _closure(op, key="cargs")
    val = lambda v: getattr(v, "value", v)
    Ca, Cb, in_dt, N, H, W = val(a[1]), (val(a[3]) if val(a[2]) else 0), val(a[4]), val(a[5]), val(a[6]), val(a[7])
    dropout = name == "indm_gn_apply_dropout"
    resample = 0 if name != "indm_gn_apply" else val(a[14])
    raw = None if dropout else val(a[15] if name == "indm_gn_apply_pp" else a[16])      # padded-pixel borders are never written
    from indm_b200 import _lib as L
    C = Ca + Cb
    Po = H * W * (4 if resample == 1 else 1) // (4 if resample == 2 else 1)
    return N * H * W * C * (4 if in_dt == L.DTYPE_F32 else 2) + N * Po * C * 2 * (2 if raw else 1)


def _fir_bytes(op):
    """Algorithmic HBM bytes of one `indm_fir_nhwc` launch (input read once + output written once); None for other launches."""
    if _closure(op, key="name") != "indm_fir_nhwc":
        return None
    from indm_b200 import _lib as L
    a = _closure(op, key="cargs")
    val = lambda v: getattr(v, "value", v)
    din, dout, N, H, W, C, mode = val(a[2]), val(a[3]), val(a[4]), val(a[5]), val(a[6]), val(a[7]), val(a[9])
    Ho, Wo = {1: (2 * H, 2 * W), 2: (H // 2, W // 2), 3: (H + 1, W + 1), 4: (H - 1, W - 1)}[mode]
    esz = lambda d: 2 if d == L.DTYPE_BF16 else 4
    return N * H * W * C * esz(din) + N * Ho * Wo * C * esz(dout)


def class_stream_ms(ops, reps=3):
    """average device time of the launches in `ops` issued back to back between two CUDA events (GPU parked in a spin while the host
    enqueues)"""
    import torch
    tot = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(60_000_000)
        e0.record()
        for op in ops:
            op()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def fir_roofline(eng, pk, pk_kind):
    """roofline_hbm of the FIR resampling class (`fir_nhwc_kernel`) of one VE score-network forward"""
    import torch
    for _ in range(2):
        eng.launch()
    torch.cuda.synchronize()
    ops, by = [], 0.0
    for op in eng.ops:
        try:
            b = _fir_bytes(op)
        except Exception:
            b = None
        if b is not None:
            ops.append(op)
            by += b
    if not ops:
        return None
    ms, fwd = class_stream_ms(ops), class_stream_ms(eng.ops)
    eng.launch()
    torch.cuda.synchronize()
    gbs = by / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": None,
            "kernel": "fir_nhwc_kernel (upfirdn2d call sites of models/up_or_down_sampling.py: up x2, down x2, pad (2,2))",
            "peak_source": pk_kind + " hbm_gbs", "launches_per_forward": len(ops), "avg_launch_ms": ms / len(ops),
            "share_of_forward_device_time": ms / fwd, "forward_device_ms": fwd,
            "algorithmic_bytes": "input read once + output written once, operand dtypes as launched (SURVEY 8d: 2 x (in + out) bytes counts r+w)"}


def kernel_rooflines(net, eng, reps=3):
    """Device time of the two kernel classes of one score-network forward, CUDA events on the launching stream, GPU first parked in
    a ~30 ms spin (`torch.cuda._sleep`) while the host enqueues, so host launch gaps are not measured.  Two views:
      * class stream: ONLY the launches of one class (all 115 igemm launches of the forward, in order, on their real buffers), back
        to back between two events -> average launch duration = elapsed / launches.  Launches pipeline as they do inside the
        sampler's CUDA graph; operands are NOT freshly written by a producer (5 GB of activations cycle through the 126 MB L2), so
        this is cold-operand, i.e. conservative.  This is the figure `achieved` uses.
      * evented: an event after every launch of the whole forward; each interval also contains the event-record / launch
        serialisation (~4-5 us per launch, visible against ncu's per-kernel durations) -> reported as `*_evented` only."""
    import torch
    for _ in range(2):
        eng.launch()
    torch.cuda.synchronize()
    ig_ops, gn_ops, ig_fl, gn_by = [], [], 0.0, 0.0
    for op in eng.ops:
        d = _closure(op, cls_name="IgemmDesc")
        if d is not None:
            ig_ops.append(op)
            ig_fl += 2.0 * d.N * d.H * d.W * d.Cout * (d.Cin * d.taps + (d.Cin2 if d.a2 else 0))
            continue
        try:
            b = _gn_apply_bytes(op)
        except Exception:
            b = None
        if b is not None:
            gn_ops.append(op)
            gn_by += b

    def stream_ms(ops):
        tot = 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(60_000_000)
            e0.record()
            for op in ops:
                op()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps

    fwd_ms = stream_ms(eng.ops)
    ig_ms = stream_ms(ig_ops)
    gn_ms = stream_ms(gn_ops) if gn_ops else None
    # evented view
    ev_ig = ev_gn = ev_all = 0.0
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(eng.ops) + 1)]
        torch.cuda._sleep(60_000_000)
        evs[0].record()
        for i, op in enumerate(eng.ops):
            op()
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i, op in enumerate(eng.ops):
            dt = evs[i].elapsed_time(evs[i + 1])
            ev_all += dt
            if op in ig_ops:
                ev_ig += dt
            elif op in gn_ops:
                ev_gn += dt
    eng.launch()                       # leave the buffers in a consistent state
    torch.cuda.synchronize()
    return dict(fwd_ms=fwd_ms, n_ig=len(ig_ops), ig_ms=ig_ms, ig_tflops=ig_fl / (ig_ms * 1e-3) / 1e12,
                ig_tflops_evented=ig_fl / (ev_ig / reps * 1e-3) / 1e12, ig_share_evented=ev_ig / ev_all,
                n_gn=len(gn_ops), gn_ms=gn_ms, gn_gbs=(gn_by / (gn_ms * 1e-3) / 1e9 if gn_ops else None),
                gn_gbs_evented=(gn_by / (ev_gn / reps * 1e-3) / 1e9 if gn_ops else None), gn_share_evented=ev_gn / ev_all,
                fwd_ms_evented=ev_all / reps)


def train_throughput(cfg_s, model, flow, sde, dev, world, timed, steps=10, warmup=16, both_modes=True):
    """Second BASELINE metric (configs[2]): training samples/s of the INDM-VP joint step `flow_step_fn_nll` (losses.py:258-320),
    128 images per GPU, data parallel (one NCCL all-reduce per flat gradient buffer: score network and flow).  One step = wolf flow
    forward in training mode (batch-statistics BatchNorm encoder, posterior sample, prior-flow KL, Neumann log-det series of all
    32 iResBlocks), score-network train-mode forward (dropout 0.1) + DSM loss + prior log-p, ONE backward through both networks
    (score dgrad + wgrad, flow first-order + second-order Neumann gradient, encoder / prior backward), then global-norm clip +
    AdamW + EMA on both.  Warm-up is 16 steps: the flow's per-(block, series length) log-det chains are captured as CUDA graphs on
    their second use, and the Poisson series lengths need a few steps to have all been seen (steady state of a training run)."""
    import torch
    from indm_b200 import configs, losses
    from indm_b200.models.ema import ExponentialMovingAverage
    cfg = configs.get_config("vp/CIFAR10/indm_nll")
    cfg.device = dev
    opt = losses.get_optimizer(cfg, model.parameters())
    state = dict(optimizer=opt, model=model, ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
    fopt = losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr)
    flow_state = dict(optimizer=fopt, model=flow, ema=ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate), step=0)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    batch_host = (torch.rand(PER_GPU_BATCH, 3, 32, 32) * 2 - 1).pin_memory()

    def one():
        b = batch_host.to(dev, non_blocking=True)             # host batch in, four per-sample loss vectors out (like run_lib.train)
        return step_fn(state, flow_state, b)

    ms, launches = timed(one, steps, warmup)
    res = {"metric": "train_samples_per_sec", "value": world * PER_GPU_BATCH / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
           "steps": steps, "warmup": warmup, "per_gpu_batch": PER_GPU_BATCH, "gpu_launches": launches,
           "config": "vp/CIFAR10/indm_nll flow_step_fn_nll: JOINT step, wolf flow (16+16 iResBlocks, idim 512, training-mode encoder) and "
                     "score network (DDPM++ nres=4, dropout 0.1) both trained: fwd + bwd + clip + AdamW + EMA on both; host batch in "
                     "(pinned, H2D inside the timed region), four per-sample loss vectors read back every step (= the e2e figure)",
           "dtype": "bf16 (score net, iResBlocks) / 3xTF32 (posterior encoder + KL)",
           "e2e": {"value": world * PER_GPU_BATCH / (ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": batch_host.numel() * 4,
                   "d2h_bytes_per_step": 4 * PER_GPU_BATCH * 4}}
    # algorithmic FLOPs of the step from the ACTUAL series lengths of the last step (they are Poisson draws): per block
    # g = conv3x3(c->idim) + 1x1(idim->idim) + conv3x3(idim->c); forward = 1 g + (n + 3) VJPs (each one g); backward (wolf_backward.py)
    # = recompute (w1 + w2) + reverse chain (g) + forward mode (w1 + w2) + gradient chain (g) + 6 weight-gradient GEMMs (2 g);
    # score network = forward + dgrad + wgrad = 3 x 21.69 GFLOP per image.
    try:
        core = flow.module
        eng = core.engine(PER_GPU_BATCH, leg="training")
        idim = eng.idim
        _, h0, w0 = core.input_shape
        fl = 0.0
        for (s_, b_, m_), nv in zip(eng.blocks, eng.vjp_per_block):
            P = PER_GPU_BATCH * (h0 >> s_) * (w0 >> s_)
            w1 = 2.0 * P * 9 * m_.channels * idim
            w2 = 2.0 * P * idim * idim
            g_ = 2 * w1 + w2
            fl += g_ * (1 + nv) + (4 * g_ + 2 * (w1 + w2))
        fl_score = 3 * GFLOP_PER_IMAGE_FWD * 1e9 * PER_GPU_BATCH
        pk, pk_kind = peaks()
        tf = (fl + fl_score) / (ms * 1e-3) / 1e12
        res["roofline"] = {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                           "frac": tf / pk["bf16_tflops_sustained"], "traffic": None, "peak_source": pk_kind + " bf16_tflops_sustained",
                           "algorithmic_tflop_per_step": (fl + fl_score) / 1e12, "flow_tflop": fl / 1e12, "score_tflop": fl_score / 1e12,
                           "flow_vjp_chains_last_step": int(sum(eng.vjp_per_block)),
                           "note": "whole-step figure (all kernels + host gaps) against the dense BF16 peak; FLOPs counted from the actual "
                                   "Poisson series lengths of the last timed step"}
    except Exception as e:
        res["roofline"] = {"error": repr(e)[:200]}
    if both_modes:
        # the same step with the iResBlocks in compensated TF32 (precision.set_policy('flow', 'training', 'tf32')): the mode in which
        # the block log-det alone also meets 1e-3 relative; reported beside the default, fewer steps (it is ~5x slower)
        from indm_b200 import precision
        try:
            precision.set_policy("flow", "training", "tf32")
            ms2, _ = timed(one, 2, 3)
            res["tf32_blocks"] = {"ms_per_step": ms2, "value": world * PER_GPU_BATCH / (ms2 * 1e-3), "unit": "samples/s", "steps": 2, "warmup": 3,
                                  "dtype": "bf16 (score net) / 3xTF32 (whole flow)"}
        except Exception as e:
            res["tf32_blocks"] = {"error": repr(e)[:200]}
        finally:
            precision.set_policy("flow", "training", "bf16")
    return res


def _randomise_zero_init(net):
    """random-init weights of that architecture; the ~0-initialised tensors (init_scale=0) get ordinary magnitudes so the network
    output is not identically ~0"""
    import torch
    with torch.no_grad():
        for n_, p_ in net.named_parameters():
            if p_.dim() > 1 and float(p_.abs().max()) < 1e-6:
                fan = p_[0].numel() + p_.shape[0] * (p_[0, 0].numel() if p_.dim() > 2 else 1)
                p_.uniform_(-1, 1).mul_((6.0 / fan) ** 0.5)


def ve_pc_leg(dev, world, timed, steps=2, warmup=1, num_scales=NUM_SCALES, global_norms=False):
    """The north-star TARGET workload: configs/ve/CIFAR10/indm.py — NCSN++ (nres=4, FIR resampling, input pyramid, Fourier
    embedding) + wolf flow, PC sampling with the reverse-diffusion predictor AND the Langevin corrector (2 NFE per step, per-sample
    norms, sampling.py:263-292,410-456), 1000 steps, 128 images per GPU.  One step = one complete pc_sampler call."""
    import torch
    from indm_b200 import configs, sde_lib, sampling
    from indm_b200.models import utils as mutils
    from indm_b200.flow_models import flow_model as fm
    cfg = configs.get_config("ve/CIFAR10/indm")
    cfg.sampling.method, cfg.sampling.predictor, cfg.sampling.corrector = "pc", "reverse_diffusion", "langevin"
    cfg.sampling.num_scales = num_scales
    cfg.sampling.global_langevin_norms = bool(global_norms)
    cfg.device = dev
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    _randomise_zero_init(model.module)
    flow = fm.create_flow_model(cfg)
    flow.eval()
    sde = sde_lib.get_sde(cfg)
    shape = (PER_GPU_BATCH, 3, 32, 32)
    fn = sampling.get_sampling_fn(cfg, sde, shape, lambda v: v, cfg.sampling.truncation_time)
    prior_host = (torch.randn(shape) * cfg.model.sigma_max).pin_memory()
    out = {}

    def call():
        out["r"] = fn(model, flow, prior=prior_host.to(dev, non_blocking=True), seed=1)
        out["host"] = out["r"][1].cpu()

    ms, launches = timed(call, steps, warmup)
    pk, pk_kind = peaks()
    eng = model.module.engine(PER_GPU_BATCH, infer=True)      # the plan the sampler runs
    res = {"metric": METRIC, "value": world * PER_GPU_BATCH / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "ms_per_pc_step": ms / num_scales,
           "steps": steps, "warmup": warmup, "per_gpu_batch": PER_GPU_BATCH, "num_scales": num_scales, "nfe_per_call": 2 * num_scales,
           "gpu_launches": launches, "finite": bool(torch.isfinite(out["host"]).all()), "dtype": "bf16 (score net) / 3xTF32 (flow inverse)",
           "config": "ve/CIFAR10/indm (NCSN++ nres=4 nf=128, fir=True, progressive_input=residual, fourier embedding) + wolf flow; "
                     "sampling.method=pc predictor=reverse_diffusion corrector=langevin snr=0.16, 1000 steps x 2 NFE, 128 images per GPU; "
                     "end to end through sampling.get_sampling_fn (host prior in, samples out)",
           "langevin_statistics": ("global batch means: 3 floats all-reduced (NCCL, inside the CUDA graph) per corrector step"
                                   if (global_norms and world > 1) else "per-rank batch means (no communication)" if world > 1 else "single batch"),
           "score_forward_tflops_algorithmic": 21.75 * PER_GPU_BATCH * 2 * num_scales / (ms * 1e-3) / 1e3}
    try:
        fr = fir_roofline(eng, pk, pk_kind)
        if fr is not None:
            res["roofline_hbm"] = fr
    except Exception as e:
        res["roofline_hbm"] = {"error": repr(e)[:200]}
    return res


def nll_leg(dev, world, timed):
    """PF-ODE likelihood (likelihood.get_likelihood_fn, BASELINE configs[2] second half) on vp/CIFAR10/indm_nll, 128 images per GPU,
    sharded by image (per-rank solvers on their sub-batches; no data-path collective), device-resident RK45, in the
    DEFAULT precision policy (score forward + Hutchinson VJP and flow log-det in compensated TF32: the mode that meets 0.01 bpd)."""
    import numpy as np
    import torch
    from indm_b200 import configs, sde_lib, likelihood
    cfg = configs.get_config("vp/CIFAR10/indm_nll")
    cfg.device = dev
    sde = sde_lib.get_sde(cfg)
    from indm_b200.models import utils as mutils
    from indm_b200.flow_models import flow_model as fm
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    flow = fm.create_flow_model(cfg)
    model.eval()
    flow.eval()
    g = torch.Generator().manual_seed(1 + int(os.environ.get("RANK", "0")))
    data = (torch.rand(PER_GPU_BATCH, 3, 32, 32, generator=g) * 2 - 1).pin_memory()
    fn = likelihood.get_likelihood_fn(cfg, sde, lambda v: (v + 1.) / 2., method="RK45-device")
    out = {}

    def call():
        np.random.seed(7)
        out["r"] = fn(model, flow, data.to(dev, non_blocking=True))
        out["bpd"] = out["r"][0].cpu()

    ms, launches = timed(call, 1, 1)
    nfe = int(out["r"][2])
    bpd_policy = out["bpd"].clone()
    # the same call with BF16 forced everywhere (round 1's timed mode: 0.03 - 0.05 bpd from the reference), side by side
    bf16 = None
    try:
        model.module.compute_mode = flow.module.compute_mode = "bf16"
        ms_b, _ = timed(call, 1, 1)
        bf16 = {"value": world * PER_GPU_BATCH / (ms_b * 1e-3), "unit": "images/s", "ms_per_call": ms_b, "nfe": int(out["r"][2]),
                "ms_per_nfe": ms_b / max(int(out["r"][2]), 1), "max_abs_bpd_diff_vs_policy": float((out["bpd"] - bpd_policy).abs().max())}
    except Exception as e:
        bf16 = {"error": repr(e)[:200]}
    finally:
        model.module.compute_mode = flow.module.compute_mode = "auto"
        out["bpd"] = bpd_policy
    return {"metric": "nll_images_per_sec", "bf16_everywhere": bf16, "value": world * PER_GPU_BATCH / (ms * 1e-3), "unit": "images/s", "ms_per_call": ms, "nfe": nfe,
            "ms_per_nfe": ms / max(nfe, 1), "per_gpu_batch": PER_GPU_BATCH, "gpu_launches": launches, "bpd_mean": float(out["bpd"].mean()),
            "finite": bool(torch.isfinite(out["bpd"]).all()), "dtype": "3xTF32 (score forward + input-VJP, flow log-det): the precision pinned at 0.01 bpd",
            "config": "vp/CIFAR10/indm_nll likelihood_fn(method='RK45-device', rtol=atol=1e-5), wolf flow forward + (20+n)-term log-det inside the "
                      "call, 128 images per GPU, sharded by image; weights = the modules' initialisers (head conv ~0: smooth ODE, few NFE) — "
                      "ms_per_nfe (1 forward + 1 input-VJP) is the transferable figure"}


def celeba_leg(dev, timed, batch=64, pc_steps=40):
    """BASELINE configs[3] / [4] (3x64x64, nres = 8, wolf flow with flow.squeeze) on ONE GPU as bounded slices: `pc_steps` steps of
    the 1000-step VE PC + Langevin schedule (every step replays the same CUDA graph, so ms per PC step is the 1000-step figure
    / 1000; the flow inverse is timed inside the slice and reported amortised over 1000), and the VP joint training step."""
    import torch
    from indm_b200 import configs, sde_lib, sampling, losses
    from indm_b200.models import utils as mutils
    from indm_b200.models.ema import ExponentialMovingAverage
    from indm_b200.flow_models import flow_model as fm
    res = {}
    torch.manual_seed(0)
    cfg = configs.get_config("ve/CELEBA/indm")
    cfg.model.num_res_blocks = 8
    cfg.device = dev
    model = mutils.create_model(cfg)
    _randomise_zero_init(model.module)
    flow = fm.create_flow_model(cfg)
    flow.eval()
    sde = sde_lib.get_sde(cfg)
    # the PC steps and the flow inverse are timed separately (the inverse's fixed-point iteration count depends on the sample, so a
    # difference of two slice lengths does not isolate it): sampler with flow.model = 'identity', then flow_forward(reverse=True)
    cfg.sampling.num_scales = pc_steps
    flow_kind = cfg.flow.model
    cfg.flow.model = "identity"
    fn = sampling.get_sampling_fn(cfg, sde, (batch, 3, 64, 64), lambda v: v, cfg.sampling.truncation_time)
    out = {}

    def call():
        out["r"] = fn(model, None, seed=1)
    ms_slice, _ = timed(call, 2, 1)
    per_step = ms_slice / pc_steps
    cfg.flow.model = flow_kind
    zlat = out["r"][0]
    flow_ms, _ = timed(lambda: fm.flow_forward(cfg, flow, zlat, log_det=None, reverse=True), 3, 1)
    res["ve_pc"] = {"config": "ve/CELEBA/indm num_res_blocks=8, 64x64, PC reverse_diffusion + langevin (2 NFE per step), wolf flow (squeeze) inverse",
                    "batch": batch, "slice_pc_steps": pc_steps, "ms_per_pc_step": per_step, "flow_inverse_ms": flow_ms,
                    "images_per_sec_at_1000_steps": batch / ((per_step * 1000 + flow_ms) * 1e-3),
                    "score_forward_tflops_algorithmic": 142.9 * batch * 2 / (per_step * 1e-3) / 1e3,
                    "note": "bounded slice of the 1000-step schedule on one GPU (every step replays the same CUDA graph), flow inverse timed "
                            "separately on the slice's output; the 1000-step figure is extrapolated and labelled as such"}
    del model, flow, fn
    torch.cuda.empty_cache()
    cfg = configs.get_config("vp/CELEBA/indm_nll")
    cfg.model.num_res_blocks = 8
    cfg.device = dev
    model = mutils.create_model(cfg)
    flow = fm.create_flow_model(cfg)
    sde = sde_lib.get_sde(cfg)
    state = dict(optimizer=losses.get_optimizer(cfg, model.parameters()), model=model,
                 ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
    flow_state = dict(optimizer=losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr), model=flow,
                      ema=ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate), step=0)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    bh = (torch.rand(batch, 3, 64, 64) * 2 - 1).pin_memory()
    ms_t, _ = timed(lambda: step_fn(state, flow_state, bh.to(dev, non_blocking=True)), 4, 10)
    res["train"] = {"config": "vp/CELEBA/indm_nll num_res_blocks=8, 64x64, flow_step_fn_nll JOINT step (flow + score)", "batch": batch,
                    "ms_per_step": ms_t, "samples_per_sec": batch / (ms_t * 1e-3),
                    "score_fwd_bwd_tflops_algorithmic": 3 * 142.9 * batch / (ms_t * 1e-3) / 1e3}
    del model, flow, state, flow_state, step_fn
    torch.cuda.empty_cache()
    return res


def ref_ops_leg(dev):
    """The reference's two native ops through the C-ABI (`indm_upfirdn2d_f32`, `indm_bias_act_f32`: what op/upfirdn2d.cpp:12-19 and
    op/fused_bias_act.cpp:11-17 bind) on SURVEY section 7's gate shape 16 x 256 x 32 x 32 FP32: GB/s of algorithmic bytes (input read
    once + output written once) against the measured HBM copy bandwidth.  Inputs are cycled through 12 distinct buffers (> L2)."""
    import torch
    from indm_b200 import _lib as L
    pk, pk_kind = peaks()
    N, C, S = 16, 256, 32
    k = torch.tensor([1., 3., 3., 1.])
    k2 = (k[:, None] * k[None, :])
    k2 = (k2 / k2.sum()).to(dev)
    res = {}
    cases = {"up2 (pad 2,1)": (2, 1, 2, 1, 4.0), "down2 (pad 1,1)": (1, 2, 1, 1, 1.0), "pad (2,2)": (1, 1, 2, 2, 1.0)}
    nbuf = 12
    xs = [torch.randn(N * C, S, S, 1, device=dev) for _ in range(nbuf)]
    for name, (up, down, p0, p1, gain) in cases.items():
        Ho = (S * up + p0 + p1 - 4) // down + 1
        ys = [torch.empty(N * C, Ho, Ho, 1, device=dev) for _ in range(nbuf)]
        kk = (k2 * gain).contiguous()

        def run(i):
            L.call("indm_upfirdn2d_f32", L.ptr(xs[i % nbuf]), L.ptr(kk), L.ptr(ys[i % nbuf]), N * C, S, S, 4, 4, up, up, down, down, p0, p1, p0, p1)
        for i in range(nbuf):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 4 * nbuf
        e0.record()
        for i in range(reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        by = 4 * N * C * (S * S + Ho * Ho)
        res["upfirdn2d " + name] = {"us": us, "GB/s": by / (us * 1e-6) / 1e9, "frac_of_hbm_peak": by / (us * 1e-6) / 1e9 / pk["hbm_gbs"], "bytes": by}
    xb = [torch.randn(N, C, S, S, device=dev) for _ in range(nbuf)]
    yb = [torch.empty(N, C, S, S, device=dev) for _ in range(nbuf)]
    bias = torch.randn(C, device=dev)

    def runb(i):
        L.call("indm_bias_act_f32", L.ptr(xb[i % nbuf]), L.ptr(bias), None, L.ptr(yb[i % nbuf]), xb[0].numel(), C, S * S, 3, 0,
               0.2, 2.0 ** 0.5)
    try:
        for i in range(nbuf):
            runb(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 4 * nbuf
        e0.record()
        for i in range(reps):
            runb(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        by = 8 * xb[0].numel()
        res["fused_leaky_relu (bias_act)"] = {"us": us, "GB/s": by / (us * 1e-6) / 1e9, "frac_of_hbm_peak": by / (us * 1e-6) / 1e9 / pk["hbm_gbs"], "bytes": by}
    except Exception as e:
        res["fused_leaky_relu (bias_act)"] = {"error": repr(e)[:200]}
    res["shape"] = f"{N} x {C} x {S} x {S} fp32, 12 rotating buffers (> L2), CUDA events over 48 launches"
    res["peak_source"] = pk_kind + " hbm_gbs"
    return res



def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from indm_b200 import sde_lib, sampling, _lib as L
    from indm_b200.models import utils as mutils
    from indm_b200.flow_models import flow_model as fm

    cfg = workload_config(dev, args.num_scales)
    flow_note = "wolf"
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    net = model.module
    _randomise_zero_init(net)
    flow = fm.create_flow_model(cfg)
    flow.eval()
    flow_note = "wolf (prior-flow sample of h + fixed-point inverse of 16+16 iResBlocks, idim 512), inside every timed step"
    sde = sde_lib.get_sde(cfg)
    shape = (PER_GPU_BATCH, 3, 32, 32)
    sampler = sampling.get_pc_sampler(cfg, sde, shape, sampling.ReverseDiffusionPredictor, sampling.NoneCorrector, lambda v: (v + 1.) / 2.,
                                      cfg.sampling.snr, n_steps=1, probability_flow=False, continuous=True, denoise=True,
                                      eps=cfg.sampling.truncation_time, device=dev)
    prior_dev = torch.randn(shape, device=dev)
    prior_host = torch.randn(shape).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.launches
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, (L.launches - l0)

    def step_resident():
        sampler(model, flow, prior=prior_dev, seed=1)

    def step_e2e():
        x0 = prior_host.to(dev, non_blocking=True)
        b, a, _ = sampler(model, flow, prior=x0, seed=1)
        return a.cpu()

    if args.train_only:
        tr = train_throughput(cfg, model, flow, sde, dev, world, timed, both_modes=(args.train_both == 1))
        if rank == 0:
            print(json.dumps(tr), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_step, launches = timed(step_resident, args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None
    e2e_steps = max(2, min(args.steps, 4))
    ms_e2e, _ = timed(step_e2e, e2e_steps, 1)
    train = None if args.skip_train else train_throughput(cfg, model, flow, sde, dev, world, timed, both_modes=(world == 1 if args.train_both < 0 else args.train_both == 1))
    extras = {}
    if not args.skip_extras:
        import traceback

        def leg(name, fn):
            try:
                extras[name] = fn()
            except Exception as e:          # an extra leg never takes the bench line down; the failure is on the line
                extras[name] = {"error": repr(e)[:300], "trace": traceback.format_exc()[-600:]}
            torch.cuda.empty_cache()
        leg("ve_pc", lambda: ve_pc_leg(dev, world, timed, num_scales=args.num_scales, global_norms=args.global_langevin_norms))
        leg("nll", lambda: nll_leg(dev, world, timed))
        if world == 1:
            leg("celeba", lambda: celeba_leg(dev, timed))
            leg("ref_ops", lambda: ref_ops_leg(dev))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * PER_GPU_BATCH / (ms_step * 1e-3)
    e2e = world * PER_GPU_BATCH / (ms_e2e * 1e-3)
    pk, pk_kind = peaks()
    traffic = igemm_traffic()
    eng = net.engine(PER_GPU_BATCH, infer=True)               # the plan the sampler runs
    kr = kernel_rooflines(net, eng)
    tf = kr["ig_tflops"]
    peak_tf = pk["bf16_tflops_sustained"]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD
                               + ("" if args.num_scales == NUM_SCALES else f" -- PROFILING SLICE with {args.num_scales} PC steps, not a bench value"),
                   "flow": flow_note, "per_gpu_batch": PER_GPU_BATCH, "num_scales": args.num_scales,
                   "l2_policy": "inputs larger than L2 (activations ~5 GB per forward), no flush",
                   "noise": "in-kernel Philox4x32-10 (no noise tensor in HBM)"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": prior_host.numel() * 4, "d2h_bytes_per_step": prior_host.numel() * 4},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                     "kernel": "igemm_kernel (tcgen05 implicit-GEMM conv/GEMM)", "peak_source": pk_kind + " bf16_tflops_sustained",
                     "launches_per_forward": kr["n_ig"], "avg_launch_ms": kr["ig_ms"] / kr["n_ig"],
                     "share_of_forward_device_time": kr["ig_ms"] / kr["fwd_ms"], "forward_device_ms": kr["fwd_ms"],
                     "algorithmic_gflop_per_image_forward": GFLOP_PER_IMAGE_FWD,
                     "timing": "two CUDA events around the forward's 115 igemm launches issued back to back on their real buffers "
                               "(cold operands), GPU parked in a spin while the host enqueues; forward_device_ms = the same for all launches",
                     "achieved_evented": kr["ig_tflops_evented"], "share_evented": kr["ig_share_evented"],
                     "evented_note": "event after every launch of the whole forward: intervals include ~4-5 us of event/launch serialisation"},
    }
    if kr["gn_gbs"] is not None:
        # second kernel class of the step (norm / elementwise kernels are HBM bound)
        out["roofline_hbm"] = {"bound": "hbm", "achieved": kr["gn_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": kr["gn_gbs"] / pk["hbm_gbs"],
                               "traffic": (traffic or {}).get("gn_apply_dram_bytes_per_launch"),
                               "kernel": "gn_apply_kernel (GroupNorm + SiLU [+ resample] [+ raw operand copy])",
                               "peak_source": pk_kind + " hbm_gbs", "launches_per_forward": kr["n_gn"], "avg_launch_ms": kr["gn_ms"] / kr["n_gn"],
                               "share_of_forward_device_time": kr["gn_ms"] / kr["fwd_ms"],
                               "achieved_evented": kr["gn_gbs_evented"], "share_evented": kr["gn_share_evented"],
                               "algorithmic_bytes": "input read once (4 B fp32 residual stream / 2 B bf16) + 2 B output per element (+ 2 B raw copy)"}
    if train is not None:
        out["train"] = train
    out["e2e"]["steps"] = e2e_steps
    out.update(extras)
    from indm_b200 import precision
    out["config"]["precision_policy"] = precision.POLICY
    if world == 1 and not args.skip_cpu:
        try:
            v, kind, sample = cpu_sample(os.cpu_count() or 1)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind, "sample": sample}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "error": repr(e)[:300]}
        if train is not None:
            try:
                tv, tsample = cpu_train_sample(os.cpu_count() or 1)
                out["train"]["cpu_baseline"] = {"value": tv, "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": tsample}
            except Exception as e:             # a reported baseline only: never let it take the bench line down
                out["train"]["cpu_baseline"] = {"value": None, "error": repr(e)[:200]}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-scales", type=int, default=NUM_SCALES, help="PC steps per sampler call (profiling slices only; the bench is 1000)")
    ap.add_argument("--skip-train", action="store_true", help="profiling: leave out the training leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling: leave out the CPU baseline sample")
    ap.add_argument("--skip-extras", action="store_true", help="leave out the extra legs (ve_pc, nll, celeba, ref_ops)")
    ap.add_argument("--train-only", action="store_true", help="development: only the training leg (prints its JSON object)")
    ap.add_argument("--train-both", type=int, default=-1, help="development: 1 / 0 forces the TF32-blocks side measurement of the training leg on / off")
    ap.add_argument("--global-langevin-norms", action="store_true",
                    help="ve_pc leg under torchrun: all-reduce the Langevin norm statistics (global batch means) instead of per-rank means")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
