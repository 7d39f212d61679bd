"""Bits-per-dimension evaluation loop of the reference (evaluation.py:388-495): NELBO (`eval.num_nelbo` passes over the evaluation
set), the NLL without the residual term ("NLL WRONG", skipped by `eval.skip_nll_wrong`), the NLL with it at eps = 1e-5 (or
`eval.truncation_time`), and — when `training.truncation_time != 1e-5` — at the training truncation time.  Every batch is
uniformly dequantised and scaled exactly as in the reference (evaluation.py:404-406).  FID / IS evaluation is out of scope."""
import logging

import numpy as np
import torch

from . import datasets, parallel


def _stats(values):
    """(mean, std, count) of a list of per-sample figures over ALL ranks: under torch.distributed every rank holds the figures of its
    own slice of each batch (evaluation shards by image, no data-path collective); three float64 scalars are summed."""
    v = np.asarray(values, dtype=np.float64)
    acc = torch.tensor([v.sum(), (v * v).sum(), float(v.size)], dtype=torch.float64)
    rank, ws = parallel.world()
    if ws > 1:
        import torch.distributed as dist
        dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
        acc = acc.to(dev)
        dist.all_reduce(acc)
        acc = acc.cpu()
    n = max(float(acc[2]), 1.0)
    mean = float(acc[0]) / n
    return mean, float(max(float(acc[1]) / n - mean * mean, 0.0) ** 0.5), int(acc[2])


def _passes(config, eval_ds, scaler, num_data, fn):
    """One pass over ceil(num_data / eval.batch_size) batches; `fn(batch) -> tuple of [B] tensors`; returns lists of floats (this
    rank's slice of every batch when torch.distributed is initialised: every rank iterates the same `eval_ds`)."""
    outs = None
    it = iter(eval_ds)
    for _ in range((num_data - 1) // config.eval.batch_size + 1):
        batch, it = datasets.get_batch(config, it, eval_ds)
        batch = parallel.shard_batch(batch)
        batch = scaler(datasets.dequantize(batch))
        vals = fn(batch)
        if outs is None:
            outs = [[] for _ in vals]
        for o, v in zip(outs, vals):
            o.extend(v.detach().cpu().numpy().reshape(-1))
    return outs


def get_bpd(config, eval_ds, scaler, nelbo_fn, nll_fn, score_model, flow_model=None, step=0, eval=False):
    """evaluation.py:388-495.  Logs what the reference logs and additionally RETURNS the figures:
    {'nelbo', 'nelbo_residual', 'nll_wrong' (or None), 'nll', 'nll_train_eps' (or None)} (means over the evaluated samples)."""
    res = dict(nelbo=None, nelbo_residual=None, nll_wrong=None, nll=None, nll_train_eps=None)
    with torch.no_grad():
        if config.flow.model != 'identity':
            flow_model.eval()
        num_data = config.eval.num_test_data if eval else 10000
        full, full_res = [], []
        for _ in range(config.eval.num_nelbo):
            a, b = _passes(config, eval_ds, scaler, num_data, lambda x: nelbo_fn(score_model, flow_model, x, None))
            (ma, sa, na), (mb, sb, nb) = _stats(a), _stats(b)
            full.append(ma)
            full_res.append(mb)
            logging.info("step: %d, num samples: %d, mean nelbo bpd: %.5e, std nelbo bpd: %.5e" % (step, na, ma, sa))
            logging.info("step: %d, num samples: %d, mean nelbo_residual bpd: %.5e, std nelbo_residual bpd: %.5e" % (step, nb, mb, sb))
        if full:
            res['nelbo'], res['nelbo_residual'] = float(np.mean(full)), float(np.mean(full_res))
            logging.info("step: %d, average nelbo bpd out of %d evaluations: %.5e" % (step, len(full), np.mean(full)))
            logging.info("step: %d, average nelbo bpd out of %d evaluations: %.5e" % (step, len(full_res), np.mean(full_res)))
        if not eval:
            num_data = num_data // 10
        eps_bpd = 1e-5 if config.eval.truncation_time == -1. else config.eval.truncation_time

        def nll(residual, eps, tag):
            (v,) = _passes(config, eval_ds, scaler, num_data,
                           lambda x: (nll_fn(score_model, flow_model, x, None, residual=residual, eps_bpd=eps)[0],))
            m, sd, cnt = _stats(v)
            logging.info("step: %d, [%s] num samples: %d, mean nll bpd: %.5e, std nll bpd: %.5e" % (step, tag, cnt, m, sd))
            return m

        if not config.eval.skip_nll_wrong:
            res['nll_wrong'] = nll(False, eps_bpd, "NLL WRONG w/ eps=%.1e" % eps_bpd)
        res['nll'] = nll(True, eps_bpd, "NLL CORRECT w/ eps=%.1e" % eps_bpd)
        if config.training.truncation_time != 1e-5:
            res['nll_train_eps'] = nll(True, config.training.truncation_time, "NLL CORRECT w/ eps=eps")
        if config.flow.model != 'identity':
            flow_model.train()
    return res
