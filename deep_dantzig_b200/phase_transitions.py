"""Drop-in for the reference's sweep driver (src/phase_transitions.py:30-63) with its missing glue defined
(B4: ``wrap_params`` / ``run_experiment``), plus the m/n-ratio x density sweep of BASELINE.json config 3 (a new
capability behind this entry point: the reference sweeps the embedding dimension only)."""
import itertools
import json

import numpy as np
import torch

from . import parallel, solver
from .benchmark import run_experiment_batch, save


def wrap_params(ep, bs, t, lr, mtm, wd, p, graph='bipartite', num_elems=64, m=50, n=20):
    return {'dataset': 'randomlp', 'graph': graph, 'elem_type': 'lp', 'num_elems': num_elems, 'p': p, 'rounds_s2v': t,
            'epochs': ep, 'batch_size': bs, 'learning_rate': lr, 'momentum': mtm, 'weight_decay': wd, 'm': m, 'n': n}


def run_experiment(params, dataset, seed, cuda, tag):
    """Return schema consumed at phase_transitions.py:52-53: res['out']['acc'], res['out']['losses']['total'] (json)."""
    assert dataset == params['dataset']
    d, model = run_experiment_batch(**params, seed=seed, cuda=cuda, tag=tag)
    hist = d['out']['results']['test']
    d['out']['acc'] = hist[-1]['accuracy']
    d['out']['losses'] = {'total': json.dumps([h['total_loss'] for h in hist])}
    return d, model


def compute_phaseTransitions(save_path, dataset, benchmark_params, cuda=False, tag=None, **lp_kwargs):
    sds, eps, bss = benchmark_params['seeds'], benchmark_params['epochs'], benchmark_params['batch_sizes']
    ts, lrs, mtms = benchmark_params['rounds_s2v'], benchmark_params['learning_rates'], benchmark_params['momentums']
    wds, ps = benchmark_params['weight_decays'], benchmark_params['ps']
    records = []
    for seed, ep, bs, t, lr, mtm, wd, p0 in itertools.product(sds, eps, bss, ts, lrs, mtms, wds, ps):
        acc, accs, losses, p = 1.0, {}, {}, p0
        while p > 1 and acc > 0.5:
            params = wrap_params(ep, bs, t, lr, mtm, wd, p, **lp_kwargs)
            print(','.join(['{0}={1}'.format(k, v) for k, v in params.items()]))
            res, model = run_experiment(params, dataset, seed, cuda, tag)
            acc = accs[p] = res['out']['acc']
            losses[p] = json.loads(res['out']['losses']['total'])[-1]
            p = p - 1
        pt = {'params': wrap_params(ep, bs, t, lr, mtm, wd, p0, **lp_kwargs), 'out': {'accs': accs, 'losses': losses},
              'dataset': dataset, 'seed': seed, 'cuda': cuda, 'tag': tag}
        save(save_path, 'pt', pt, None)
        records.append(pt)
    return records


def warm_up_sweep(n=100, ratios=(1.25, 1.5, 2.0, 3.0, 4.0), densities=(1.0, 0.5, 0.1), chunk=2048, device=0):
    """One untimed chunk of EVERY cell on this rank (kernels loaded, scratch sized for the chunk), whatever cells the timed
    sweep will deal to it."""
    for r in ratios:
        for d in densities:
            solver.generate_solve_label(0, 0, chunk, int(round(r * n)), n, density=d, device=device)
    torch.cuda.synchronize(device)


def sweep_ratio_density(n=100, ratios=(1.25, 1.5, 2.0, 3.0, 4.0), densities=(1.0, 0.5, 0.1), per_cell=10000, chunk=2048,
                        key=0, device=0):
    """BASELINE.json config 3: grid over m/n and density of A; the (cell, chunk) list is dealt round-robin to the ranks
    and every rank solves its chunks with the fused generate -> solve -> label call; per-cell statistics are summed with
    one all-reduce.  Returns {(ratio, density): {'instances', 'optimal', 'unbounded', 'other', 'mean_pivots', 'ties'}}
    on every rank."""
    cells = [(r, d) for r in ratios for d in densities]
    work = []
    for ci, (r, d) in enumerate(cells):
        for lo in range(0, per_cell, chunk):
            work.append((ci, lo, min(chunk, per_cell - lo)))
    mine = parallel.shard_round_robin(work)
    dev = torch.device('cuda', device)
    stats = torch.zeros(len(cells), 6, dtype=torch.float64, device=dev)
    for ci, lo, cnt in mine:
        r, d = cells[ci]
        m = int(round(r * n))
        res = solver.generate_solve_label(key + ci, lo, cnt, m, n, density=d, device=device)
        st = res['status']
        stats[ci, 0] += cnt
        stats[ci, 1] += (st == 2).sum()
        stats[ci, 2] += (st == 5).sum()
        stats[ci, 3] += ((st != 2) & (st != 5)).sum()
        stats[ci, 4] += res['pivots'][:, 3].sum()
        stats[ci, 5] += (res['ties'] * (st == 2)).sum()
    if parallel.world()[1] > 1:
        torch.distributed.all_reduce(stats)
    stats = stats.cpu().numpy()
    out = {}
    for ci, cell in enumerate(cells):
        s = stats[ci]
        out[cell] = {'instances': int(s[0]), 'optimal': int(s[1]), 'unbounded': int(s[2]), 'other': int(s[3]),
                     'mean_pivots': float(s[4] / max(s[0], 1)), 'ties': int(s[5]), 'm': int(round(cell[0] * n)), 'n': n}
    return out
