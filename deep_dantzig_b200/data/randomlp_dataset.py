"""Drop-in for the reference's ``data.randomlp_dataset.RandomLPDataset`` (src/data/randomlp_dataset.py:12-128).

Same constructor, same per-item dictionaries, same seed schedule -- but the serial generate/solve/label loop
(``_generate_problems``, :58-63) is one batched GPU call.  Two generation modes:

* ``generator='numpy'`` (default, parity mode): instances are drawn on the host from numpy's legacy global stream
  exactly as the reference does (:37-42, :76-84), so (A, b, c) are bit-identical to the reference's;
* ``generator='philox'`` (throughput mode): instances come from the on-device counter-based generator
  (``ddb_generate_dev``), keyed by ``seed`` and indexed by the instance number.
"""
import numpy as np
import torch
from torch.utils.data.dataset import Dataset

from .. import solver
from .._lib import DEFAULT_THRESHOLD, ST_LOADED, ST_OPTIMAL
from .gurobi_lp import STATUSCODES


def _seed_schedule(seed, num_lps):
    np.random.seed(seed)                                        # randomlp_dataset.py:37
    step = np.random.randint(1, 1000)                           # :41
    return [seed + i * step for i in range(num_lps)]            # :42


def _draw_instance(m, n, seed):
    if seed is not None:
        np.random.seed(seed)                                    # :77-81
    A = np.random.randn(m, n)                                   # :82
    b = A.dot(np.random.randn(n)) + np.absolute(np.random.randn(m))   # :83
    c = np.absolute(np.random.randn(n))                         # :84
    return A, b, c


def _assemble(A, b, c, status, labels_row, objval, seed, with_stats, warn=True):
    """Label list + stats dictionary of create_lp_problem (:91-128)."""
    m, n = A.shape
    success = status in (ST_LOADED, ST_OPTIMAL)
    if success:
        active = np.flatnonzero(labels_row).astype(np.int64)
    else:
        active = []
        if warn:
            print(STATUSCODES.get(status, str(status)))
            print('WARNING: Linear program did not succeed!')
    flags = labels_row if success else np.zeros(m, np.uint8)
    labels = [(i, int(flags[i])) for i in range(m)]
    nactive = int(np.sum(flags))
    if warn and nactive != m - nactive:
        print('WARNING: class inbalance')
    stats = None
    if with_stats:
        stats = {'id': seed, 'm': m, 'n': n, 'eq': 0, 'ineq': m, 'active': len(active), 'sc': status,
                 'objval': objval if success else None, 'success': success}
    return {'A': A, 'b': b, 'c': c, 'active': active, 'labels': labels, 'stats': stats}


class RandomLPDataset(Dataset):

    def __init__(self, m, n, num_lps=1, test=False, seed=3231, generator='numpy', device=0, verbose=False):
        self.m, self.n = m, n
        self.seed, self.test_mode = seed, test
        self.generator, self.device, self.verbose = generator, device, verbose
        self._seeds = _seed_schedule(seed, num_lps)
        self._problems = self._generate_problems()

    def __len__(self):
        return len(self._problems)

    def __getitem__(self, idx):
        p = self._problems[idx % len(self._problems)]
        return {'lp': {'A': p['A'], 'b': p['b'], 'c': p['c']}, 'labels': p['labels']}

    def get_lp_params(self):
        return [p['stats'] for p in self._problems]

    def _generate_problems(self):
        N, m, n = len(self._seeds), self.m, self.n
        if N == 0:
            return []
        if self.generator == 'numpy':
            A = np.empty((N, m, n)); b = np.empty((N, m)); c = np.empty((N, n))
            for i, s in enumerate(self._seeds):
                A[i], b[i], c[i] = _draw_instance(m, n, s)
            res = solver.solve_label_host(A, b, c, DEFAULT_THRESHOLD, device=self.device)
        elif self.generator == 'philox':
            r = solver.generate_solve_label(self.seed, 0, N, m, n, device=self.device, keep_instances=True)
            torch.cuda.synchronize(self.device)
            A, b, c = r['A'].cpu().numpy(), r['b'].cpu().numpy(), r['c'].cpu().numpy()
            res = {k: r[k].cpu().numpy() for k in ('status', 'x', 'obj', 'labels', 'n_active', 'pivots', 'ties')}
        else:
            raise ValueError("generator must be 'numpy' or 'philox'")
        self.solve_result = res
        return [_assemble(A[i], b[i], c[i], int(res['status'][i]), res['labels'][i], float(res['obj'][i]),
                          self._seeds[i], True, warn=self.verbose) for i in range(N)]

    @staticmethod
    def create_lp_problem(m, n, seed=None, with_stats=False, device=0):
        A, b, c = _draw_instance(m, n, seed)
        res = solver.solve_label_host(A[None], b[None], c[None], DEFAULT_THRESHOLD, device=device)
        return _assemble(A, b, c, int(res['status'][0]), res['labels'][0], float(res['obj'][0]), seed, with_stats)
