"""Drop-in for the reference's ``data.plnn_dataset.DatasetPLNN`` (src/data/plnn_dataset.py:23-187): LPs stored as MPS files
under ``<root>/problem_*/`` directories, each with a ``.info`` side file (active constraints + optimal x).  The reference
lets Gurobi read the files and expects the ``.info`` files to exist; here :mod:`data.mps` reads them and
``DatasetPLNN.write_infos`` produces the side files with the B200 solver (``LinProg.solve_mps``)."""
import os

import numpy as np
from torch.utils.data.dataset import Dataset

from .gurobi_lp import LinProg as LP
from .mps2numpy import mps2numpy


class DatasetPLNN(Dataset):
    TRAIN_PCT = 0.90

    def __init__(self, dataset='mnist', graph='bipartite', num_elems=None, elem_type='lp', seed=1111, test=False, root=None):
        """Same arguments as the reference plus ``root`` (the reference takes it from the ROOT environment variable,
        plnn_dataset.py:189-197).  elem_type: 'property' (split the problem directories), 'lp' (split the LPs of the
        largest directory), 'constraint' (the largest LP of the first directory)."""
        if elem_type not in ('property', 'lp', 'constraint'):
            raise ValueError('elem_type not recognised')
        self.test, self.seed, self.graph_structure = test, seed, graph
        self._root = root
        np.random.seed(self.seed)
        property_dirs = list(np.random.permutation(self.get_prop_dirs(dataset, root)))
        ok = lambda d: d['num_ineq'] > 0                                     # noqa: E731
        if elem_type in ('property', 'lp'):
            if elem_type == 'property':
                train_props, test_props = self._train_test_split(property_dirs, num_elems, self.TRAIN_PCT)
                train_lps = [os.path.join(d, f) for d in train_props for f in sorted(os.listdir(d)) if f.endswith('.mps')]
                test_lps = [os.path.join(d, f) for d in test_props for f in sorted(os.listdir(d)) if f.endswith('.mps')]
            else:
                pdir = max(property_dirs, key=lambda d: len(os.listdir(d)))   # the directory with the most problems
                fs = [os.path.join(pdir, f) for f in sorted(os.listdir(pdir)) if f.endswith('.mps')]
                fs_ok = [d['path'] for d in (LP.ineq_num(f) for f in fs) if ok(d)]
                print('\t %d/%d LPs with inequality constraints' % (len(fs_ok), len(fs)))
                train_lps, test_lps = self._train_test_split(fs_ok, num_elems, self.TRAIN_PCT)
            print('Test set' if self.test else 'Train set')
            fs = test_lps if self.test else train_lps
            fs_ineq = [LP.ineq_num(f) for f in fs]
            fs_ok = [d['path'] for d in fs_ineq if ok(d)]
            print('\t %d/%d LPs with inequality constraints' % (len(fs_ok), len(fs)))
            print('(%s) %d/%d problems' % (elem_type, len(fs), len(train_lps) + len(test_lps)))
            stats = [d for d in fs_ineq if ok(d)]
        else:
            pdir = property_dirs[0]
            fs = [os.path.join(pdir, f) for f in sorted(os.listdir(pdir)) if f.endswith('.mps')]
            d = max((LP.ineq_num(f) for f in fs), key=lambda x: x['num_constrs'])
            fs_ok, stats = [d['path']], [d]
        self._fpaths = fs_ok
        self.n_pos = sum(d['num_pos'] for d in stats)
        self.n_neg = sum(d['num_neg'] for d in stats)
        self.n_eq = sum(d['num_eq'] for d in stats)
        self.n_ineq = sum(d['num_ineq'] for d in stats)
        self.n_inact_ineq = sum(d['num_inactive_ineq'] for d in stats)
        self.n_total = self.n_pos + self.n_neg
        self.print_baselines()
        self.weight = [self.n_pos / max(self.n_total, 1), self.n_neg / max(self.n_total, 1)]     # plnn_dataset.py:116
        print(self.weight)
        self._items = [self._fpath2item(f) for f in self._fpaths]

    def _fpath2item(self, fpath):
        if self.graph_structure == 'complete':
            return LP.getitem_complete(fpath)
        if self.graph_structure == 'bipartite':
            return LP.getitem_bipartite(fpath)
        raise ValueError

    def print_baselines(self):
        print('%s set' % ('Test' if self.test else 'Train'))
        print('%d total LPs' % len(self._fpaths))
        print('\t %d total constraints\n\t %d positive\n\t %d negative\n\t %d equality\n\t %d inequality\n\t %d active_inequality\n'
              % (self.n_total, self.n_pos, self.n_neg, self.n_eq, self.n_ineq, self.n_inact_ineq))

    def override_fpaths(self, lps):
        self._fpaths = [os.path.join(h, f) for h in lps for f in sorted(os.listdir(h)) if f.endswith('.mps')]

    def __len__(self):
        return len(self._items)

    def __getitem__(self, idx):
        return self._items[idx]

    def get_source_dir(self):
        return list(set(os.path.dirname(f) for f in self._fpaths))

    def _train_test_split(self, items, num_items=None, TRAIN_PCT=0.90):
        """plnn_dataset.py:166-187."""
        assert len(items) > 1
        items = list(np.random.permutation(items))
        if num_items is not None:
            assert num_items > 1
            if num_items > len(items):
                raise ValueError('num_items > len(items)')
            items = items[:num_items]
        n_train = min(int(len(items) * TRAIN_PCT), len(items) - 1)
        n_test = len(items) - n_train
        assert 0 < n_train < len(items) and 0 < n_test < len(items)
        test_items = items[n_train:] if n_test > 1 else [items[-1]]
        train_items = items[:n_train] if n_train > 1 else [items[0]]
        return train_items, test_items

    @staticmethod
    def get_lp_dir(dataset=None, root=None):
        root = root if root is not None else os.environ.get('ROOT', '.')
        return os.path.join(root, 'data/mnist/problems') if dataset == 'mnist' else os.path.join(root, 'data/plnn')

    @staticmethod
    def get_prop_dirs(dataset, root=None):
        h = DatasetPLNN.get_lp_dir(dataset, root)
        return [os.path.join(h, f) for f in sorted(os.listdir(h)) if f.startswith('problem_')]

    @staticmethod
    def get_mps_paths(ext='.mps', num_lps=None, seed=1111, dataset=None, root=None):
        ds = [d for d in DatasetPLNN.get_prop_dirs(dataset, root) if os.path.isdir(d)]
        if num_lps and num_lps < len(ds):
            np.random.seed(seed)
            ds = np.random.choice(ds, size=num_lps, replace=False).tolist()
        return [os.path.join(d, f) for d in ds for f in sorted(os.listdir(d)) if f.endswith(ext)], ds

    @staticmethod
    def write_infos(dataset=None, root=None, device=0, overwrite=False):
        """Solve every MPS file below the problem directories on the GPU and write its ``.info`` side file (what the
        reference obtained from an offline Gurobi run).  Returns the number of files written."""
        fs, _ = DatasetPLNN.get_mps_paths(dataset=dataset, root=root)
        done = 0
        for f in fs:
            if overwrite or not os.path.exists(os.path.splitext(f)[0] + '.info'):
                LP.solve_mps(f, device=device, write_info=True)
                done += 1
        return done

    @staticmethod
    def extract_lp_problem(fpath, standardize=True, device=0):
        """plnn_dataset.py:207-260: solve one MPS file and report status, names, slacks, x, objective and the solve time."""
        from timeit import default_timer as timer
        from .mps import read_mps
        A, b, c, ops, obj = mps2numpy(fpath, standardize)
        lp = LP(A, b, c, obj, ops, device=device)
        start = timer()
        lp.optimize()
        total_time = timer() - start
        model = read_mps(fpath)
        constr_names = [q.ConstrName for q in model.getConstrs()]
        var_names = [v.VarName for v in model.getVars()]
        d = {'sc': lp.model.status, 'constrs': constr_names, 'vars': var_names, 'num_bounds': A.shape[0] - len(constr_names),
             'constr_sense': {q.ConstrName: q.Sense for q in model.getConstrs()}, 'time': total_time, 'source': fpath,
             'num_constrs': len(constr_names), 'num_vars': len(var_names), 'upper_bounds': {}, 'lower_bounds': {},
             'slacks': {}, 'x': {}, 'obj_val': None}
        if lp.model.status == 2:
            slack = b - A.dot(lp.x)
            d['slacks'] = {name: float(slack[i]) for i, name in enumerate(constr_names)}
            d['x'] = {name: float(lp.x[j]) for j, name in enumerate(var_names)}
            d['obj_val'] = lp.model.objVal
        return d
