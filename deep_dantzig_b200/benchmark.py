"""Drop-in for the reference's ``benchmark`` driver (src/benchmark.py:27-106): the random-LP dataset (``dataset='randomlp'``,
the hot path) and the reference's own MPS / PLNN datasets (``dataset='plnn' | 'mnist'``).

``save`` keeps the sha1-stamped file naming (:27-44); ``run_experiment_batch`` keeps the parameter names and the result
record schema (:46-95) with the reference's defects fixed (B5: ``params`` undefined, no ``randomlp`` dataset);
``run_benchmark`` expands a dict-of-lists grid (:97-106)."""
import hashlib
import itertools
import json
import os

import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader

from .data.randomlp_dataset import RandomLPDataset
from .ml.models.s2v import Model
from .ml.train import train_net
from .ml.utils import class_weights, collate_randomlp


def save(save_path, stype, res, model=None):
    dset = res['dataset']
    s = json.dumps(res, sort_keys=True)
    stamp = hashlib.sha1(s.encode()).hexdigest()[0:11]
    fname = '%s_%s_res_%s.json' % (stype, dset, stamp)
    with open(os.path.join(save_path, fname), 'w') as outfile:
        json.dump(res, outfile)
    if model is not None:
        fname = '%s_%s_model_%s.json' % (stype, dset, stamp)      # binary despite the suffix, as in the reference
        torch.save(model.state_dict(), os.path.join(save_path, fname))
    return stamp


def run_experiment_batch(dataset, graph, elem_type, num_elems, p, rounds_s2v, epochs, batch_size, learning_rate,
                         momentum, weight_decay, seed, cuda=True, tag=None, m=50, n=20, generator='numpy', device=0, root=None):
    if dataset != 'randomlp':
        return _run_experiment_batch_plnn(dataset, graph, elem_type, num_elems, p, rounds_s2v, epochs, batch_size, learning_rate,
                                          momentum, weight_decay, seed, cuda, tag, device, root)
    params = dict(dataset=dataset, graph=graph, elem_type=elem_type, num_elems=num_elems, p=p, rounds_s2v=rounds_s2v,
                  epochs=epochs, batch_size=batch_size, learning_rate=learning_rate, momentum=momentum,
                  weight_decay=weight_decay, seed=seed, m=m, n=n)
    trainset = RandomLPDataset(m, n, num_lps=num_elems, test=False, seed=seed, generator=generator, device=device)
    testset = RandomLPDataset(m, n, num_lps=max(num_elems // 4, 1), test=True, seed=seed + 7919, generator=generator,
                              device=device)                      # B14: the reference reuses the train seed
    trainloader = DataLoader(trainset, batch_size=batch_size, shuffle=True, collate_fn=collate_randomlp)
    testloader = DataLoader(testset, batch_size=batch_size, shuffle=False, collate_fn=collate_randomlp)
    torch.manual_seed(seed)
    model = Model(graph, p, rounds_s2v, cuda, verbose_init=False)
    dev = torch.device('cuda', device) if cuda else torch.device('cpu')
    weight = torch.tensor(class_weights(trainset), dtype=torch.float32, device=dev)
    criterion = nn.NLLLoss(weight=weight, reduction='sum')        # size_average=False, reduce=True (benchmark.py:75)
    optimizer = optim.SGD(model.parameters(), lr=learning_rate, momentum=momentum, weight_decay=weight_decay)
    results = train_net(model, criterion, optimizer, trainloader, testloader, epochs, batch_size, cuda=cuda, verbose=False)
    out = {'lps': 'randomlp(m=%d,n=%d,seed=%d)' % (m, n, seed), 'results': results}
    d = {'params': params, 'out': out, 'dataset': dataset, 'seed': seed, 'cuda': cuda, 'tag': tag}
    return d, model


def _run_experiment_batch_plnn(dataset, graph, elem_type, num_elems, p, rounds_s2v, epochs, batch_size, learning_rate, momentum,
                               weight_decay, seed, cuda, tag, device, root):
    """The reference's own experiment (src/benchmark.py:46-95): ``DatasetPLNN`` (MPS problem directories below
    ``<root>/data/{plnn, mnist/problems}``; ``root`` defaults to the ROOT environment variable as plnn_dataset.py:189-197),
    class weights from the dataset, per-item training loop.  ``DatasetPLNN.write_infos`` produces the ``.info`` side files with
    the B200 solver when they are missing."""
    from .data.plnn_dataset import DatasetPLNN
    from .ml.train import collate_items
    params = dict(dataset=dataset, graph=graph, elem_type=elem_type, num_elems=num_elems, p=p, rounds_s2v=rounds_s2v,
                  epochs=epochs, batch_size=batch_size, learning_rate=learning_rate, momentum=momentum,
                  weight_decay=weight_decay, seed=seed)
    trainset = DatasetPLNN(dataset, graph, num_elems, elem_type, seed, test=False, root=root)
    testset = DatasetPLNN(dataset, graph, num_elems, elem_type, seed, test=True, root=root)
    trainloader = DataLoader(trainset, batch_size=batch_size, shuffle=True, collate_fn=collate_items)
    testloader = DataLoader(testset, batch_size=batch_size, shuffle=False, collate_fn=collate_items)
    torch.manual_seed(seed)
    model = Model(graph, p, rounds_s2v, cuda, verbose_init=False)
    dev = torch.device('cuda', device) if cuda else torch.device('cpu')
    weight = torch.tensor(trainset.weight, dtype=torch.float32, device=dev)
    criterion = nn.NLLLoss(weight=weight, reduction='sum')
    optimizer = optim.SGD(model.parameters(), lr=learning_rate, momentum=momentum, weight_decay=weight_decay)
    results = train_net(model, criterion, optimizer, trainloader, testloader, epochs, batch_size, cuda=cuda, verbose=False)
    out = {'lps': trainset.get_source_dir(), 'results': results}
    d = {'params': params, 'out': out, 'dataset': dataset, 'seed': seed, 'cuda': cuda, 'tag': tag}
    return d, model


def run_benchmark(params, save_path, cuda=False, tag=None):
    assert(all([type(v) is list for k, v in params.items()]))
    params_keys = list(params.keys())
    stamps = []
    for params_values in itertools.product(*[params[k] for k in params_keys]):
        params_batch = {k: v for k, v in zip(params_keys, params_values)}
        print(','.join(['{0}={1}'.format(k, v) for k, v in params_batch.items()]))
        res, model = run_experiment_batch(**params_batch, cuda=cuda, tag=tag)
        stamps.append(save(save_path, 'benchmark', res, model))
    return stamps
