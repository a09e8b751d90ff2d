"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's epoch loop
(graphembed/graphembed/train.py): `_train` (:198-228: one randperm per epoch from the global CPU generator, node
batches, tail batches shorter than drop_last_n dropped, zero_grad / backward / step), `_run_epoch` (:168-196:
check-best, stabilize, validate) and `_validate` (:230-265: sqrt of both distance vectors, pearsonr and
average_distortion), on top of the oracle's manifolds, losses and optimizer steps.

Parity status: PINNED against tests/golden/engine_runs_f64.npz, which holds the step losses, per-epoch metrics, final
points and best-loss bookkeeping of the real TrainingEngine run in the build container
(tests/golden/make_golden.py::make_engine_runs); see tests/test_oracle_pinned.py."""
import torch

import manifolds_oracle as O


def run_engine(oracles, xs, scales, hops_condensed, loss_fn, step_fn, n_epochs, alpha, batch_size=None,
               drop_last_n=50, stabilize_every_epochs=None, seed=None):
    """oracles/xs/scales: one entry per factor (xs are plain tensors, updated functionally).
    loss_fn(targets, mdists, alpha, epoch) -> scalar; step_fn(f, oracle, x, grad, state) -> new x.
    Returns dict(step_loss, pearsonr, average_distortion, xs, best=(epoch, loss))."""
    dtype = xs[0].dtype
    cond = O.dataset_targets(hops_condensed, dtype)
    n = xs[0].shape[0]
    dense = torch.zeros(n, n, dtype=dtype)
    iu = torch.triu_indices(n, n, 1)
    dense[iu[0], iu[1]] = cond
    dense = dense + dense.T
    states = [dict() for _ in xs]
    out = dict(step_loss=[], pearsonr=[], average_distortion=[])
    best = (0, 1e8)
    if stabilize_every_epochs is None:
        stabilize_every_epochs = n_epochs + 1
    if seed is not None:
        torch.manual_seed(seed)
    bs = n if batch_size is None else min(n, batch_size)
    for epoch in range(1, n_epochs + 1):
        perm = torch.randperm(n)
        total = 0.0
        for i in range(0, n, bs):
            idx = perm[i:i + bs]
            if len(idx) < drop_last_n:
                break
            leaves = [x.clone().requires_grad_() for x in xs]
            m = O.product_dist2(oracles, leaves, scales, lambda o, x: o.pdist2(x[idx]))
            loss = loss_fn(O.batch_targets(dense, idx), m, alpha, epoch)
            loss.backward()
            xs = [step_fn(f, o, x, leaf.grad, st).detach()
                  for f, (o, x, leaf, st) in enumerate(zip(oracles, xs, leaves, states))]
            out['step_loss'].append(loss.item() / len(idx))
            total += loss.item()
        if total < best[1]:
            best = (epoch, total)
        if epoch % stabilize_every_epochs == 0:
            xs = [o.projx(x) for o, x in zip(oracles, xs)]
        vm = O.validation_metrics(oracles, xs, scales, cond)
        out['pearsonr'].append(vm['pearsonr'])
        out['average_distortion'].append(vm['average_distortion'])
    out['xs'] = xs
    out['best'] = best
    return out
