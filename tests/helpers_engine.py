"""Shared description of the four golden TrainingEngine runs (tests/golden/make_golden.py::make_engine_runs)."""
import numpy as np
import torch

from helpers import GOLDEN

RUNS = {
    # tag: (factors, objective, optimizer, engine kwargs)
    'spd3_full': ([('spd', 3)], 'quotient', ('rsgd', dict(lr=0.01, max_grad_norm=20, exact=True)), dict(alpha=1.0)),
    'spd3_batched': ([('spd', 3)], 'quotient', ('rsgd', dict(lr=0.01, max_grad_norm=20, exact=True)),
                     dict(alpha=1.0, batch_size=52)),
    'spd2_kl': ([('spd', 2)], 'kl', ('radam', dict(lr=0.01, max_grad_norm=100, exact=False)),
                dict(alpha=10.0, stabilize_every_epochs=2)),
    'prod_radam': ([('spd', 3), ('lorentz', 5)], 'quotient', ('radam', dict(lr=0.01, max_grad_norm=100, exact=True)),
                   dict(alpha=1.0)),
}
ENGINE_SEED = 1234
N_EPOCHS = 3


def load_engine_golden():
    with np.load(f'{GOLDEN}/engine_runs_f64.npz') as z:
        return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in 'fiu' else z[k]) for k in z.files}
