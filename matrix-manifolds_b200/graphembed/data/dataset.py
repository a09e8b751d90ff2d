"""Graph-distance targets as the reference's GraphDataset stores them
(graphembed/data/dataset.py:7-30): a dense N x N matrix of squared distances
divided by their maximum; `dataset[idx]` is the condensed upper triangle of the
idx x idx sub-matrix in batch-position order."""
import torch
from torch.utils.data import Dataset

from ..utils import squareform1


class GraphDataset(Dataset):

    def __init__(self, pdists):
        sq = pdists.pow(2)
        sq = sq / sq.max()
        self.pdists = squareform1(sq).contiguous()

    @property
    def device(self):
        return self.pdists.device

    def to(self, *args, **kwargs):
        self.pdists = self.pdists.to(*args, **kwargs).contiguous()
        return self

    def __getitem__(self, node_indices=None):
        if node_indices is None:
            sub = self.pdists
        else:
            idx = node_indices.to(self.device)
            sub = self.pdists[idx][:, idx]
        i, j = torch.triu_indices(len(sub), len(sub), 1, device=self.device)
        return sub[i, j]

    def __len__(self):
        return len(self.pdists)
