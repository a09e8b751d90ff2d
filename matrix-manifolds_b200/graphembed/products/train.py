"""TrainingEngine for product-of-Universal embeddings (reference: graphembed/products/train.py:4-19): the norm
constraint of products.Embedding.stabilize must be imposed every epoch."""
from ..train import TrainingEngine as Base


class TrainingEngine(Base):

    def __init__(self, *args, **kwargs):
        if 'stabilize_every_epochs' in kwargs and kwargs['stabilize_every_epochs'] > 1:
            raise ValueError('For product-space training using the Universal manifold we need to stabilize every '
                             'epoch in order to impose the norm constraint.')
        super().__init__(*args, **kwargs)
        self.stabilize_every_epochs = 1

    def _burnin(self, graph_dataset):
        # the reference's loop reads `self.emb`, which does not exist; the intent -- burn-in
        # epochs each followed by a stabilisation of the embedding being trained -- is what is implemented here
        for epoch in range(1, self.burnin_epochs):
            _ = self._train(graph_dataset, self.alpha, epoch)
            self.embedding.stabilize()
