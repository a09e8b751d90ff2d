from .embedding import Embedding
from .train import TrainingEngine
