from .dense import MLP, Dense
from .embedding import EmbeddingTable, MultiTableEmbedding
from .interaction import AttentionPooling, CrossNet, FMSecondOrder

__all__ = ["Dense", "MLP", "EmbeddingTable", "MultiTableEmbedding", "FMSecondOrder", "CrossNet", "AttentionPooling"]
