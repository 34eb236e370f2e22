"""Feature columns — same public surface as the reference's ``torchrec.feature_column``
(torchrec/feature_column/__init__.py:4-10).  Columns hold no parameters: they pull one tensor out
of the batch dict.  Embedding tables are built *from* them (``model.layer.MultiTableEmbedding``).
"""
from .columns import (CategoricalColumn, CategoricalColumnWithIdentity, CrossedColumn, DenseColumn,
                      FeatureColumn, NormalizationMode, NumericColumn)

__all__ = ["CategoricalColumn", "CategoricalColumnWithIdentity", "CrossedColumn", "DenseColumn",
           "FeatureColumn", "NormalizationMode", "NumericColumn"]
