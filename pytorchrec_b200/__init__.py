"""pytorchrec_b200 — B200 (sm_100a) implementation of PyTorchRec's data-parallel hot path:
sparse-feature embedding lookup with pooling, the FM / DCN / DIN interaction layers and the sparse
optimizer update, behind the reference's ``feature_column`` / ``model`` / ``optim`` Python surface.
The compute lives in ``libptrec_b200.so`` (C ABI: ``include/ptrec_b200.h``); there is no CPU path.
"""
__version__ = "0.1.0"

from . import feature_column, loss, metric, model, optim, utils  # noqa: F401
