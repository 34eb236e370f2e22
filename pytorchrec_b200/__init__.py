"""pytorchrec_b200 — B200 (sm_100a) implementation of PyTorchRec's data-parallel hot path:
sparse-feature embedding lookup with pooling, the FM / DCN / DIN interaction layers and the sparse
optimizer update, behind the reference's ``feature_column`` / ``model`` / ``optim`` Python surface.
The compute lives in ``libptrec_b200.so`` (C ABI: ``include/ptrec_b200.h``); there is no CPU path.
"""
__version__ = "0.1.0"

from . import feature_column, loss, metric, model, optim, utils  # noqa: F401


def _drain_at_exit() -> None:
    """Models are reference cycles (``weight._ptrec_table``), so the pinned staging buffers, copy streams and events they
    own are normally released by the cyclic collector — at interpreter shutdown in arbitrary order, when a pinned buffer
    whose last copy was issued on an already-destroyed stream aborts the process.  Collect them while CUDA is alive."""
    try:
        import gc

        import torch
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            torch.cuda.synchronize()
        from .model.IModel import _LIVE_MODELS
        for m in list(_LIVE_MODELS):   # release the ingest state in a fixed order: batches and events, then packers and streams
            m.__dict__.get("_prefetched", []).clear()
            m.__dict__.get("_prefetch", {}).clear()
        gc.collect()
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            torch.cuda.synchronize()
    except Exception:  # noqa: BLE001 — never let shutdown hygiene raise
        pass


import atexit  # noqa: E402

atexit.register(_drain_at_exit)
