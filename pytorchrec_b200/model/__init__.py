"""Models with the reference's ``torchrec.model`` surface (``IModel`` + name registry,
torchrec/model/models.py:8-30)."""
from typing import Dict, Type

from .IModel import History, IModel
from .ctr import DCN, DIN, FM, DeepFM
from .mf import NCF, SVDPP, FunkSVD

_model_classes: Dict[str, Type[IModel]] = {
    "funksvd": FunkSVD,
    "svdpp": SVDPP,
    "ncf": NCF,
    "fm": FM,
    "deepfm": DeepFM,
    "dcn": DCN,
    "din": DIN,
}
model_name_list = _model_classes.keys()


def get_model_type(model_name: str) -> Type[IModel]:
    if (not isinstance(model_name, str)) or (model_name not in _model_classes):
        raise ValueError(f"invalid model_name: {model_name}")
    return _model_classes[model_name]


__all__ = ["IModel", "History", "FM", "DeepFM", "DCN", "DIN", "FunkSVD", "SVDPP", "NCF", "get_model_type", "model_name_list"]
