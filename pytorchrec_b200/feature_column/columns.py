"""Feature-column classes.  Behavioural mirror of the reference (file:line cited per class); the
wire format is the reference's: a batch is ``Dict[str, Tensor]`` keyed by feature name, id 0 = PAD.
"""
from abc import ABC, abstractmethod
from enum import Enum, unique
from typing import Any, Dict, List, Optional

from torch import Tensor


class FeatureColumn(ABC):
    """Base column with a free-form info dict (torchrec/feature_column/FeatureColumn.py:10-26)."""

    def __init__(self):
        self._info: Dict[str, Any] = {}

    def set_info(self, key: str, value: Any) -> None:
        self._info[key] = value

    def get_info(self) -> Dict[str, Any]:
        return self._info

    @abstractmethod
    def get_feature_data(self, *args, **kwargs) -> Tensor:
        """Extract this column's tensor from a batch."""


class CategoricalColumn(FeatureColumn, ABC):
    """Discrete column; ``category_num`` is the table height (CategoricalColumn.py:9-14)."""

    def __init__(self, category_num: int):
        super().__init__()
        self.category_num = category_num


class CategoricalColumnWithIdentity(CategoricalColumn):
    """Integer ids in ``[0, category_num)``; returns ``batch[name].long()``
    (CategoricalColumnWithIdentity.py:12-22).  ``from_series`` sizes the table as ``max + 1``
    (:24-37)."""

    def __init__(self, category_num: int, feature_name: str):
        super().__init__(category_num)
        self.feature_name = feature_name

    def get_feature_data(self, batch: Dict[str, Tensor]) -> Optional[Tensor]:
        return batch.get(self.feature_name).long()

    @staticmethod
    def from_series(feature_name: str, series, other_info: Optional[Dict[str, Any]] = None):
        from pandas.api import types
        assert types.is_integer_dtype(series), series.dtypes
        column = CategoricalColumnWithIdentity(category_num=int(series.max()) + 1, feature_name=feature_name)
        column.set_info("min", series.min())
        column.set_info("max", series.max())
        for key, value in (other_info or {}).items():
            column.set_info(key, value)
        return column

    def __repr__(self):
        s = f"name: {self.feature_name}, category_num: {self.category_num}"
        for key, value in self.get_info().items():
            s += f", {key}: {value}"
        return s

    __str__ = __repr__


class CrossedColumn(CategoricalColumn):
    """Mixed-radix cross of categorical columns: id = sum_i coeff_i * id_i, with
    ``coeff_i = prod_{j>i} category_num_j`` (CrossedColumn.py:14-27).  Integer, bit-exact."""

    def __init__(self, categorical_columns: List[CategoricalColumn]):
        total = 1
        for c in categorical_columns:
            total *= c.category_num
        super().__init__(total)
        self.categorical_columns = categorical_columns
        coeffs = [1] * len(categorical_columns)
        for i in range(len(categorical_columns) - 2, -1, -1):
            coeffs[i] = coeffs[i + 1] * categorical_columns[i + 1].category_num
        self.coefficients = coeffs

    def get_feature_data(self, batch: Dict[str, Any]) -> Tensor:
        out = None
        for coeff, col in zip(self.coefficients, self.categorical_columns):
            term = coeff * col.get_feature_data(batch)
            out = term if out is None else out + term
        return out


class DenseColumn(FeatureColumn, ABC):
    """Marker base for columns that feed dense layers directly (DenseColumn.py:9-11)."""


@unique
class NormalizationMode(Enum):
    """NormalizationMode.py:8-12"""
    NOP = "nop"
    MAX_MIN = "max_min"
    Z_SCORE = "z_score"


class NumericColumn(DenseColumn):
    """Float feature with optional max-min / z-score normalisation (NumericColumn.py:14-34)."""

    def __init__(self, feature_name: str, min_value: float, max_value: float, mean_value: float,
                 std_value: float):
        super().__init__()
        self.feature_name = feature_name
        self.min_value = min_value
        self.max_value = max_value
        self.mean_value = mean_value
        self.std_value = std_value

    def get_feature_data(self, batch: Dict[str, Any],
                         normalization_mode: NormalizationMode = NormalizationMode.NOP) -> Tensor:
        x = batch[self.feature_name].float()
        if normalization_mode == NormalizationMode.NOP:
            return x
        if normalization_mode == NormalizationMode.MAX_MIN:
            return (x - self.min_value) / (self.max_value - self.min_value)
        if normalization_mode == NormalizationMode.Z_SCORE:
            return (x - self.mean_value) / self.std_value
        raise Exception("NormalizationMode is wrong!")

    @staticmethod
    def from_series(feature_name: str, series):
        from pandas.api import types
        assert types.is_numeric_dtype(series), series.dtypes
        return NumericColumn(feature_name=feature_name, min_value=series.min(), max_value=series.max(),
                             mean_value=series.mean(), std_value=series.std())

    def __repr__(self):
        s = (f"name: {self.feature_name}, min: {self.min_value}, max: {self.max_value}, "
             f"mean: {self.mean_value}, std: {self.std_value}")
        for key, value in self.get_info().items():
            s += f", {key}: {value}"
        return s

    __str__ = __repr__
