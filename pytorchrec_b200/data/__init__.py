from .synthetic import criteo_columns, criteo_batch  # noqa: F401
from .tensor_reader import SplitDataset, TensorDataReader  # noqa: F401
