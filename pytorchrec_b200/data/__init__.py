from .synthetic import amazon_batch, amazon_columns, criteo_batch, criteo_columns  # noqa: F401
from .tensor_reader import SplitDataset, TensorDataReader  # noqa: F401
