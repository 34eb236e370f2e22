from .synthetic import criteo_columns, criteo_batch  # noqa: F401
