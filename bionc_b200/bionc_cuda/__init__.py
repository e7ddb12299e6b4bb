from .model_table import ModelTable, UnsupportedModelError  # noqa: F401
