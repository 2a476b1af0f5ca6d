from . import save_results  # noqa: F401
