from .combo import estimate_logz, fmt_val_err  # noqa: F401
