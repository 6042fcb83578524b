"""Model plugins, discovered like reference run.py:320 does (`util.dynamic_load_module(models, prefix="models")`):
importing this package imports every plugin module, each of which self-registers (models/transformer.py:289)."""
from . import model  # noqa: F401
from . import transformer  # noqa: F401
