"""sbr_b200 -- import alias of the package directory ``sequence-based-recommendations_b200/``.

The package directory keeps the repository's name (which is not a valid Python identifier);
this stub makes it importable as ``sbr_b200`` by pointing ``__path__`` at it.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
__path__ = [_os.path.join(_ROOT, "sequence-based-recommendations_b200")]
__version__ = "0.1.0"
