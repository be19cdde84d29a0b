"""astrolink_b200: B200-native implementation of AstroLink's hot path behind the
reference's `AstroLink(P, ...).run()` interface."""
from .astrolink import AstroLink  # noqa: F401
from . import io  # noqa: F401

__all__ = ["AstroLink", "io"]
__version__ = "0.1.0"
