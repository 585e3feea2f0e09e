"""pyglet stand-in (TEST INFRASTRUCTURE ONLY): import-time names only, no GUI."""
from . import window, gl  # noqa: F401
