"""pyglet stand-in (TEST INFRASTRUCTURE ONLY): import-time names and inert constructors, no GUI."""
from . import window, gl  # noqa: F401


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


class app:
    EventLoop = _Inert


class text:
    Label = _Inert


class clock:
    @staticmethod
    def schedule_interval(*a, **k):
        pass

    @staticmethod
    def schedule_once(*a, **k):
        pass

    @staticmethod
    def unschedule(*a, **k):
        pass
