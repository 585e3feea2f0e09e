"""pyglet.window stand-in (TEST INFRASTRUCTURE ONLY): a window that never opens."""


class Window:
    def __init__(self, *a, **k):
        pass

    def set_caption(self, *a, **k):
        pass


class FPSDisplay:
    def __init__(self, *a, **k):
        pass


key = mouse = None
