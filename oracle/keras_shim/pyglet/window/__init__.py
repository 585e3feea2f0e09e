"""pyglet.window stand-in (TEST INFRASTRUCTURE ONLY)."""


class Window:
    def __init__(self, *a, **k):
        raise RuntimeError('GUI is out of scope')


key = mouse = None
