"""pymunk stand-in (TEST INFRASTRUCTURE ONLY): lets `/root/reference/src/main.py` be imported so
its relation-construction loops (main.py:66-81) can be executed; physics is out of scope."""


class Vec2d(tuple):
    def __new__(cls, x=0.0, y=0.0):
        return super().__new__(cls, (x, y))
