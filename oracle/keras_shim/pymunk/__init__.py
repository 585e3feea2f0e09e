"""pymunk stand-in (TEST INFRASTRUCTURE ONLY): lets `/root/reference/src/main.py`, `JengaBuilder.py` and `TowerCreator.py` be
imported and their LAYOUT code (create_world, put_box(es)) be executed, so that the relation-construction loops (main.py:66-81)
and the layout samplers (spwgnn_b200/synth.py) can be pinned to the reference.  No physics: bodies only remember what they
were given; stepping the space is out of scope."""


class Vec2d(tuple):
    def __new__(cls, x=0.0, y=0.0):
        return super().__new__(cls, (x, y))

    @property
    def x(self):
        return self[0]

    @property
    def y(self):
        return self[1]


def moment_for_box(mass, size):
    return mass * (size[0] ** 2 + size[1] ** 2) / 12.0


class Body:
    def __init__(self, mass=0.0, moment=0.0):
        self.mass, self.moment = mass, moment
        self.position = Vec2d(0.0, 0.0)
        self.velocity = Vec2d(0.0, 0.0)
        self.angle = 0.0


class _Shape:
    friction = 0.0


class Segment(_Shape):
    def __init__(self, body, a, b, radius):
        self.body, self.a, self.b, self.radius = body, a, b, radius


class Poly(_Shape):
    def __init__(self, body, size):
        self.body, self.size = body, size
        self.area = float(size[0]) * float(size[1])

    @staticmethod
    def create_box(body, size):
        return Poly(body, size)


class Space:
    def __init__(self):
        self.static_body = Body()
        self.gravity = Vec2d(0.0, 0.0)
        self.sleep_time_threshold = 0.0
        self.items = []

    def add(self, *objs):
        self.items.extend(objs)

    def remove(self, *objs):
        for o in objs:
            if o in self.items:
                self.items.remove(o)

    def step(self, dt):
        raise RuntimeError('physics is out of scope')
