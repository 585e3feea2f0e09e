"""pymunk.pyglet_util stand-in (TEST INFRASTRUCTURE ONLY)."""


class DrawOptions:
    DRAW_SHAPES = 1
    flags = 0
    collision_point_color = (0, 0, 0, 0)
