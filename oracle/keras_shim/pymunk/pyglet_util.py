"""pymunk.pyglet_util stand-in (TEST INFRASTRUCTURE ONLY)."""
