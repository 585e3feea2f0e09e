"""pyglet.gl stand-in (TEST INFRASTRUCTURE ONLY)."""
