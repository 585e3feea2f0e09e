"""keras.utils stand-in (TEST INFRASTRUCTURE ONLY)."""


class Sequence:
    pass
