"""keras.regularizers stand-in (TEST INFRASTRUCTURE ONLY)."""


def l2(l=0.01):
    return ('l2', l)
