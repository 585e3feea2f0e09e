"""tensorflow stand-in (TEST INFRASTRUCTURE ONLY): the reference imports tf but never calls it
on the hot path (`/root/reference/src/Networks.py:2`)."""
