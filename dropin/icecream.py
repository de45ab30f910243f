"""`from icecream import ic` shim (prune.py:24 and others): pass-through debug print."""


def ic(*args):
    if not args:
        return None
    return args[0] if len(args) == 1 else args
