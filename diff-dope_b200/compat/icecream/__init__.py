"""icecream stand-in: ic(*args) prints its arguments and returns them."""
import pprint


def ic(*args):
    for a in args:
        print("ic| " + (a if isinstance(a, str) else pprint.pformat(a)))
    if not args:
        return None
    return args[0] if len(args) == 1 else args
