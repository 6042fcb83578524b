"""tensorflow.python.util.nest stand-in (test infrastructure; see oracle/tf1_shim/tensorflow/__init__.py)."""


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, dict):
        return {k: map_structure(fn, *[s[k] for s in structs]) for k in s0}
    if _is_namedtuple(s0):
        return type(s0)(*[map_structure(fn, *[s[i] for s in structs]) for i in range(len(s0))])
    if isinstance(s0, (list, tuple)):
        return type(s0)(map_structure(fn, *[s[i] for s in structs]) for i in range(len(s0)))
    return fn(*structs)


def flatten(s):
    if isinstance(s, dict):
        out = []
        for k in sorted(s):
            out.extend(flatten(s[k]))
        return out
    if isinstance(s, (list, tuple)):
        out = []
        for v in s:
            out.extend(flatten(v))
        return out
    return [s]
