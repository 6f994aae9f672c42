"""plum.dispatch stand-in: a global registry keyed by function name (plum's default dispatcher does the same:
galax registers every overload with a bare ``@dispatch``, potential/_src/register_funcs.py:11,33,86)."""
import inspect
import typing

REGISTRY: dict[str, list] = {}


def _matches(ann, value) -> bool:
    if ann is inspect.Parameter.empty or ann is typing.Any or ann is object:
        return True
    try:
        return isinstance(value, ann)
    except TypeError:
        return True


def _specificity(ann) -> int:
    return 0 if ann in (inspect.Parameter.empty, typing.Any, object) else 1


class _Generic:
    def __init__(self, name):
        self.name = name

    def __call__(self, *args, **kwargs):
        best, best_score = None, -1
        for fn in REGISTRY[self.name]:
            sig = inspect.signature(fn)
            try:
                bound = sig.bind(*args, **kwargs)
            except TypeError:
                continue
            pos = [p for p in sig.parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
            if len(args) > len(pos) and not any(p.kind == p.VAR_POSITIONAL for p in sig.parameters.values()):
                continue
            hints = {k: v.annotation for k, v in sig.parameters.items()}
            if any(isinstance(h, str) for h in hints.values()):  # plum cannot resolve names local to another function
                raise TypeError(f"{fn.__qualname__}: unresolved (string) annotation {hints}")
            if all(_matches(hints[k], v) for k, v in bound.arguments.items() if k in hints and k in [p.name for p in pos]):
                score = sum(_specificity(hints[p.name]) for p in pos if p.name in bound.arguments)
                if score > best_score:
                    best, best_score = fn, score
        if best is None:
            raise LookupError(f"no overload of {self.name} for {[type(a).__name__ for a in args]}")
        return best(*args, **kwargs)


def dispatch(fn):
    REGISTRY.setdefault(fn.__name__, []).append(fn)
    return _Generic(fn.__name__)
