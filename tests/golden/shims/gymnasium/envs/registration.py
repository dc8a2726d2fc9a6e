registry = {}


def register(id, entry_point, kwargs=None, **_):
    registry[id] = (entry_point, dict(kwargs or {}))


def make(id, **kw):
    entry_point, kwargs = registry[id]
    return entry_point(**{**kwargs, **kw})
