def _missing(*a, **kw):
    raise ImportError("matplotlib is not installed; this is an import-level stand-in")


figure = subplots = plot = show = imshow = savefig = close = axis = title = _missing


def get_cmap(*a, **kw):
    _missing()
