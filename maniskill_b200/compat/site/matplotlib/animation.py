def FuncAnimation(*a, **kw):
    raise ImportError("matplotlib is not installed; this is an import-level stand-in")
