from . import creation


def Box(extents=(1, 1, 1), transform=None, **kw):
    return creation.box(extents=extents, transform=transform)
