"""sapien.wrapper.coacd: approximate convex decomposition is not available; the file is used as it is (its parts, when it has several)."""


def do_coacd(filename, **params):
    return filename
