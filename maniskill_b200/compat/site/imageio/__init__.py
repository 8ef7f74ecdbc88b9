"""Import-level stand-in for imageio (video export of the reference's RecordEpisode; outside the hot path)."""


def get_writer(*a, **kw):
    raise ImportError("imageio is not installed; this is an import-level stand-in")


mimsave = imwrite = imread = get_writer
