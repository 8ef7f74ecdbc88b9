def display(*a, **kw):
    pass


class Video:
    def __init__(self, *a, **kw):
        pass


class HTML(Video):
    pass
