def close(*args, **kwargs):
    pass


def __getattr__(name):
    raise AttributeError("matplotlib stub: pyplot.%s is plotting and not part of the accelerated path" % name)
