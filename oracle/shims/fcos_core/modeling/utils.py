def cat(*a, **k):  # name only
    raise NotImplementedError
