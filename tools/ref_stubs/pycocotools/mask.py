def decode(*a, **k):
    raise NotImplementedError('stub')
