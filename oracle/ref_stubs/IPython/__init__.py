def embed(*a, **k):
    raise NotImplementedError("stub")
