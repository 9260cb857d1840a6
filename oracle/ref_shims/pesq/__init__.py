def pesq(*a, **k):
    raise RuntimeError("pesq unavailable")
