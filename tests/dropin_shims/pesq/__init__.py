def pesq(*a, **k):
    raise RuntimeError("pesq is not installed in this image")
