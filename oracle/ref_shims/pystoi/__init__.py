def stoi(*a, **k):
    raise RuntimeError("pystoi unavailable")
