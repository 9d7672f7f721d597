import contextlib


def trace(*_a, **_k):
    return contextlib.nullcontext()


def start_trace(*_a, **_k):
    return None


def stop_trace(*_a, **_k):
    return None
