import contextlib


@contextlib.contextmanager
def gather_params_ctx(module, modifier_rank=0):
    yield
