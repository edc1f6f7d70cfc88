"""The 002 configs import the reference's training-time degradation / cropping / coordinate-sampling
pipeline classes at module level (configs/002_real_wogan...py:9-11).  Those run in the dataloader during
training, which is outside this package's scope; the names exist so that the
config files load unchanged, and say so when somebody tries to use them."""


def placeholder(name, where):
    class _Unavailable:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                f"{name} ({where}) is a training data-pipeline step of the reference; ciaosr_b200 covers "
                "the inference path only (SURVEY.md 8f #4)")
    _Unavailable.__name__ = _Unavailable.__qualname__ = name
    return _Unavailable
