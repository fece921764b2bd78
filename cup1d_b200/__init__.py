"""cup1d_b200: the batched Lyman-alpha P1D log-likelihood of igmhub/cup1d on NVIDIA B200.

    from cup1d_b200 import B200Likelihood          # mirrors cup1d's Likelihood on the hot path
    from cup1d_b200.attach import attach, detach   # rebinds those methods on a live reference object

Names are resolved lazily so that importing the package needs neither torch nor the CUDA library.
"""

__all__ = ["B200Likelihood", "B200Theory"]


def __getattr__(name):
    if name == "B200Likelihood":
        from .likelihood import B200Likelihood

        return B200Likelihood
    if name == "B200Theory":
        from .theory import B200Theory

        return B200Theory
    raise AttributeError("module 'cup1d_b200' has no attribute %r" % name)
