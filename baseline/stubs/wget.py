"""Stand-in for the `wget` PyPI module, which the reference imports at package import time
(pygrank/benchmarks/download.py:3) but which is not installed here (no network either)."""


def download(*args, **kwargs):
    raise RuntimeError("no network: dataset download is unavailable in this environment")
