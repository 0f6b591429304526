"""src/truncation_parameters.jl:1-14 (oracle; test-only)."""
import sys

DEFAULT_MAXDIM = sys.maxsize
DEFAULT_MINDIM = 1
DEFAULT_CUTOFF = 0.0


def get_or_last(x, i):
    """`get_or_last(x, i) = (i >= length(x)) ? last(x) : x[i]` with 1-based sweep index i."""
    if isinstance(x, (list, tuple)):
        return x[-1] if i >= len(x) else x[i - 1]
    return x


def truncation_parameters(sweep, *, cutoff=DEFAULT_CUTOFF, maxdim=DEFAULT_MAXDIM, mindim=DEFAULT_MINDIM):
    return dict(cutoff=get_or_last(cutoff, sweep), mindim=get_or_last(mindim, sweep),
                maxdim=get_or_last(maxdim, sweep))
