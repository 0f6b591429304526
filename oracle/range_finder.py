"""Randomised range finder (oracle; test-only).  Restates
src/sketched_linear_algebra/range_finder.jl:6-47 and :54-64.  Vectors are numpy arrays."""
import sys

import numpy as np


def range_finder(sample_from_range, *, cutoff=0.0, domain_size=sys.maxsize, max_rank=sys.maxsize,
                 orthogonal_threshold=1e-12, oversample=2, range_size=None, north_pass=2):
    range_vectors = []
    if max_rank <= 0:
        return range_vectors
    if range_size is None:
        q = sample_from_range()
        qnorm = np.linalg.norm(q)
        if qnorm < orthogonal_threshold:
            return range_vectors
        range_vectors = [q / qnorm]
        range_size = q.size
    max_rank = min(max_rank, range_size, domain_size)
    sketch_size = min(max_rank + oversample, range_size, domain_size)
    if sketch_size <= 0:
        return range_vectors
    for k in range(len(range_vectors) + 1, sketch_size + 1):
        q = sample_from_range()
        for _ in range(north_pass):
            for qprev in range_vectors:
                q = q - np.vdot(qprev, q) * qprev
        qnorm = np.linalg.norm(q)
        if qnorm < orthogonal_threshold:
            break
        q = q / qnorm
        range_vectors.append(q)
        if qnorm < cutoff:
            break
    return range_vectors


def range_finder_map(linear_map, random_vector, *, domain_size=sys.maxsize, **kws):
    v = random_vector()
    vsize = v.size
    if domain_size < sys.maxsize and vsize != domain_size:
        raise ValueError(
            f"length of random_vector() output (={v.size}) should equal domain_size (={domain_size})")
    return range_finder(lambda: linear_map(random_vector()), domain_size=vsize, **kws)
