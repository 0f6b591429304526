"""BASELINE config 3 (Hubbard chain with QN conservation) for bench.py --config 3; filled in with the block-sparse path."""


def run_config3(args, emit, _line, ClockSampler, roofline_from_profile, pinned_array):
    raise SystemExit("bench.py --config 3: not available in this build")
