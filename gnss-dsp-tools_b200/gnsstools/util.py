"""Command-line list parsers (reference: gnsstools/util.py:1-14)."""


def parse_list_ranges(s, sep='-'):
    """'1-3,7' -> [1, 2, 3, 7]; ranges are inclusive, `sep` is the range mark."""
    out = []
    for item in s.split(','):
        ends = item.split(sep)
        if len(ends) == 1:
            out.append(int(ends[0]))
        else:
            out.extend(range(int(ends[0]), int(ends[1]) + 1))
    return out


def parse_list_floats(s):
    """'-7000,7000,200' -> [-7000.0, 7000.0, 200.0]."""
    return [float(tok) for tok in s.split(',')]
