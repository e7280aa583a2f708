"""Deterministic test files for the pager (more(), nuts333.c:2205-2322)."""
import random

TOKENS = [b"~OL", b"~FR", b"~RS", b"~BB", b"/~FG", b"~", b"~XX", b"/", b"word", b"help", b" ", b" ", b"text", b"~FT", b"~UL"]


def make_file(seed: int, n_lines: int, trailing_newline: bool = True, long_line_at: int = -1) -> bytes:
    rng = random.Random(seed)
    lines = []
    for i in range(n_lines):
        if i == long_line_at:
            # longer than fgets(text,1999): split into chunks, with "/~" and "~FR" across the seams
            body = (b"x" * 1996 + b"/~" + b"y" * 1995 + b"~FR" + b"tail of the long line")
        else:
            k = rng.choice([0, 1, 3, 6, 10, 14, 20, 30])
            body = b"".join(rng.choice(TOKENS) for _ in range(k))
        lines.append(body)
    data = b"\n".join(lines)
    if trailing_newline and n_lines:
        data += b"\n"
    return data


CASES = [
    dict(seed=1, n_lines=60, trailing_newline=True),       # three pages
    dict(seed=2, n_lines=10, trailing_newline=False),      # last line never shown (feof)
    dict(seed=3, n_lines=40, trailing_newline=True, long_line_at=5),
    dict(seed=4, n_lines=0),                               # empty file
    dict(seed=5, n_lines=23, trailing_newline=True),
    dict(seed=6, n_lines=1, trailing_newline=False),
]
