#! /usr/bin/env python3

"""Minimal single-file scanner: prints ``line_number:line`` for every match (reference hypergrep/scanner.py)."""

import argparse

import hypergrep_b200 as hypergrep


def on_match(matches: list, count: int) -> None:
    """Batch callback from the native engine: 0-based line number, then the line without its line ending."""
    for position in range(count):
        record = matches[position]
        print(f"{record.line_number}:{record.line.decode(errors='ignore').rstrip()}")


def parse_args() -> argparse.Namespace:
    """Command line of the demo scanner."""
    parser = argparse.ArgumentParser()
    parser.add_argument("pattern", help="Regular expression to use.")
    parser.add_argument("file", help="File to process: plain text, gzip or zstd.")
    return parser.parse_args()


def main() -> None:
    """Scan one file with one pattern."""
    args = parse_args()
    hypergrep.scan(args.file, [args.pattern], on_match)


if __name__ == "__main__":
    main()
