#! /usr/bin/env python3

"""grep-style command line over the GPU scan engine.

Mirror of the reference CLI (reference hypergrep/multiscanner.py): same options, same BRE/GNU pattern rewrites,
same output formats and exit codes (0 match, 1 no match, 2 error, 130 interrupted), one ``grep`` job per file.
Files are spread over the visible GPUs by the native library (one device per ``hyperscan`` call, round-robin),
which replaces the reference's "one CPU thread per file" scaling.
"""

from __future__ import annotations

import argparse
import multiprocessing
import re
import sys
import textwrap
from multiprocessing.pool import ThreadPool
from typing import Any, Iterable, Iterator

import hypergrep_b200 as hypergrep

_UNSUPPORTED_URL = "https://intel.github.io/hyperscan/dev-reference/compilation.html#unsupported-constructs"
# Characters that are literals in POSIX basic regular expressions unless escaped (and operators when escaped).
_BRE_SWAPPED = "+?(){}|"


def _grep_with_index(index: int, args: Iterable, kwargs: dict[str, Any]) -> tuple[int, Any]:
    """Pool job: run one grep and tag the outcome (result or exception) with its position in the file list."""
    try:
        outcome = hypergrep.grep(*args, **kwargs)
    except Exception as error:  # pylint: disable=broad-except
        outcome = error
    return index, outcome


def get_argparse_files(args: argparse.Namespace) -> list[str]:
    """Files named on the command line (reference multiscanner.py:27-43).

    As with GNU grep, once -e/-f supplied a pattern the bare positional "pattern" is really the first file.
    """
    files: list[str] = []
    if args.pattern and (args.patterns or args.pattern_files):
        files.append(args.pattern)
    files.extend(args.files or [])
    return files


def _validate_python_regex(pattern: str) -> None:
    try:
        re.compile(pattern)
    except Exception as error:
        raise ValueError(f"hyperscanner: invalid regex: {error}") from error


def get_argparse_patterns(args: argparse.Namespace) -> list[str]:
    """Patterns from the positional, -e and -f options, validated (reference multiscanner.py:46-83).

    Raises ValueError for patterns Python's ``re`` rejects and for patterns the engine's compiler rejects.
    """
    patterns: list[str] = []
    if args.patterns:
        patterns += args.patterns
    elif args.pattern and not args.pattern_files:
        patterns.append(args.pattern)
    for name in args.pattern_files or []:
        with open(name, "rt", encoding="utf-8") as handle:
            patterns += [line.rstrip("\n") for line in handle.readlines()]
    for pattern in patterns:
        _validate_python_regex(pattern)
    # The compile check runs in the parent, before any pool exists, and never touches CUDA (fork-safe).
    if hypergrep.check_compatibility(patterns):
        raise ValueError(f"hyperscanner: incompatible regex: for more information visit {_UNSUPPORTED_URL}")
    return patterns


class _OrderedPrinter:  # pylint: disable=too-many-instance-attributes
    """Collects per-file outcomes from the pool and prints them, in file order unless told otherwise."""

    def __init__(self, files: list, options: dict[str, Any]) -> None:
        self.files = files
        self.opt = options
        self.waiting: dict[int, Any] = {}
        self.cursor = 0
        self.total = 0
        self.matched = False
        self.errored = False

    def __call__(self, tagged: tuple[int, Any]) -> None:
        index, outcome = tagged
        if self.opt["ordered_results"] and index != self.cursor:
            self.waiting[index] = outcome
            return
        self._report(index, outcome)

    def _advance(self) -> None:
        self.cursor += 1
        if self.cursor in self.waiting:
            self(((self.cursor), self.waiting.pop(self.cursor)))

    def _report(self, index: int, outcome: Any) -> None:
        name = self.files[index]
        if isinstance(outcome, Exception):
            print(f"hyperscanner: {name}: {outcome}")
            self.errored = True
            self._advance()
            return
        found, code = outcome
        if code:
            self.errored = True
        if found:
            self.matched = True
            if self.opt["quiet"]:
                return
        opt = self.opt
        if opt["files_without_match"]:
            if not found:
                print(name)
        elif opt["files_with_matches"]:
            if found:
                print(name)
        elif opt["total_results"]:
            self.total += found
        elif opt["count_results"]:
            print(f"{name}:{found}" if opt["with_file_name"] else f"{found}")
        else:
            try:
                print_results(found, name, with_file_name=opt["with_file_name"], with_line_number=opt["with_line_number"])
            except BrokenPipeError:
                pass  # e.g. piped into `head`; keep draining the pool
        self._advance()


def parallel_grep(  # pylint: disable=too-many-arguments,too-many-locals
    files: list,
    patterns: list[str],
    ignore_case: bool = False,
    ordered_results: bool = True,
    count_results: bool = False,
    total_results: bool = False,
    with_file_name: bool = False,
    with_line_number: bool = False,
    use_multithreading: bool = True,
    only_matching: bool = False,
    no_messages: bool = False,
    max_match_count: int = 0,
    files_without_match: bool = False,
    files_with_matches: bool = False,
    quiet: bool = False,
) -> int:
    """Scan files and print grep-formatted results (reference multiscanner.py:86-223).

    Returns the grep exit code: 2 if any file errored, else 1 if nothing matched, else 0.
    """
    if files_without_match or files_with_matches or quiet:
        max_match_count = 1  # these modes only need to know whether a file matches at all
    printer = _OrderedPrinter(
        files,
        {
            "ordered_results": ordered_results,
            "count_results": count_results,
            "total_results": total_results,
            "with_file_name": with_file_name,
            "with_line_number": with_line_number,
            "files_without_match": files_without_match,
            "files_with_matches": files_with_matches,
            "quiet": quiet,
        },
    )
    grep_kwargs = {
        "ignore_case": ignore_case,
        "count_only": count_results or total_results,
        "only_matching": only_matching,
        "no_messages": no_messages,
        "max_match_count": max_match_count,
    }
    workers = min(max(multiprocessing.cpu_count() - 1, 1), len(files))
    pool_type = ThreadPool if use_multithreading else multiprocessing.Pool
    with pool_type(processes=workers) as pool:
        pending = [
            pool.apply_async(_grep_with_index, (index, (name, patterns), grep_kwargs), callback=printer)
            for index, name in enumerate(files)
        ]
        for job in pending:
            job.get()
            if quiet and printer.matched:
                pool.terminate()
                break
    if total_results:
        print(printer.total)
    if printer.errored:
        return 2
    return 0 if printer.matched else 1


def print_results(results: list, file_name: str, with_file_name: bool = False, with_line_number: bool = False) -> None:
    """Print matched lines with the requested prefixes (reference multiscanner.py:226-255)."""
    prefix = f"{file_name}:" if with_file_name else ""
    if with_line_number:
        sys.stdout.write("".join(f"{prefix}{number}:{text}" for number, text in results))
    else:
        sys.stdout.write("".join(f"{prefix}{text}" for _, text in results))


def read_stdin() -> Iterator[str]:
    """File names piped on stdin, one per line, until the first blank line or EOF."""
    for raw in iter(sys.stdin.readline, ""):
        name = raw.strip()
        if not name:
            return
        yield name


def to_basic_regular_expressions(patterns: list[str]) -> list[str]:
    """Rewrite PCRE/ERE syntax into what the same text means as a POSIX BRE (reference multiscanner.py:273-305).

    In a BRE the characters ``+?(){}|`` are literals and their escaped forms are the operators, so escaped and
    unescaped occurrences trade places.  Raises ValueError if the rewritten pattern is not a valid regex.
    """
    converted = []
    for pattern in patterns:
        out: list[str] = []
        for char in pattern:
            if char in _BRE_SWAPPED:
                if out and out[-1] == "\\":
                    out[-1] = char  # escaped in the input: becomes the bare operator
                else:
                    out.append("\\" + char)  # bare in the input: a literal in BRE
            else:
                out.append(char)
        rewritten = "".join(out)
        _validate_python_regex(rewritten)
        converted.append(rewritten)
    return converted


def to_gnu_regular_expressions(patterns: list[str]) -> list[str]:
    r"""Translate GNU word-edge escapes ``\<`` and ``\>`` to ``\b`` (reference multiscanner.py:308-328)."""
    return [re.sub(r"(?<!\\)\\[<>]", lambda _match: r"\b", pattern) for pattern in patterns]


_DESCRIPTION = """\
GPU grep (Global Regular Expression Print) for large log files.

Multi-pattern regex matching on NVIDIA B200 GPUs behind a grep-compatible command line:
    1. All patterns are compiled into one automaton set and matched in a single pass over the file.
    2. Files are read and decompressed (gzip, zstd) on host threads and scanned on the GPU.
    3. Several files are processed concurrently and spread over all visible GPUs.

Differences from standard "grep":
    1. Only the options listed here are supported.
    2. Patterns must avoid constructs that need backtracking (look-around, back-references, ...).
       More details: """ + _UNSUPPORTED_URL + """

Examples:
    hyperscanner <regex> <file(s)>
    find <args> | hyperscanner <regex>"""


def parse_args(args: list | None = None) -> argparse.Namespace:
    """Parse the command line; option names and destinations follow the reference (multiscanner.py:331-548)."""
    parser = argparse.ArgumentParser(
        formatter_class=argparse.RawTextHelpFormatter, add_help=False, description=textwrap.dedent(_DESCRIPTION)
    )
    parser.add_argument("pattern", nargs="?", help="Regex pattern to use.")
    parser.add_argument("files", nargs="*", help="Files to scan.")

    generic = parser.add_argument_group("Generic Program Information")
    generic.add_argument("--help", action="help", default=argparse.SUPPRESS, help="show this help message and exit")

    syntax = parser.add_argument_group("Pattern Syntax").add_mutually_exclusive_group()
    syntax.set_defaults(regexp="bre")
    for short, long_name, value, text in (
        ("-E", "--extended-regexp", "ere", "Interpret PATTERNS as extended regular expressions (EREs)."),
        ("-G", "--basic-regexp", "bre", "Interpret PATTERNS as basic regular expressions. This is the default."),
        ("-P", "--perl-regexp", "pcre", "Interpret PATTERNS as Perl-compatible regular expressions (PCREs)."),
    ):
        syntax.add_argument(short, long_name, dest="regexp", action="store_const", const=value, help=text)

    matching = parser.add_argument_group("Matching Control")
    matching.add_argument("-e", "--regexp", action="append", dest="patterns", metavar="pattern",
                          help="Use PATTERNS as the patterns; may be repeated and combined with -f.")
    matching.add_argument("-f", "--file", action="append", dest="pattern_files", metavar="file",
                          help="Obtain patterns from FILE, one per line; may be repeated and combined with -e.")
    matching.add_argument("-i", "--ignore-case", action="store_true", help="Perform case insensitive matching.")

    output = parser.add_argument_group("General Output Control")
    output.add_argument("-c", "--count", action="store_true",
                        help="Suppress normal output; print a count of matching lines for each input file.")
    output.add_argument("-L", "--files-without-match", action="store_true",
                        help="Suppress normal output; print the name of each input file without a match.")
    output.add_argument("-l", "--files-with-matches", action="store_true",
                        help="Suppress normal output; print the name of each input file with a match.")
    output.add_argument("-m", "--max-count", type=int, default=0, help="Stop reading a file after NUM matching lines.")
    output.add_argument("-o", "--only-matching", action="store_true",
                        help="Print only the matched (non-empty) parts of a matching line, one per output line.")
    output.add_argument("-q", "--quiet", "--silent", action="store_true",
                        help="Quiet; write nothing to standard output. Exit with zero status on the first match.")
    output.add_argument("-s", "--no-messages", action="store_true",
                        help="Suppress error messages about nonexistent or unreadable files.")

    prefix = parser.add_argument_group("Output Line Prefix Control")
    name_mode = prefix.add_mutually_exclusive_group()
    name_mode.add_argument("-H", "--with-filename", action="store_true", default=None,
                           help="Print the file name for each match (default with more than one file).")
    name_mode.add_argument("-h", "--no-filename", action="store_true", default=None,
                           help="Suppress the file name prefix (default with a single file).")
    prefix.add_argument("-n", "--line-number", action="store_true",
                        help="Prefix each line of output with the 1-based line number within its input file.")

    selection = parser.add_argument_group("File and Directory Selection")
    selection.add_argument("-a", "--text", action="store_true",
                           help="Accepted for grep compatibility; files are always processed as bytes.")

    extra = parser.add_argument_group("Unique arguments to hyperscanner")
    extra.add_argument("-t", "--total", action="store_true",
                       help="Suppress normal output; print a count of matching lines across all input files.")
    extra.add_argument("--no-gnu", dest="gnu_regexp", action="store_false",
                       help="Disable GNU grep compatibility rewrites of BRE/ERE patterns (\\< and \\> become \\b).")
    extra.add_argument("--no-order", dest="ordered", action="store_false",
                       help="Print results as files finish instead of in file order.")
    extra.add_argument("--no-sort", dest="sort_files", action="store_false",
                       help="Keep the given file order instead of sorting.")
    extra.add_argument("--mp", action="store_false", dest="use_multithreading",
                       help="Use a multiprocessing pool instead of threads.")

    parser.set_defaults(parser=parser)
    return parser.parse_intermixed_args(args=args)


def _usage_exit(args: argparse.Namespace) -> None:
    args.parser.print_usage()
    raise SystemExit(2)


def main() -> None:
    """Entry point of the ``hyperscanner`` command (reference multiscanner.py:551-606)."""
    args = parse_args()
    try:
        patterns = get_argparse_patterns(args)
        if not patterns:
            _usage_exit(args)
        if args.regexp == "bre":
            patterns = to_basic_regular_expressions(patterns)
    except ValueError as error:
        print(error)
        raise SystemExit(2) from error
    if args.gnu_regexp and args.regexp != "pcre":
        patterns = to_gnu_regular_expressions(patterns)

    files = get_argparse_files(args) or list(read_stdin())
    if args.sort_files:
        files = sorted(files)
    if not files:
        _usage_exit(args)

    if args.no_filename is not None:
        with_file_name = False
    elif args.with_filename is not None:
        with_file_name = True
    else:
        with_file_name = len(files) > 1

    raise SystemExit(
        parallel_grep(
            files=files,
            patterns=patterns,
            ignore_case=args.ignore_case,
            ordered_results=args.ordered,
            count_results=args.count,
            total_results=args.total,
            with_file_name=with_file_name,
            with_line_number=args.line_number,
            use_multithreading=args.use_multithreading,
            only_matching=args.only_matching,
            no_messages=args.no_messages,
            max_match_count=args.max_count,
            quiet=args.quiet,
            files_without_match=args.files_without_match,
            files_with_matches=args.files_with_matches,
        )
    )


if __name__ == "__main__":
    try:
        main()
    except KeyboardInterrupt as interrupt:
        raise SystemExit(130) from interrupt
