# The `zot` dispatcher: same command line as zotmer/cli.py:21-59 (`zot [options] <command> [<args>...]`, `zot help`,
# `zot help <command>`, temp files of library.file.tmpfile removed on the way out) over a static table of the commands
# this package implements; the grammars live in zotmer_b200/usage.py.
import importlib
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200 import usage
from zotmer_b200.library.file import autoremove

__doc__ = usage.CLI
VERSION = 'Zotmer k-mer toolkit 0.1'
UNKNOWN = "unable to load command `%s', use `zot help` for help."


def _module(name):
    """the module behind a command word, or None (as the reference: any importable zotmer.commands.<name>)"""
    if name not in usage.USAGE:
        return None
    return importlib.import_module('zotmer_b200.commands.' + name)


def _overview():
    lines = [__doc__, "Available commands:"]
    lines += ['\t' + name for name in sorted(usage.USAGE)]
    lines.append('\nuse "zot help <command>" for command specific help.')
    return '\n'.join(lines)


def mainInner(argv=None):
    parsed = docopt.docopt(__doc__, argv, version=VERSION, options_first=True)
    word, rest = parsed['<command>'], parsed['<args>']
    if word == 'help':
        if len(rest) != 1:
            print(_overview())
            return 0
        mod = _module(rest[0])
        if mod is None:
            print(UNKNOWN % (word,), file=sys.stderr)      # sic: the reference names `help` here, not the argument
            return 1
        print(mod.__doc__)
        return 0
    mod = _module(word)
    if mod is None:
        print(UNKNOWN % (word,), file=sys.stderr)
        return 1
    return mod.main([word] + rest)


def main(argv=None):
    try:
        with autoremove():
            mainInner(argv)
    except KeyboardInterrupt:
        pass


if __name__ == '__main__':
    main()
