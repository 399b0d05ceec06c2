"""
Usage:
    zot [options] <command> [<args>...]

options:
    --help          print usage information
    -V, --version   print version information
"""
# Mirrors zotmer/cli.py:1-63: docopt with options_first, `zot help`, dynamic import of
# zotmer_b200.commands.<command>, temp files removed on exit.
import importlib
import pkgutil
import sys

from zotmer_b200 import docopt_mini as docopt
from zotmer_b200.library.file import autoremove
from zotmer_b200 import commands


def mainInner(argv=None):
    args = docopt.docopt(__doc__, argv, version='Zotmer k-mer toolkit 0.1', options_first=True)

    if args['<command>'] == 'help' and len(args['<args>']) != 1:
        print(__doc__)
        print("Available commands:")
        for _, name, is_pkg in pkgutil.iter_modules([commands.__path__[0]]):
            print('\t' + name)
        print('\nuse "zot help <command>" for command specific help.')
        return 0

    if args['<command>'] == 'help' and len(args['<args>']) == 1:
        modname = commands.__name__ + '.' + args['<args>'][0]
        try:
            m = importlib.import_module(modname)
            print(m.__doc__)
            return 0
        except ImportError:
            print("unable to load command `%s', use `zot help` for help." % (args['<command>'],), file=sys.stderr)
            return 1

    modname = commands.__name__ + '.' + args['<command>']
    try:
        m = importlib.import_module(modname)
    except ImportError:
        print("unable to load command `%s', use `zot help` for help." % (args['<command>'],), file=sys.stderr)
        return 1
    argv = [args['<command>']] + args['<args>']
    return m.main(argv)


def main(argv=None):
    try:
        with autoremove():
            mainInner(argv)
    except KeyboardInterrupt:
        pass


if __name__ == '__main__':
    main()
