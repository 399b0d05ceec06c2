"""
Run the UNMODIFIED reference sources (Python 2) under Python 3 by translating them in memory.

This is golden-vector tooling only (see make_golden.py).  Nothing is copied into the repo: the
sources are read from /root/reference at generation time, a handful of *syntactic* py2->py3
rewrites are applied to the text in memory (print statement, xrange, .next(), dict views,
integer '/', binary file modes) and the result is exec'd into synthetic modules registered under
the reference's own package names.  No arithmetic line is altered (SURVEY.md section 8c lists the
same rewrite set).  The translated modules are the *reference itself* as far as results go; the
committed fixtures under tests/golden/ are therefore "outputs of the reference run here".

Only usable inside the build container (/root/reference does not exist on the GPU box).
"""
import io
import os
import re
import sys
import types

REF_ROOT = os.environ.get("ZOT_REFERENCE", "/root/reference")

# module name -> path relative to REF_ROOT (only the hot-path files of SURVEY.md section 2.1)
_MODULES = {
    "zotmer.library.bits": "zotmer/library/bits.py",
    "zotmer.library.basics": "zotmer/library/basics.py",
    "zotmer.library.file": "zotmer/library/file.py",
    "zotmer.library.codec64": "zotmer/library/codec64.py",
    "zotmer.library.container.casket": "zotmer/library/container/casket.py",
    "zotmer.library.files": "zotmer/library/files.py",
    "zotmer.library.kmers": "zotmer/library/kmers.py",
    "zotmer.library.misc": "zotmer/library/misc.py",
    "zotmer.library.dist": "zotmer/library/dist.py",
    "zotmer.library.stats": "zotmer/library/stats.py",
    "zotmer.library.exceptions": "zotmer/library/exceptions.py",
    "zotmer.library.timer": "zotmer/library/timer.py",
    "zotmer.library.reads": "zotmer/library/reads.py",
    "zotmer.commands.kmerize": "zotmer/commands/kmerize.py",
    "zotmer.commands.merge": "zotmer/commands/merge.py",
    "zotmer.commands.dist": "zotmer/commands/dist.py",
    "zotmer.commands.jaccard": "zotmer/commands/jaccard.py",
    "zotmer.commands.trim": "zotmer/commands/trim.py",
    "zotmer.commands.hist": "zotmer/commands/hist.py",
    "zotmer.commands.info": "zotmer/commands/info.py",
    "zotmer.commands.dump": "zotmer/commands/dump.py",
    "zotmer.commands.sample": "zotmer/commands/sample.py",
    "zotmer.commands.project": "zotmer/commands/project.py",
}

_PKGS = ["zotmer", "zotmer.library", "zotmer.library.container", "zotmer.commands"]


def _next(o):
    """py2 `o.next()`: classes in the reference define .next(); generators need next()."""
    if hasattr(o, "next"):
        return o.next()
    return next(o)


class _BinFile(object):
    """Binary file whose write() accepts py2-style str (casket TOC / meta JSON)."""

    def __init__(self, fn, mode):
        self._f = io.open(fn, mode + "b")

    def write(self, d):
        if isinstance(d, str):
            d = d.encode("latin-1")
        return self._f.write(d)

    def __getattr__(self, nm):
        return getattr(self._f, nm)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self._f.close()
        return False

    def __iter__(self):
        return iter(self._f)


def _open_bin(fn, mode="r"):
    return _BinFile(fn, mode.replace("b", ""))


def _open_text(fn, mode="r"):
    # py2 str == bytes: latin-1 keeps ord() == byte value; newline='\n' keeps '\r' in the data
    # exactly as py2's non-universal 'r' mode does on POSIX.
    if "r" in mode:
        return io.open(fn, "r", encoding="latin-1", newline="\n")
    return _BinFile(fn, mode.replace("b", ""))


def translate(src):
    out = []
    for line in src.split("\n"):
        m = re.match(r"^(\s*)print\s*>>\s*sys\.stderr\s*,\s*(.*)$", line)
        if m:
            line = "%sprint(%s, file=sys.stderr)" % (m.group(1), m.group(2))
        else:
            m = re.match(r"^(\s*)print\s+(.*)$", line)
            if m:
                line = "%sprint(%s)" % (m.group(1), m.group(2))
            elif re.match(r"^\s*print\s*$", line):
                line = line.replace("print", "print()")
        out.append(line)
    s = "\n".join(out)
    s = re.sub(r"\bxrange\b", "range", s)
    s = re.sub(r"\blong\(", "int(", s)
    s = s.replace(".iteritems()", ".items()")
    # gen.next() -> _next(gen)   (but keep `def next(self)` and `self.next()` inside classes working)
    s = re.sub(r"\b([A-Za-z_][A-Za-z_0-9]*(?:\.[A-Za-z_][A-Za-z_0-9]*)*)\.next\(\)", r"_next(\1)", s)
    # dict views that get .sort()ed / indexed later
    s = re.sub(r"=\s*([A-Za-z_][A-Za-z_0-9\.\[\]']*)\.(items|keys)\(\)\s*$", r"= list(\1.\2())", s, flags=re.M)
    # the single true-division-on-ints site on the path (files.py:59)
    s = s.replace("(len(s) / 8,)", "(len(s) // 8,)")
    # py2 string exception
    s = s.replace('raise "Jensen-Shannon cannot be computed over k-mer lists"',
                  'raise TypeError("Jensen-Shannon cannot be computed over k-mer lists")')
    # casket returns '' at EOF in py2
    s = s.replace("return ''\n", "return b''\n")
    return s


class _Tqdm(object):
    def __init__(self, *a, **k):
        pass

    def update(self, *a):
        pass

    def set_postfix(self, **k):
        pass

    def close(self):
        pass


_loaded = False


def load(docopt_func=None):
    """Install the translated reference under its own module names; returns the `zotmer` package."""
    global _loaded
    if _loaded:
        return sys.modules["zotmer"]
    for p in _PKGS:
        m = types.ModuleType(p)
        m.__path__ = []
        sys.modules[p] = m
    # stand-ins for third-party deps that carry no arithmetic
    d = types.ModuleType("docopt")
    d.docopt = docopt_func if docopt_func is not None else (lambda doc, argv=None, **kw: {})
    sys.modules["docopt"] = d
    t = types.ModuleType("tqdm")
    t.tqdm = _Tqdm
    sys.modules.setdefault("tqdm", t)

    for name, rel in _MODULES.items():
        with io.open(os.path.join(REF_ROOT, rel), "r", encoding="latin-1") as f:
            src = translate(f.read())
        mod = types.ModuleType(name)
        mod.__file__ = os.path.join(REF_ROOT, rel)
        mod.__dict__["_next"] = _next
        if name in ("zotmer.library.container.casket", "zotmer.library.files",
                    "zotmer.commands.kmerize", "zotmer.commands.merge"):
            mod.__dict__["open"] = _open_bin
        elif name == "zotmer.library.file":
            mod.__dict__["open"] = _open_text
        sys.modules[name] = mod
        code = compile(src, mod.__file__, "exec")
        exec(code, mod.__dict__)
        if name == "zotmer.library.reads":
            mod.reads.__next__ = mod.reads.next  # py2 iterator protocol (reads.py:62)
        parent, _, leaf = name.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    for p in _PKGS[1:]:
        parent, _, leaf = p.rpartition(".")
        setattr(sys.modules[parent], leaf, sys.modules[p])
    _loaded = True
    return sys.modules["zotmer"]


def set_docopt(func):
    sys.modules["docopt"].docopt = func
