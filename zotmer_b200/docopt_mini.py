"""
A small docopt-compatible parser (docopt itself is not available offline).

Covers the grammar features the `zot` commands use: a single usage pattern with
`[options]`, explicit option groups such as `[-c CUTOFF]`, `[-M measure]...` (repeatable) and stacked
shorts `[-abp P]`, positional `<name>` / `<name>...` / `[<args>...]`, an "Options:" section with
`-x ARG` / `--long` descriptions and `[default: V]`, plus `options_first`.

Deliberate difference from docopt 0.6: every option DESCRIBED in the options section is present in
the result even when the usage pattern does not mention it (real docopt omits those, which makes the
reference's `zot trim` raise KeyError('-C') on every run -- zotmer/commands/trim.py:3,75).
"""
import re
import sys


class DocoptExit(SystemExit):
    usage = ""

    def __init__(self, message=""):
        SystemExit.__init__(self, (message + "\n" + self.usage).strip())


class _Opt(object):
    def __init__(self, short=None, long=None, takes_arg=False, default=None):
        self.short, self.long, self.takes_arg, self.default = short, long, takes_arg, default
        self.repeat = False

    @property
    def name(self):
        return self.long or self.short


def _section(doc, name):
    out = []
    take = False
    for line in doc.split("\n"):
        if re.match(r"^\s*%s:" % name, line, re.I):
            take = True
            rest = line.split(":", 1)[1]
            if rest.strip():
                out.append(rest)
            continue
        if take:
            if line.strip() == "" and out:
                if name.lower() == "usage":
                    break
                out.append("")
                continue
            if re.match(r"^\S.*:\s*$", line):  # next section header
                break
            out.append(line)
    return out


def _parse_option_descriptions(doc):
    opts = []
    lines = _section(doc, "options")
    i = 0
    while i < len(lines):
        line = lines[i]
        m = re.match(r"^\s*(-\S.*)$", line)
        if m:
            desc_block = [line]
            j = i + 1
            while j < len(lines) and not re.match(r"^\s*-\S", lines[j]):
                desc_block.append(lines[j])
                j += 1
            text = " ".join(desc_block)
            spec = re.split(r"\s{2,}", m.group(1).strip(), maxsplit=1)[0]
            o = _Opt()
            for tok in spec.replace(",", " ").replace("=", " ").split():
                if tok.startswith("--"):
                    o.long = tok
                elif tok.startswith("-"):
                    o.short = tok
                else:
                    o.takes_arg = True
            d = re.search(r"\[default: (.*?)\]", text, re.I)
            if o.takes_arg and d:
                o.default = d.group(1)
            opts.append(o)
            i = j
        else:
            i += 1
    return opts


def docopt(doc, argv=None, help=True, version=None, options_first=False):
    argv = list(sys.argv[1:] if argv is None else argv)
    usage_lines = _section(doc, "usage")
    DocoptExit.usage = "Usage:\n" + "\n".join(usage_lines)
    pattern = " ".join(usage_lines).split()
    # drop program name and sub-command words (everything before the first [ < or -)
    k = 0
    while k < len(pattern) and not re.match(r"^[\[<\-(]", pattern[k]):
        k += 1
    words = pattern[:k]
    pattern = pattern[k:]
    opts = _parse_option_descriptions(doc)
    by_name = {}
    for o in opts:
        if o.short:
            by_name[o.short] = o
        if o.long:
            by_name[o.long] = o

    # ---- walk the usage pattern: positionals, and options that only appear there (e.g. -b of jaccard)
    positionals = []  # (name, required, repeat)
    text = " ".join(pattern)
    for grp, rep in re.findall(r"\[(-[^\]]*)\](\.\.\.)?", text):
        toks = grp.split()
        first = toks[0]
        if first.startswith("--"):
            names = [first]
        else:
            names = ["-" + ch for ch in first[1:]]
        for nm in names:
            if nm not in by_name:
                o = _Opt(short=nm if not nm.startswith("--") else None, long=nm if nm.startswith("--") else None)
                by_name[nm] = o
                opts.append(o)
        if len(toks) > 1:  # the last stacked short takes the argument
            by_name[names[-1]].takes_arg = True
        if rep:
            for nm in names:
                by_name[nm].repeat = True
    stripped = re.sub(r"\[(-[^\]]*)\](\.\.\.)?", " ", text).replace("[options]", " ")
    for m in re.finditer(r"(\[)?\s*(<[^>]+>)\s*(\.\.\.)?\s*(\])?\s*(\.\.\.)?", stripped):
        positionals.append((m.group(2), m.group(1) is None, bool(m.group(3) or m.group(5))))

    result = {}
    for o in opts:
        if o.repeat:
            result[o.name] = []
        elif o.takes_arg:
            result[o.name] = o.default
        else:
            result[o.name] = False
    if help and "--help" not in by_name:
        pass

    # ---- scan argv
    args = []
    i = 0
    seen_positional = False
    while i < len(argv):
        tok = argv[i]
        if tok == "--":
            args.extend(argv[i + 1:])
            break
        if options_first and seen_positional:
            args.append(tok)
            i += 1
            continue
        if tok.startswith("--") and len(tok) > 2:
            nm, eq, val = tok.partition("=")
            if help and nm == "--help":
                print(doc.strip("\n"))
                sys.exit(0)
            if version is not None and nm == "--version":
                print(version)
                sys.exit(0)
            o = by_name.get(nm)
            if o is None:
                raise DocoptExit("%s is not recognized" % nm)
            if o.takes_arg:
                if not eq:
                    i += 1
                    if i >= len(argv):
                        raise DocoptExit("%s requires argument" % nm)
                    val = argv[i]
                _store(result, o, val)
            else:
                _store(result, o, True)
        elif tok.startswith("-") and len(tok) > 1:
            j = 1
            while j < len(tok):
                nm = "-" + tok[j]
                if help and nm == "-h" and nm not in by_name:
                    print(doc.strip("\n"))
                    sys.exit(0)
                if version is not None and nm == "-V" and by_name.get(nm) is not None and by_name[nm].long == "--version":
                    print(version)
                    sys.exit(0)
                o = by_name.get(nm)
                if o is None:
                    raise DocoptExit("%s is not recognized" % nm)
                if o.takes_arg:
                    val = tok[j + 1:]
                    if not val:
                        i += 1
                        if i >= len(argv):
                            raise DocoptExit("%s requires argument" % nm)
                        val = argv[i]
                    _store(result, o, val)
                    break
                _store(result, o, True)
                j += 1
        else:
            args.append(tok)
            seen_positional = True
        i += 1

    # ---- bind positionals
    ai = 0
    for w in words[1:]:                     # sub-command literal, e.g. "zot kmerize ..."
        if ai >= len(args) or args[ai] != w:
            raise DocoptExit()
        result[w] = True
        ai += 1
    for idx, (nm, required, repeat) in enumerate(positionals):
        remaining_required = sum(1 for p in positionals[idx + 1:] if p[1])
        if repeat:
            take = len(args) - ai - remaining_required
            if required and take < 1:
                raise DocoptExit()
            result[nm] = args[ai:ai + max(take, 0)]
            ai += max(take, 0)
        else:
            if ai < len(args) - remaining_required or (required and ai < len(args)):
                result[nm] = args[ai]
                ai += 1
            elif required:
                raise DocoptExit()
            else:
                result[nm] = None
    if ai != len(args):
        raise DocoptExit()
    return result


def _store(result, o, val):
    if o.repeat:
        result[o.name].append(val)
    else:
        result[o.name] = val
