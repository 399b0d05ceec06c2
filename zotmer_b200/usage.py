"""
The docopt grammars of the `zot` commands this package implements -- the externally visible command surface of
zotmer/cli.py and zotmer/commands/<command>.py, kept character for character so that every invocation the reference
accepts parses to the same options here.  A command module sets its own __doc__ from this table (`zot help <command>`
prints it, docopt parses it).
"""

CLI = """
Usage:
    zot [options] <command> [<args>...]

options:
    --help          print usage information
    -V, --version   print version information
"""

KMERIZE = """
Usage:
    zot kmerize [options] <k> <output> <input>...

Kmerize FASTA or FASTQ inputs to produce a standard container object.

Arguments:
    <k>         the length of the k-mers. Recommended values: 10-30
    <output>    the name of the output file.
                recommended naming convention
                    - mykmers.k25 for a k-mer set of 25-mers
                    - mykmers.kf25 for a k-mer frequency set of 25-mers
                    - mykmers.e25 for an expanded k-mer set of 25-mers

Options:
    -m MEM      in-memory buffer size (in MB)
    -C BAITS    capture mode - use kmers from the given FASTA file.
    -D FRAC     subsample k-mers, using FRAC proportion of k-mers
    -S SEED     if -D is given, give a seed for determining the
                subspace (defaults to 0).
    -v          produce verbose progress messages
"""

MERGE = """
Usage:
    zot merge <output> <input>...
"""

DIST = """
Usage:
    zot dist [-M measure]... <k> <input>...

Options:
    -M measure  use "measure" for the distance between k-mer frequency sets.
                Use "-M list" to get a list of available measures.
"""

JACCARD = """
Usage:
    zot jaccard [-abp P] <input>...

Compute Jaccard indexes between k-mer sets. By default, indexes are
computed only between the first k-mer set and all the remaining
k-mer sets. If the -a option is given, all pairwise indexes are
computed.  If the -p P option is given, a Null hypothesis test is
performed for the hypothesis that the underlying Jaccard Index is
less than P. This is particularly useful if subsets of k-mers are
being used (NB, if the k-mer sets are large, the statistics can be
very expensive to compute).

Options:
    -a          print all pairwise distances
    -p P        Jaccard distance thresshhold for p-value computation
"""

TRIM = """
Usage:
    zot trim [-c CUTOFF] <output> <input>

Options:
    -c CUTOFF   discard k-mers with frequency less than CUTOFF. A
                cutoff of 0 (the default) indicates that cutoff
                inference should be used. [default: 0]
    -C CUTOFF   discard k-mers with frequency greater than CUTOFF.
                A cutoff of 0 (the default) indicates that the
                cutoff value should be effectively infinite.
                [default: 0]
"""

HIST = """
Usage:
    zot hist <input>...

Options:
    -u              update the input container to include the histogram
"""

INFO = """
Usage:
    zot info <input>...
"""

DUMP = """
Usage:
    zot dump <input>
"""

SAMPLE = """
Usage:
    zot sample [-DS SEED] [-P PROBABILITY] <output> <input>

Options:
    -D              use deterministic sampling
    -P PROBABILITY  the proportion of samples to include in the output.
                    default: 0.01
    -S SEED         use the given seed for the sampling
"""

PROJECT = """
Usage:
    zot project <ref> <output> <input>

Project one or more inputs on to a reference set. For each k-mer in <ref>,
a whitespace separated 0 or a 1 is printed indicating whether that k-mer
was present in the input, with a separate line for each input k-mer set.
"""


USAGE = {"kmerize": KMERIZE, "merge": MERGE, "dist": DIST, "jaccard": JACCARD, "trim": TRIM, "hist": HIST, "info": INFO, "dump": DUMP, "sample": SAMPLE, "project": PROJECT}
