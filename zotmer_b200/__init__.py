"""
zotmer_b200 -- B200-native implementation of the k-mer hot path of drtconway/zotmer.

Layout mirrors the reference: `cli` (the `zot` dispatcher), `commands/` (one module per
sub-command, the module docstring is its docopt grammar) and `library/` (container format, stream
files, k-mer basics, distance formulas).  All per-k-mer work is done by libzot_b200.so (sm_100a
CUDA kernels behind the C ABI in include/zotmer_b200.h), reached through `_native`.
"""
__version__ = "0.1"
