class MismatchedK(Exception):
    """zotmer/library/exceptions.py: raised by `zot dist` when a file's K is below the requested K."""

    def __init__(self, k1, k2):
        self.k1 = k1
        self.k2 = k2

    def __str__(self):
        return 'incompatible values of K: %d & %d' % (self.k1, self.k2)
