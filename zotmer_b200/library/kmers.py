"""k-mer set file = casket + JSON `__meta__` (mirrors zotmer/library/kmers.py:8-21)."""
import json

from zotmer_b200.library.casket import casket


class kmers(casket):
    def __init__(self, fn, mode):
        super(kmers, self).__init__(fn, mode)
        self.meta = {}
        if mode == 'r':
            self.meta = json.loads(self.open('__meta__').read())

    def close(self):
        if self.fo is not None and self.mode == 'w':
            self.add_content('__meta__', json.dumps(self.meta))
        super(kmers, self).close()
