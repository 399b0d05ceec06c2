# A k-mer set file = a casket (library/casket.py) whose entry '__meta__' holds the JSON metadata (K, stream names, count
# histogram, acgt frequencies, number of records) -- zotmer/library/kmers.py:8-21.  Reading parses the metadata up front;
# a container opened for writing appends it when it is closed.
import json

from zotmer_b200.library.casket import casket

META = '__meta__'


class kmers(casket):
    def __init__(self, fn, mode):
        casket.__init__(self, fn, mode)
        self.meta = json.loads(self.open(META).read()) if mode == 'r' else {}

    def close(self):
        writing = self.mode == 'w' and self.fo is not None
        if writing:
            self.add_content(META, json.dumps(self.meta))
        casket.close(self)
