"""Packaging of zotmer_b200: the `zot` console script of the reference (setup.py:25-27 there: zot=zotmer.cli:main)
over this package's dispatcher.  The CUDA library is built in-tree first (make -C zotmer_b200/csrc, or
`python -c "import __graft_entry__ as g; g.build()"`) and shipped as package data; there is no CPU fallback, so an
install without libzot_b200.so fails at the first command with a message that says how to build it."""
from setuptools import setup, find_packages

setup(name='zotmer_b200',
      version='0.2',
      description='B200 (sm_100a) implementation of the zotmer k-mer hot path: kmerize, merge, dist, jaccard, trim, hist',
      packages=find_packages(include=['zotmer_b200', 'zotmer_b200.*']),
      package_data={'zotmer_b200': ['libzot_b200.so']},
      install_requires=['numpy'],
      entry_points={'console_scripts': ['zot=zotmer_b200.cli:main']},
      zip_safe=False)
