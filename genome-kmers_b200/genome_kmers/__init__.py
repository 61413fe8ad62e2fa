"""
genome_kmers -- B200-native drop-in for the hot path of mrperkett/genome-kmers.

    from genome_kmers.sequence_collection import SequenceCollection
    from genome_kmers.kmers import Kmers

Same import names and call signatures as the reference package; Kmers.sort() and the k-mer
group counting run as CUDA kernels for sm_100a (libgkb200.so, see include/gkb200.h).
"""
__version__ = "0.1.0"
