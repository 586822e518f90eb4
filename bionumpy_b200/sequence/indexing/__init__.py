from .kmer_indexing import KmerIndex, KmerLookup
